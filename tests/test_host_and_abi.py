"""CPU-side checks: the host mirror's known answers from the reference's own tests, and that the
C-ABI library loads and exports every symbol include/pixie_cuda.h declares (no compute, no GPU)."""
import ctypes as C
import os

import numpy as np
import pytest

from pixie_b200 import host
from pixie_b200.common import PixieError

KIND = "Z M L H V C S Q T A m l h v c s q t a".split()
NPAR = dict(zip(range(19), [0, 2, 2, 1, 1, 6, 4, 4, 2, 7, 2, 2, 1, 1, 6, 4, 4, 2, 7]))


def to_string(path):
    """`$`(path) (paths.nim:83-117)."""
    c, i, out = path.commands, 0, []
    while i < len(c):
        k = int(c[i])
        i += 1
        vals = []
        for _ in range(NPAR[k]):
            v = float(c[i])
            vals.append(str(int(v)) if v == int(v) else repr(v))
            i += 1
        out.append(KIND[k] + " ".join(vals))
    return " ".join(out)


def test_parse_path_known_answers():  # tests/test_paths.nim:3-49
    assert to_string(host.parsePath("\n  m 1 2 3 4 5 6\n  ")) == "m1 2 l3 4 l5 6"
    assert to_string(host.parsePath("\n  l 1 2 3 4 5 6\n  ")) == "l1 2 l3 4 l5 6"
    s = "m 1 2\n l 3 4\n h 5\n v 6\n c 0 0 0 0 0 0\n q 1 1 1 1\n t 2 2\n a 7 7 7 7 7 7 7\n z\n"
    assert to_string(host.parsePath("\n" + s)) == "m1 2 l3 4 h5 v6 c0 0 0 0 0 0 q1 1 1 1 t2 2 a7 7 7 7 7 7 7 Z"
    assert to_string(host.parsePath("\n" + s.upper().replace("Z", "z"))) == \
        "M1 2 L3 4 H5 V6 C0 0 0 0 0 0 Q1 1 1 1 T2 2 A7 7 7 7 7 7 7 Z"
    host.parsePath("M 0.1E-10 0.1e10 L2+2 L3-3 L0.1E+10-1")
    assert len(host.parsePath("").commands) == 0


def test_parse_path_errors():
    with pytest.raises(PixieError, match="wrong number of parameters"):
        host.parsePath("M 1 2 3")
    with pytest.raises(PixieError, match="unexpected parameters"):
        host.parsePath("M 1 2 z 4")


def test_builder_errors_match_reference():
    p = host.newPath()
    with pytest.raises(PixieError, match="Invalid polygon sides value"):  # paths.nim:638-639
        p.polygon(0, 0, 10, 2)
    with pytest.raises(PixieError, match="negative radius"):  # paths.nim:422-423
        p.arc(0, 0, -1, 0, 1)
    with pytest.raises(PixieError, match="Invalid line dash value"):  # paths.nim:2037-2039
        host.stroke_segments("M 0 0 L 10 10", dashes=(2.0, 0.0))
    with pytest.raises(PixieError):  # tests/test_paths.nim:659-678: huge arc cannot be discretised
        host.stroke_segments("L -16370.0 -18156.0 A 4100 4100 0 1 0 -19670 -14134 Z")


def test_gaussian_kernel_known_answers():  # SURVEY.md 3.4 (derived from internal.nim:17-34)
    for r, total, centre, first5 in [(10, 65278, 5850, [520, 824, 1243, 1787, 2448]),
                                     (20, 65279, 2935, [261, 330, 413, 511, 624]),
                                     (32, 65275, 1837, [163, 190, 219, 252, 288])]:
        k = host.gaussianKernel(r)
        assert len(k) == 2 * r + 1 and int(k.sum()) == total and int(k[r]) == centre and list(k[:5]) == first5
        assert (k == k[::-1]).all()


def test_segments_are_quantised_and_oriented():  # shapesToSegments, paths.nim:1059-1090
    segs = host.fill_segments("M 10.3 5.123 L 40.7 9.999 L 20 33.3337 L 5 5.123 z")
    y = segs.xyxy[:, [1, 3]]
    assert (y * 256 == np.floor(y * 256)).all()          # y quantised to 1/256
    assert (segs.xyxy[:, 1] < segs.xyxy[:, 3]).all()     # at.y < to.y, horizontals dropped
    assert set(segs.winding.tolist()) <= {1, -1}
    assert len(host.fill_segments("M 0 0 L 0 1 L 0 0 Z")) == 2
    assert len(host.fill_segments("M -65.4,9z")) == 0     # the tiger's empty path


def test_stroke_counts():
    # a stroke is a fill of one rectangle per flattened segment plus joins/caps (paths.nim:2041-2080)
    segs = host.stroke_segments("M 10 10 L 50 60 90 90", strokeWidth=10.0)
    assert len(segs) > 8
    assert np.isfinite(segs.xyxy).all()


def test_cuda_library_exports_every_declared_symbol():
    from pixie_b200 import device

    names = device.declared_symbols()
    assert len(names) >= 40 and "pixie_cuda_fill_batch" in names and "pixie_cuda_blur_rows" in names
    lib = device.lib()  # binds every declared symbol; AttributeError if one is missing
    for n in names:
        assert getattr(lib, n) is not None
    assert set(device._SIGNATURES) | {"pixie_cuda_last_error"} == set(names)


def test_host_library_exports_header():
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "pixie_host.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(pixie_host_[a-z0-9_]+)\s*\(", text)))
    lib = C.CDLL(os.path.join(root, "pixie_b200", "libpixie_host.so"))
    for n in names:
        assert getattr(lib, n) is not None


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product fails loudly instead of falling back to the CPU."""
    from pixie_b200 import device

    if device.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(PixieError):
        device.init(0)
    with pytest.raises(PixieError):
        device.DeviceImage(8, 8)


def test_path_batch_packing():
    """PathBatch (the host side of pixie_cuda_cmdlist_create_from_paths): command counts, host fallback for arcs /
    round joins / dashes, and the ctypes struct has the header's layout (80 bytes)."""
    import ctypes

    from pixie_b200 import host
    from pixie_b200.device import PathBatch, PathDesc, _scan_commands

    assert ctypes.sizeof(PathDesc) == 80
    p = host.parsePath("M 1 2 L 3 4 C 1 1 2 2 3 3 Q 1 1 5 5 Z")
    assert _scan_commands(p.commands) == (5, False)
    arc = host.parsePath("M 1 2 A 5 5 0 1 0 9 9 Z")
    assert _scan_commands(arc.commands) == (3, True)
    b = PathBatch()
    b.add_fill(p, None, 0xFF0000FF, 0, 0, 0)
    b.add_fill(arc, None, 0xFF0000FF, 0, 0, 0)
    b.add_stroke(p, None, 2.0, host.RoundCap, host.MiterJoin, 4.0, (), 0xFF0000FF, 0, 0, 0)
    b.add_stroke(p, None, 2.0, host.ButtCap, host.BevelJoin, 4.0, (), 0xFF0000FF, 0, 0, 1)
    b.add_stroke(p, None, 2.0, host.ButtCap, host.BevelJoin, 4.0, (3.0, 1.0), 0xFF0000FF, 0, 0, 1)
    assert [d.kind for d in b.descs] == [0, 2, 2, 1, 2] and b.host_paths == 3
    descs, cmds, raw, rw = b.packed()
    assert len(cmds) == 2 * len(p.commands) and descs[3].begin == len(p.commands) and descs[3].num_commands == 5
    assert descs[1].begin == 0 and descs[1].end == len(host.fill_segments(arc))
    assert raw.shape[0] == len(rw) == descs[4].end
    assert descs[3].layer == 1 and descs[3].line_join == host.BevelJoin
