"""Tiger from path COMMANDS: device-side flattening + stroking (pixie_cuda_cmdlist_create_from_paths) against the
host flattener (libpixie_host.so) + pixie_cuda_cmdlist_create from segments."""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixie_b200 import device as dev, host, svg as psvg  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev.init(0)
doc = psvg.parseSvg(open(os.path.join(ROOT, "tests", "golden", "tiger.svg")).read(), size, size)
t0 = time.perf_counter()
arrays = psvg.svg_fill_batch(doc).arrays()
t_host = time.perf_counter() - t0
# the C++ flattener alone: the fill_segments / stroke_segments calls of the render loop on pre-parsed paths
calls = []
for d, props in doc.elements:
    if not (props.display and props.opacity > 0):
        continue
    path = host.parsePath(d)
    if props.fill != "none":
        calls.append((host.fill_segments, (path, props.transform)))
    if props.stroke != 0 and props.strokeWidth > 0:
        calls.append((host.stroke_segments, (path, props.transform, props.strokeWidth, props.strokeLineCap, props.strokeLineJoin,
                                            props.strokeMiterLimit, props.strokeDashArray)))
tc = []
for it in range(5):
    t0 = time.perf_counter()
    for fn, a in calls:
        fn(*a)
    tc.append(time.perf_counter() - t0)
t_cpp = statistics.median(tc)
pb = psvg.svg_path_batch(doc)
packed = pb.packed()
img = dev.DeviceImage(size, size)
th, td, tr = [], [], []
for it in range(8):
    dev.sync()
    t0 = time.perf_counter()
    cl = dev.CmdList(size, size, 1, arrays)
    dev.sync()
    th.append(time.perf_counter() - t0)
    del cl
    t0 = time.perf_counter()
    cl = dev.CmdList.from_paths(size, size, 1, pb, packed)
    dev.sync()
    td.append(time.perf_counter() - t0)
    img.fill(0)
    dev.timer_begin()
    cl.run(img)
    tr.append(dev.timer_end())
    del cl
ncmd = sum(d.num_commands for d in pb.descs)
print(f"tiger {size}^2: {len(pb)} paths, {ncmd} commands, {len(packed[1]) * 4} B of commands (+{len(pb) * 80} B of path headers) vs "
      f"{len(arrays['winding']) * 18} B of segments; host paths {pb.host_paths}")
print(f"  host: {len(calls)} fill_segments / stroke_segments calls into libpixie_host.so {t_cpp * 1e3:.2f} ms (whole Python render loop incl. parsePath {t_host * 1e3:.2f} ms); "
      f"cmdlist_create from segments {statistics.median(th[2:]) * 1e3:.3f} ms")
print(f"  device: cmdlist_create_from_paths {statistics.median(td[2:]) * 1e3:.3f} ms (H2D commands, resolve/count/scan/emit/stroke/bounds, 3 small readbacks); run {statistics.median(tr[2:]):.3f} ms")
