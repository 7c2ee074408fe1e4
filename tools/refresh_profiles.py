"""gpurun_out/ (scratch) -> profiles/ (tracked): bench lines, launch lists, ncu summaries of the last tools/gpu_round.sh,
and a SASS opcode histogram of every hot kernel of the in-tree pixie_cuda.so (cuobjdump, runs without a GPU).
usage: refresh_profiles.py <gpurun tag, e.g. r2a> <profiles tag, e.g. r02>"""
import collections
import glob
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
src, tag = sys.argv[1], sys.argv[2]


def last_line(path):
    with open(path) as f:
        lines = [ln for ln in f.read().strip().splitlines() if ln.startswith("{")]
    return lines[-1]


for a, b in (("bench.json", "bench_n1.json"), ("bench_ref.json", "bench_reference_arm.json")):
    p = os.path.join(G, f"{src}_{a}")
    if os.path.exists(p):
        with open(os.path.join(P, f"{tag}_{b}"), "w") as f:
            f.write(last_line(p) + "\n")
for p in glob.glob(os.path.join(G, f"{src}_launches_*.csv")) + glob.glob(os.path.join(G, f"{src}_*_metrics.txt")) + \
        glob.glob(os.path.join(G, f"{src}_*_details.txt")):
    shutil.copyfile(p, os.path.join(P, tag + os.path.basename(p)[len(src):]))


def sass_histogram(out):
    """opcode histogram per kernel: what proves tcgen05 (UTCHMMA), TMA (UTMALDG), TMEM (LDTM/STTM) are in the binary"""
    so = os.path.join(ROOT, "pixie_b200", "pixie_cuda.so")
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    hot = re.compile(r"blur_tc|blur_mma|raster_kernel|plan_|partition_kernel|count_kernel|blend_rect_vec4|spread_|draw_smooth|gradient|flatten|stroke|halo_")
    cur, hist = None, collections.OrderedDict()
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = name if hot.search(name) else None
            if cur:
                hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            hist[cur][m.group(1).split(".")[0]] += 1
    with open(out, "w") as f:
        f.write("SASS opcode histogram per kernel (cuobjdump -sass pixie_b200/pixie_cuda.so, sm_100a); opcode = mnemonic before the first dot\n")
        f.write("UTCHMMA = tcgen05.mma kind::f16, UTMALDG = cp.async.bulk.tensor (TMA), LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier\n\n")
        for k, c in hist.items():
            short = re.sub(r"\(.*", "", k)
            if len(k) > 160:
                k = k[:160] + "..."
            f.write(f"{k}\n  total {sum(c.values())}: " + ", ".join(f"{op} {n}" for op, n in c.most_common(28)) + "\n")
            special = {op: n for op, n in c.items() if re.match(r"UTC|UTMA|LDTM|STTM|SYNCS|HMMA|IMMA|LDSM|LDGSTS|REDUX|ATOM|RED", op)}
            if special:
                f.write(f"  special: {special}\n")
            f.write("\n")


sass_histogram(os.path.join(P, f"{tag}_sass_histogram.txt"))
print("profiles refreshed:", sorted(os.listdir(P)))
