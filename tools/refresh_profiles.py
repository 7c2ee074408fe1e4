"""gpurun_out/ (scratch) -> profiles/ (tracked): bench lines, launch list, ncu summaries of the last tools/gpu_round.sh."""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"


def last_line(path):
    with open(path) as f:
        lines = [ln for ln in f.read().strip().splitlines() if ln.startswith("{")]
    return lines[-1]


for src, dst in (("bench_cur.json", f"{tag}_bench_n1.json"), ("bench_cur_ref.json", f"{tag}_bench_reference_arm.json"),
                 ("bench_n2_cur.json", f"{tag}_bench_n2.json")):
    p = os.path.join(G, src)
    if os.path.exists(p):
        with open(os.path.join(P, dst), "w") as f:
            f.write(last_line(p) + "\n")
shutil.copyfile(os.path.join(G, "launches_cur.csv"), os.path.join(P, f"{tag}_launches_bench_tiger.csv"))
for rep, name in (("prof_tiger_cur", "tiger"), ("prof_blur_cur", "blur_mma"), ("prof_shadow_cur", "shadow"), ("prof_blend_cur", "blend"), ("prof_draw_cur", "draw")):
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "summarize_ncu.py"), os.path.join(G, rep + ".ncu-rep"),
                           os.path.join(P, f"{tag}_{name}")])
d = json.loads(last_line(os.path.join(G, "bench_cur.json")))
r = json.loads(last_line(os.path.join(G, "bench_cur_ref.json")))
x = d["extras"]
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "ref", r["value"], r["ms_per_step"])
print("roofline", d["roofline"])
print("blur", x["blur_r32_16384"]["ms"], x["blur_r32_16384"]["x_pass_ms"], x["blur_r32_16384"]["y_pass_ms"], x["blur_r32_16384"]["frac_hbm"],
      "shadow", x["shadow_16384"], "icons", x["icons_512_batch"])
print({k: v["frac_hbm"] for k, v in x["blend_8192_masked"]["modes"].items()})
print({k: (v["ms"], v["frac_hbm"]) for k, v in x["draw_paint_8192"]["ops"].items()})
