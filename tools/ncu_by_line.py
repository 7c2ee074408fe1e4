"""Aggregates an ncu SASS source page (csv) by CUDA source line using nvdisasm -g line annotations.
usage: ncu_by_line.py <sass.csv from `ncu -i X --page source --csv`> <nvdisasm -g -c output> <kernel mangled name> [top]"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, dis_txt, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
sass = [(r[ci['Source']].strip(), int(r[ci['Warp Stall Sampling (All Samples)']] or 0),
         int(r[ci['Instructions Executed']] or 0)) for r in rows[2:] if r and r[0].startswith('0x')]
cur_line, cur_fn, seq = None, None, []
for ln in open(dis_txt):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\.text\.(\S+):', ln)
    if m:
        cur_fn = m.group(1)
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m and cur_fn == kname:
        seq.append(cur_line)
assert len(seq) == len(sass), (len(seq), len(sass))
agg = defaultdict(lambda: [0, 0])
for (txt, smp, ex), line in zip(sass, seq):
    agg[line][0] += smp
    agg[line][1] += ex
ts = sum(v[0] for v in agg.values())
te = sum(v[1] for v in agg.values())
print(f"total samples {ts}, warp instructions {te}")
src_cache = {}
def src(line):
    if not line:
        return ""
    f, n = line
    if f not in src_cache:
        import glob
        c = glob.glob(f"pixie_b200/csrc/cuda/{f}")
        src_cache[f] = open(c[0]).read().split("\n") if c else []
    L = src_cache[f]
    return L[n - 1].strip()[:100] if 0 < n <= len(L) else ""
for line, (smp, ex) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * smp / ts:5.1f}% stall  {100 * ex / te:5.1f}% inst  {line}  {src(line)}")
