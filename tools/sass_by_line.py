"""Per-source-line instruction counts of one kernel from an ncu SASS source page (csv) + nvdisasm -g -c listing.
usage: sass_by_line.py <sass.csv> <nvdisasm listing> <mangled kernel name> [top]  (regions: raster.cu functions)"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, dis_txt, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
rows = list(csv.reader(open(sass_csv)))
secs, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        secs.append(cur)
    elif r and r[0] == 'Address':
        cur['ci'] = {h: j for j, h in enumerate(r)}
    elif r and r[0].startswith('0x'):
        cur['rows'].append(r)
short = re.search(r'pixie(\d+)([a-z_]+)', kname)
short = short.group(2)[:int(short.group(1))] if short else kname
sec = [s for s in secs if ('::' + short + '(') in s['name'] or ('::' + short + '<') in s['name']][-1]
ci = sec['ci']
seq, cur_line, cur_fn = [], None, None
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
assert len(seq) == len(sec['rows']), (len(seq), len(sec['rows']))
agg = defaultdict(lambda: [0, 0, 0])
for l, r in zip(seq, sec['rows']):
    agg[l][0] += int(r[ci['Instructions Executed']] or 0)
    agg[l][1] += int(r[ci['Warp Stall Sampling (All Samples)']] or 0)
    agg[l][2] += int(r[ci['Thread Instructions Executed']] or 0)
te = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
print('warp instructions', te, 'samples', ts)
import glob
cache = {}
def src(k):
    if not k:
        return ''
    if k[0] not in cache:
        c = glob.glob('pixie_b200/csrc/cuda/' + k[0])
        cache[k[0]] = open(c[0]).read().split('\n') if c else []
    L = cache[k[0]]
    return L[k[1] - 1].strip()[:90] if 0 < k[1] <= len(L) else ''
# function regions of the .cu file the kernel lives in (top-level definitions)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * v[0] / te:5.1f}% inst {100 * v[1] / max(ts, 1):5.1f}% stall  thr/warp {v[2] / max(v[0], 1):4.1f}  {k} {src(k)}")
