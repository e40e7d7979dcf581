#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel: joins the SASS page of an ncu report (executed instructions per SASS
instruction) with nvdisasm's line table of the same cubin (built with -lineinfo).
  python tools/ncu_by_line.py <report.ncu-rep> <object.o> <kernel-section-substring> [top]
Runs HERE (needs ncu, cuobjdump, nvdisasm; no GPU)."""
import csv, os, re, subprocess, sys, tempfile, collections

rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], cwd=tmp, check=True, capture_output=True, text=True).stdout.splitlines()
# instruction -> innermost and outermost (file, line)
lines, inside, cur, chain = [], False, None, []
for ln in dis:
    if ln.startswith('.text.'):
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        chain = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', ln):
        outer = (os.path.basename(chain[-1][0]), int(chain[-1][1])) if chain else cur
        lines.append((cur, outer, ln.split('*/', 1)[1].strip()))
csvtxt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
h = rows[hi]
ii, si = h.index('Instructions Executed'), h.index('Source')
sass = [(r[si].strip(), int(r[ii])) for r in rows[hi + 1:] if len(r) > ii and r[ii].isdigit()]
print(f'# {len(lines)} instructions in the cubin section, {len(sass)} in the report')
n = min(len(lines), len(sass))
tot = sum(c for _, c in sass)
inner, outerc = collections.Counter(), collections.Counter()
for (cur, outer, _), (_, c) in zip(lines[:n], sass[:n]):
    inner[cur] += c
    outerc[outer] += c
print(f'# total executed warp-instructions {tot}')
for title, cnt in (('innermost source line', inner), ('outermost (kernel body) line', outerc)):
    print(f'## by {title}')
    for k, c in cnt.most_common(top):
        print(f'{c:>12} {100 * c / tot:5.1f}%  {k[0]}:{k[1]}' if k else f'{c:>12} {100 * c / tot:5.1f}%  ?')
