"""Join an ncu source page (SASS, per-instruction counters) with nvdisasm line info -> executed warp instructions,
active lanes and stall samples per source line.
  ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv
  cuobjdump -xelf all libdsrt.so; nvdisasm -g -c dsrt_api.sm_100a.cubin > dis.txt
  python tools/sass_by_line.py src.csv dis.txt 'k_traceILb1ELb0' 2 [top]
(third argument: mangled-name substring of the kernel's .text section; fourth: index of the kernel table in the csv)"""
import csv, re, sys
from collections import defaultdict

src, dis, sect, kidx = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
rows = list(csv.reader(open(src)))
ks = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'rows': []}; ks.append(cur); continue
    if r and r[0] == 'Address': cur['hdr'] = r; continue
    if cur is not None and r: cur['rows'].append(r)
k = ks[kidx]; h = k['hdr']
iS, iI, iT, iM = h.index('Source'), h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
# line info per instruction, in order
lines = []; inside = False; curline = ('?', 0)
for l in open(dis):
    if l.startswith('\t.section'):
        inside = ('.text.' in l) and (sect in l); continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: curline = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4}\*/', l): lines.append((curline, l.strip()))
print('kernel', k['name'][:60], 'ncu instr', len(k['rows']), 'nvdisasm instr', len(lines))
n = min(len(lines), len(k['rows']))
agg = defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for i in range(n):
    r = k['rows'][i]; a = agg[lines[i][0]]
    for j, c in enumerate((iI, iT, iM)):
        a[j] += int(r[c]); tot[j] += int(r[c])
print('total warp-instr %d  lanes %.1f  samples %d' % (tot[0], tot[1] / max(tot[0], 1), tot[2]))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%-18s %5d  inst %5.1f%%  lanes %5.1f  samples %5.1f%%' % (key[0], key[1], 100 * a[0] / tot[0], a[1] / max(a[0], 1), 100 * a[2] / max(tot[2], 1)))
