"""Per-source-line warp-stall samples from `ncu --page source --csv --print-source cuda,sass` output.
usage: python tools/ncu_lines.py cs.csv [first_line last_line]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 10**9)
cur, hdr, out = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 3 and r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2 or r[0] == '':
        continue
    try:
        samples, inst = float(r[6]), float(r[7])
    except ValueError:
        continue
    stalls = {h: float(v or 0) for h, v in zip(hdr, r) if h.startswith('stall_') and 'Not Issued' not in h}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    out.append((samples, inst, cur, int(r[0]), r[1][:90], top))
tot = sum(o[0] for o in out)
sel = [o for o in out if lo <= o[3] <= hi and (len(sys.argv) <= 3 or o[2] == 'decoder_bf16.cuh')]
print('total samples', tot, ' selected', sum(o[0] for o in sel), ' warp-instr', sum(o[1] for o in sel))
for o in sorted(sel, key=lambda o: -o[0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    print('%7.0f %5.1f%% %10.0f %s:%d  %s   %s' % (o[0], 100 * o[0] / tot, o[1], o[2], o[3], o[4].strip()[:80], [(k[6:], int(v)) for k, v in o[5]]))
