"""Per-source-line summary of one kernel from an ncu report captured with --import-source on:
   python scripts/ncu_lines.py report.ncu-rep [min_pct]   (stall samples, instructions, shared-memory wavefronts per CUDA line)"""
import csv, subprocess, sys, io
rep = sys.argv[1]
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Line No')
hdr = rows[hi]
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
def g(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0
data = [r for r in rows[hi + 1:] if len(r) > 10 and r[2] == '-']
tot = sum(g(r, '# Samples') for r in data) or 1
ti = sum(g(r, 'Instructions Executed') for r in data) or 1
tw = sum(g(r, 'L1 Wavefronts Shared') for r in data)
te = sum(g(r, 'L1 Wavefronts Shared Excessive') for r in data)
print('samples %d  warp instructions %d  shared wavefronts %d (excessive %d)' % (tot, ti, tw, te))
for r in data:
    sp, ip = 100 * g(r, '# Samples') / tot, 100 * g(r, 'Instructions Executed') / ti
    if sp >= minp or ip >= minp or g(r, 'L1 Wavefronts Shared Excessive') > 0:
        print('%-4s smp %5.1f%% inst %5.1f%% wf %9d exc %9d | %s' % (r[0], sp, ip, g(r, 'L1 Wavefronts Shared'),
              g(r, 'L1 Wavefronts Shared Excessive'), r[1].strip()[:100]))
