"""Per-source-line instruction / stall-sample shares from an .ncu-rep (needs -lineinfo and --import-source on)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == 'Line No'][0]
hdr = rows[hi]; iS = hdr.index('# Samples'); iI = hdr.index('Instructions Executed')
f = lambda x: float(x) if x.replace('.', '').isdigit() else 0.0
lines = [r for r in rows[hi + 1:] if len(r) > iI and r[2] == '-']
ti = sum(f(r[iI]) for r in lines); ts = sum(f(r[iS]) for r in lines)
print('total warp instructions %.4g, samples %d' % (ti, ts))
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
agg = sorted(((sum(f(r[i]) for r in lines), h) for i, h in stall), reverse=True)[:8]
print('stall samples:', ', '.join('%s %.1f%%' % (h[6:], 100 * v / ts) for v, h in agg))
for r in sorted(lines, key=lambda r: -f(r[iI]))[:top]:
    print('%5.1f%% inst %5.1f%% samp  L%-4s %s' % (100 * f(r[iI]) / ti, 100 * f(r[iS]) / ts, r[0], r[1].strip()[:105]))
