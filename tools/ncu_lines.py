"""Per-source-line instruction / stall-sample shares per kernel from an .ncu-rep
(needs -lineinfo and --import-source on).  usage: ncu_lines.py report.ncu-rep [kernel-substring] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
f = lambda x: float(x) if x.replace('.', '').isdigit() else 0.0
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Function Name']
seen = set()
for k, st in enumerate(starts):
    end = starts[k + 1] if k + 1 < len(starts) else len(rows)
    name = rows[st][1]
    sub = rows[st:end]
    hi = [i for i, r in enumerate(sub) if len(r) > 5 and r[0] == 'Line No']
    if not hi or want not in name:
        continue
    hdr = sub[hi[0]]; iS = hdr.index('# Samples'); iI = hdr.index('Instructions Executed')
    lines = [r for r in sub[hi[0] + 1:] if len(r) > iI and r[2] == '-']
    ti = sum(f(r[iI]) for r in lines); ts = sum(f(r[iS]) for r in lines)
    if ti < 1e5 or (name, round(ti)) in seen:
        continue
    seen.add((name, round(ti)))
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    agg = sorted(((sum(f(r[i]) for r in lines), h) for i, h in stall), reverse=True)[:7]
    print('== %s\n   file %s: warp instructions %.4g, samples %d' % (name[:90], rows[st - 1][1] if st else '', ti, ts))
    print('   stall samples:', ', '.join('%s %.1f%%' % (h[6:], 100 * v / max(ts, 1)) for v, h in agg))
    for r in sorted(lines, key=lambda r: -f(r[iI]))[:top]:
        print('%5.1f%% inst %5.1f%% samp  L%-4s %s' % (100 * f(r[iI]) / ti, 100 * f(r[iS]) / max(ts, 1), r[0], r[1].strip()[:105]))
