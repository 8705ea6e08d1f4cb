"""Per-source-line instruction shares of one kernel in SOURCE ORDER plus headline counters.
usage: ncu_phase.py report.ncu-rep kernel-substring [min_share]"""
import csv, io, subprocess, sys
rep, want = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr = rr[0]
keys = ['gpu__time_duration.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic']
for r in rr[2:]:
    if want in r[hdr.index('Kernel Name')]:
        print({k.split('.')[0]: r[hdr.index(k)] for k in keys if k in hdr})
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
f = lambda x: float(x) if x.replace('.', '').isdigit() else 0.0
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Function Name']
acc = {}; stalls = {}
done = set()
for k, st in enumerate(starts):
    end = starts[k + 1] if k + 1 < len(starts) else len(rows)
    name = rows[st][1]
    if want not in name: continue
    fname = rows[st - 1][1] if st else ''
    sub = rows[st:end]
    hi = [i for i, r in enumerate(sub) if len(r) > 5 and r[0] == 'Line No']
    if not hi: continue
    hdr2 = sub[hi[0]]; iS = hdr2.index('# Samples'); iI = hdr2.index('Instructions Executed')
    sig = (name, fname, sum(f(r[iI]) for r in sub[hi[0] + 1:] if len(r) > iI and r[2] == '-'))
    if sig in done: continue
    done.add(sig)
    st_cols = [(i, h) for i, h in enumerate(hdr2) if h.startswith('stall_') and 'Not Issued' not in h]
    for r in sub[hi[0] + 1:]:
        if len(r) > iI and r[2] == '-':
            key = (fname.split('/')[-1], int(r[0]))
            a = acc.setdefault(key, [0, 0, r[1].strip()[:90]])
            a[0] += f(r[iI]); a[1] += f(r[iS])
            for i, h in st_cols: stalls[h] = stalls.get(h, 0) + f(r[i])
tot = sum(v[0] for v in acc.values()); ts = sum(v[1] for v in acc.values())
print('total warp instr %.4g samples %d' % (tot, ts))
print('stalls:', ', '.join('%s %.1f%%' % (h[6:], 100 * v / max(ts, 1)) for h, v in sorted(stalls.items(), key=lambda x: -x[1])[:8]))
for k, v in sorted(acc.items()):
    if v[0] / tot > thr or v[1] / max(ts, 1) > 2 * thr:
        print('%-14s %4d %5.1f%% inst %5.1f%% samp  %s' % (k[0][:14], k[1], 100 * v[0] / tot, 100 * v[1] / max(ts, 1), v[2]))
