"""Turn ncu outputs in gpurun_out/ into the small tracked summaries under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python tools/ncu_summary.py full gpurun_out/prof_r1b.ncu-rep profiles/r1_ncu_full.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict


def short(name):
    m = re.search(r"(\w+_kernel)", name)
    return m.group(1) if m else name[:40]


def launches(src, dst):
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"])
        per.setdefault(k, []).append(float(r["Metric Value"]) / 1e3)
    total = sum(sum(v) for v in per.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write(f"source: `{src}` — {len(rows)} launches captured inside the timed steps of `bench.py`; times are cold-cache and "
                "serialised, so compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|\n")
        for k, v in per.items():
            f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v):.1f} | {sum(v):.1f} | {100*sum(v)/total:.1f} % |\n")
        f.write(f"| all | {len(rows)} | | {total:.1f} | 100 % |\n")
    print(open(dst).read())


WANT = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__t_bytes.sum", "L1 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__shared_mem_per_block_static", "static smem/block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) blocks"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
])


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    with open(dst, "w") as f:
        f.write("# ncu --set full summary\n\n")
        f.write(f"source: `{src}` (`ncu --set full --clock-control none --import-source on`), one row block per captured launch.\n"
                "Numbers under a profiler are for diagnosis only; bench values come from `bench.py` CUDA events.\n\n")
        for r in rows[2:]:
            if len(r) < len(hdr):
                continue
            name = short(r[idx["Kernel Name"]])
            f.write(f"## `{name}` (launch id {r[idx['ID']]})\n\n| metric | value |\n|---|---|\n")
            for key, label in WANT.items():
                if key in idx:
                    f.write(f"| {label} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |\n")
            if "dram__bytes_read.sum" in idx and "dram__bytes_write.sum" in idx:
                def tobytes(v, u):
                    v = float(v.replace(",", ""))
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                tr = tobytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                    tobytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
                f.write(f"| **traffic** = DRAM read + write | {tr/1e6:.1f} MB |\n")
            st = sorted(((float(r[idx[c]] or 0), c) for c in stall_cols), reverse=True)[:6]
            if st:
                f.write("\ntop warp stall reasons (warps stalled per issue-active cycle): " +
                        ", ".join(f"{c.split('issue_stalled_')[1].split('_per_')[0]} {v:.2f}" for v, c in st) + "\n")
            f.write("\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
