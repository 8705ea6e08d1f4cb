"""Write profiles/r2_traffic.json from one `ncu --set full` capture of the fused traversal: DRAM bytes of that launch and
the sha1 of the traverse.cu it was taken on (bench.py reports roofline.traffic only while the source still has that hash).
usage: python tools/traffic_from_ncu.py capture.ncu-rep out.json"""
import csv, hashlib, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, dst = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, r = rows[0], rows[1], rows[2]
idx = {h: i for i, h in enumerate(hdr)}
def tobytes(key):
    v = float(r[idx[key]].replace(",", ""))
    return int(round(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[idx[key]]]))
rd, wr = tobytes("dram__bytes_read.sum"), tobytes("dram__bytes_write.sum")
src = open(os.path.join(ROOT, "naivedynamics.jl_b200", "csrc", "traverse.cu"), "rb").read()
json.dump({"traverse_kernel": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "traverse_cu_sha1": hashlib.sha1(src).hexdigest(), "workload": "c3 (1M LJ+Coulomb), melted 600 steps",
           "kernel": r[idx["Kernel Name"]][:80], "duration_under_ncu_us": r[idx["gpu__time_duration.sum"]] + " " + units[idx["gpu__time_duration.sum"]],
           "note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (ncu --set full --clock-control none, launch 611 of "
                   "tools/stage_bench.py c3 after 600 melt steps); below the algorithmic bytes because positions and the force "
                   "reductions stay in the 126 MB L2 across kernels",
           "command": "NB200_NO_GRAPH=1 NB200_PRESTEPS=600 ncu --set full --clock-control none --import-source on -k regex:traverse_kernel "
                      "-s 610 -c 1 -o gpurun_out/r2_traverse_final python tools/stage_bench.py c3 30"}, open(dst, "w"), indent=1)
print(open(dst).read())
