"""Key metrics of an `ncu --set full` report as text (what profiles/*_summary.txt hold) and, with --traffic, the DRAM bytes per launch
as JSON for profiles/traffic.json.   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--traffic]"""
import csv
import io
import json
import subprocess
import sys

WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = WANT + [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
ki = hdr.index("Kernel Name")
seen, traffic = {}, {}
for r in rows[2:]:
    name = r[ki]
    seen[name] = seen.get(name, 0) + 1
    if seen[name] > 1 and "--all" not in sys.argv:
        continue                      # one launch per kernel (the first captured one)
    if "--traffic" in sys.argv:
        def val(m):
            v, u = float(r[hdr.index(m)]), units[hdr.index(m)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        traffic[name.split("(")[0].split("::")[-1].split("<")[0].replace("void ", "")] = {"dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum")}
        continue
    print("----", name[:100])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("   %-95s %s %s" % (w, r[i], units[i]))
if "--traffic" in sys.argv:
    print(json.dumps(traffic, indent=1))
