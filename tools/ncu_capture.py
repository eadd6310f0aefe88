"""profiles/r02_ncu_score_kernel.json from an `ncu --set full` report of the bench command (run HERE, on the report
brought back in gpurun_out/; ncu reads reports without a GPU):

    python tools/ncu_capture.py gpurun_out/prof_bench.ncu-rep [launch index]

bench.py quotes this file in its `limiter` block -- after checking that the kernel captured is the kernel the launcher
selects today (frisk_b200_score_kernel_name); a stale capture is reported as an error there, never quoted."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, row = rows[0], rows[1], rows[2 + idx]


def num(name):
    return float(row[hdr.index(name)].replace(",", "")) if name in hdr else None


name = row[hdr.index("Kernel Name")]
m = re.search(r"(score_windows_\w+|gen_score_kernel)\s*(<[^>]*>)?", name)
short = (m.group(1) + (m.group(2) or "")) if m else name
dur_us = num("gpu__time_duration.sum")
if units[hdr.index("gpu__time_duration.sum")] == "ns":
    dur_us /= 1e3
stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(row[i])
          for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h and float(row[i]) > 0.1}
out = {
    "kernel": short,
    "kernel_full": name,
    "report": os.path.basename(rep),
    "duration_us_under_ncu": dur_us,
    "warp_instructions": num("smsp__inst_executed.sum"),
    "issue_slots_pct_of_peak": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "lsu_data_pipe_pct_of_peak": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    "shared_wavefronts": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "shared_bank_conflict_wavefronts": num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    "l2_sectors_read_by_sm": num("lts__t_sectors_srcunit_tex_op_read.sum"),
    "fp64_pipe_pct_of_peak": num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "alu_pipe_pct_of_peak": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "l2_throughput_pct": num("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    "dram_bytes": (num("dram__bytes_read.sum") or 0.0) + (num("dram__bytes_write.sum") or 0.0),
    "registers_per_thread": num("launch__registers_per_thread"),
    "ctas_per_sm_limit_shared": num("launch__occupancy_limit_shared_mem"),
    "ctas_per_sm_limit_registers": num("launch__occupancy_limit_registers"),
    "warps_stalled_per_issue": stalls,
    "note": "pct_of_peak values are this kernel's utilisation of each pipe, not a fraction of necessary work; "
            "dram_bytes is the `traffic` of bench.py's roofline block",
}
for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    if k in hdr and units[hdr.index(k)] not in ("byte", "Byte", ""):
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "KB": 1e3, "MB": 1e6, "GB": 1e9}.get(units[hdr.index(k)])
        if scale:
            out["dram_bytes"] = sum((num(x) or 0.0) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "KB": 1e3, "MB": 1e6, "GB": 1e9}.get(units[hdr.index(x)], 1.0)
                                    for x in ("dram__bytes_read.sum", "dram__bytes_write.sum") if x in hdr)
path = os.path.join(ROOT, "profiles", "r02_ncu_score_kernel.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
