#!/usr/bin/env python
"""Turn an .ncu-rep (brought back from the GPU box in gpurun_out/) into the small text summary that
is committed under profiles/.   python profiles/summarize.py gpurun_out/X.ncu-rep [units_per_launch]

`units_per_launch` (optional) = warp-steps per launch, to print executed instructions per warp per
env-step.  Uses `ncu -i ... --page raw --csv` and `--page source --csv` (no GPU needed)."""
import csv
import io
import subprocess
import sys
from collections import Counter

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, unit = rows[0], rows[1]
    print(f"# {rep}")
    for r in rows[2:]:
        print(f"\n## {r[hdr.index('Kernel Name')]}")
        for m in METRICS:
            if m in hdr:
                print(f"{m:72s} {r[hdr.index(m)]:>16s} {unit[hdr.index(m)]}")
    src = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    for b in blocks[:1]:
        h = b["hdr"]
        i_s, i_e, i_n = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        mix, samples, total = Counter(), Counter(), 0
        for r in b["rows"]:
            toks = r[i_s].split()
            op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            mix[op] += int(r[i_e]); samples[op] += int(r[i_n]); total += int(r[i_e])
        print(f"\n## executed SASS mix of {b['name']} (warp-level instructions)")
        print(f"total {total}" + (f" = {total / units:.1f} per warp per env-step" if units else ""))
        for op, n in mix.most_common(16):
            print(f"  {op:10s} {100.0 * n / total:5.1f} %" + (f"  {n / units:6.1f}/step" if units else "") + f"   stall samples {samples[op]}")


if __name__ == "__main__":
    main()
