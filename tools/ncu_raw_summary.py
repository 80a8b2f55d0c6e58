#!/usr/bin/env python
"""Key figures of one kernel from an `ncu --page raw --csv` export.
usage: tools/ncu_raw_summary.py raw.csv [frames_per_launch]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    frames = float(sys.argv[2]) if len(sys.argv) > 2 else 0
    hdr, vals = rows[0], rows[-1]
    d = dict(zip(hdr, vals))
    print(d.get("Kernel Name", ""))
    for k in KEYS:
        if k in d:
            extra = ""
            if frames and k.endswith(".sum") and "bytes" not in k and "duration" not in k:
                extra = f"   ({float(d[k]) / frames:.1f} per frame)"
            print(f"  {k} = {d[k]}{extra}")
    st = [(float(d[h]), h) for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and d[h]]
    print("  stalls (warps per issue):", ", ".join(f"{h.split('issue_stalled_')[1].split('_per_')[0]}={v:.2f}" for v, h in sorted(st, reverse=True)[:9]))


if __name__ == "__main__":
    main()
