"""raw ncu csv (`ncu -i rep --page raw --csv`) of the GEMM-family launches of one forward -> the per-launch summary csv that
bench.py reads (profiles/*_ncu_full_summary.csv): selected columns, one row per launch, units in the second row.
usage: python scripts/ncu_gemm_summary.py raw.csv > profiles/rN_sM_gemm_kernels_ncu_full_summary.csv"""
import csv
import sys

COLS = ["ID", "Kernel Name", "launch__grid_size", "launch__cluster_dim_x", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = [hdr.index(c) if c in hdr else -1 for c in COLS]
w = csv.writer(sys.stdout)
w.writerow(COLS)
for r in rows[1:]:
    out = []
    for c, i in zip(COLS, idx):
        v = r[i] if 0 <= i < len(r) else ""
        if c == "Kernel Name":
            v = v.split("(")[0].replace("rs::", "")
        out.append(v)
    w.writerow(out)
