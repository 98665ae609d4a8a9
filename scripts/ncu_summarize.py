"""Summarise an .ncu-rep of ONE kernel launch into text: key raw metrics + the instructions with the most stall samples.
usage: python scripts/ncu_summarize.py <file.ncu-rep> > profiles/<name>.txt   (needs ncu on PATH; no GPU)"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
print(f"# {rep}")
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith("smsp__average_warps_issue_stalled_"):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"\n# warp-state samples: {tot} over {len(data)} SASS instructions; instructions with >= 0.6 % of the samples")
for i, r in enumerate(data):
    n = int(r[ix["# Samples"]] or 0)
    if n >= 0.006 * tot:
        st = sorted(((int(r[ix[s]] or 0), s.replace("stall_", "")) for s in stalls), reverse=True)[:2]
        print(f"{i:5d} {100 * n / tot:5.1f}%  {r[ix['Source']][:70]:70s} {st}")
print("\n# samples per 200-instruction window (code order)")
for a in range(0, len(data), 200):
    n = sum(int(r[ix["# Samples"]] or 0) for r in data[a:a + 200])
    ex = sum(int(r[ix["Instructions Executed"]] or 0) for r in data[a:a + 200])
    print(f"[{a:5d},{a + 200:5d}) {100 * n / tot:5.1f}%  warp-instructions executed {ex}")
