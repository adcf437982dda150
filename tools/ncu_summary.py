"""Selected metrics of every kernel in an .ncu-rep (one `--set full` capture) as a small CSV for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof_step.ncu-rep profiles/rNN_full_summary.csv"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
           "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_bytes.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
           "smsp__inst_executed.sum", "sm__cycles_active.avg", "launch__grid_size", "launch__block_size",
           "launch__shared_mem_per_block_dynamic"]
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
cols = [m for m in METRICS if m in ci]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["Kernel Name"] + cols)
    w.writerow([""] + [units[ci[m]] for m in cols])
    for r in data:
        w.writerow([r[ci["Kernel Name"]]] + [r[ci[m]] for m in cols])
print("wrote", out, len(data), "kernels")
