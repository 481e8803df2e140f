"""Selected columns of an ncu report's raw page as CSV: python tools/ncu_summary.py report.ncu-rep > profiles/summary.csv"""
import csv
import subprocess
import sys

COLS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.max",
]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    def find(c):
        for i, h in enumerate(hdr):
            if h == c or h.endswith("." + c):
                return i
        return -1
    idx = [(c, find(c)) for c in COLS]
    idx = [(c, i) for c, i in idx if i >= 0]
    w = csv.writer(sys.stdout)
    w.writerow([c for c, _ in idx])
    for r in rows[1:]:
        w.writerow([r[i] for _, i in idx])


if __name__ == "__main__":
    main()
