"""Extracts the judged metrics from an `ncu --set full` report into CSV (one row per captured launch):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_x_summary.csv"""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    out = csv.writer(sys.stdout)
    out.writerow([f"{w} [{units[i]}]" if units[i] else w for w, i in idx])
    for d in data:
        out.writerow([d[i][:120] for _, i in idx])


if __name__ == "__main__":
    main(sys.argv[1])
