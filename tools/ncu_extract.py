"""Compact per-launch table of the metrics the roofline uses from an `ncu --set full` report:
    python tools/ncu_extract.py gpurun_out/r02_full_512.ncu-rep > profiles/r02_ncu_full_512.csv"""
import csv
import io
import re
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read_bytes",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
         "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    cols = {}
    for i, h in enumerate(head):
        for k, short in WANT.items():
            if h == k or h.endswith("." + k) or h.endswith(k):
                cols.setdefault(short, i)
    name_i, grid_i = head.index("Kernel Name"), head.index("Grid Size")
    out = csv.writer(sys.stdout)
    keys = list(cols)
    out.writerow(["id", "kernel", "grid"] + [k + ("_us" if k == "duration" else ("_bytes" if k.startswith("dram_") else "")) for k in keys])
    for r in rows[2:]:
        if len(r) <= name_i:
            continue
        m = re.search(r"(\w+_kernel(?:<[^>]*>)?)", r[name_i])
        vals = []
        for k in keys:
            v = r[cols[k]].replace(",", "")
            try:
                f = float(v) * SCALE.get(units[cols[k]], 1.0)
                vals.append(f"{f:.6g}")
            except ValueError:
                vals.append(v)
        out.writerow([r[0], m.group(1) if m else r[name_i][:50], r[grid_i]] + vals)


if __name__ == "__main__":
    main(sys.argv[1])
