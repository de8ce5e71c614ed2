"""Summarises `ncu --set full` reports (gpurun_out/*.ncu-rep) into the per-kernel lines kept under profiles/:
duration, DRAM bytes read / written and % of peak, L2 -> L1/SM bytes, tensor-pipe and SM activity, achieved occupancy,
registers and shared memory.  Needs only the ncu CLI (no GPU).

    python tools/ncu_summary.py gpurun_out/r2e_*.ncu-rep > profiles/r2e_ncu_summary.txt
"""
import csv
import io
import os
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("lts__t_sectors_srcunit_tex.sum", "l2_sectors_from_sm"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_read"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__cycles_active.avg", "smsp_cycles"),
]


def fmt(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    if unit in ("byte", "Kbyte", "Mbyte", "Gbyte"):
        x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        return "%.1f MB" % (x / 1e6)
    if unit in ("ns", "us", "ms", "second"):
        x *= {"ns": 1e-3, "us": 1, "ms": 1e3, "second": 1e6}[unit]
        return "%.1f us" % x
    if unit == "%":
        return "%.1f%%" % x
    return ("%.0f" % x) if x == int(x) else ("%.2f" % x)


def main():
    for path in sys.argv[1:]:
        r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
        if r.returncode != 0:
            print("%s: ncu failed: %s" % (path, r.stderr[-200:]))
            continue
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if len(rows) < 3:
            print("%s: empty report" % path)
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print("== %s" % os.path.basename(path))
        for row in rows[2:]:
            name = row[col["Kernel Name"]] if "Kernel Name" in col else "?"
            parts = []
            for metric, label in WANT:
                if metric in col and row[col[metric]] != "":
                    parts.append("%s=%s" % (label, fmt(row[col[metric]], units[col[metric]])))
            print("  %s\n    %s" % (name[:110], "  ".join(parts)))


if __name__ == "__main__":
    main()
