"""One line per launch from an `ncu --set full` report: python tools/ncu_raw_summary.py report.ncu-rep > summary.csv
(duration, DRAM bytes / %, tensor-pipe %, issue-active %, warps-active %, executed warp instructions, registers)."""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("gpu__time_duration.sum", "duration_us"),
        ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("smsp__inst_executed.sum", "warp_instructions"), ("launch__registers_per_thread", "registers"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts")]
w = csv.writer(sys.stdout)
w.writerow([c[1] for c in cols])
units = rows[1]
for r in rows[2:]:
    out = []
    for name, short in cols:
        if name not in ix:
            out.append("")
            continue
        v = r[ix[name]]
        if short == "kernel":
            v = v.replace("void <unnamed>::", "").split("(CUtensorMap")[0][:60]
        elif short.endswith("_MB"):
            u = units[ix[name]]
            f = float(v.replace(",", ""))
            v = f"{f * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0):.1f}"
        elif short == "duration_us":
            u = units[ix[name]]
            f = float(v.replace(",", ""))
            v = f"{f * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(u, 1.0):.1f}"
        else:
            try:
                v = f"{float(v.replace(',', '')):.1f}"
            except ValueError:
                pass
        out.append(v)
    w.writerow(out)
