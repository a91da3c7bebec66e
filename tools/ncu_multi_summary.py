"""One block per kernel of an .ncu-rep that holds SEVERAL kernels (read here, no GPU): duration, DRAM bytes and
throughput, achieved HBM GB/s, issue / LSU utilisation.  usage: python tools/ncu_multi_summary.py rep.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
seen = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    seen.setdefault(name, []).append(r)


def to_bytes(v, u):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_s(v, u):
    v = float(v)
    return v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}.get(u, 1)


for name, rs in seen.items():
    r = rs[-1]  # last launch of the kernel (warm)
    print(f"== {name.split('(')[0]}   ({len(rs)} launches captured; last one shown)")
    for w in want:
        if w in col:
            print(f"  {w:70s} {r[col[w]]:>16s} {units[col[w]]}")
    try:
        t = to_s(r[col['gpu__time_duration.sum']], units[col['gpu__time_duration.sum']])
        b = to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']]) + \
            to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
        print(f"  {'DRAM bytes / duration':70s} {b / t / 1e9:16.1f} GB/s  ({b / 1e6:.1f} MB in {t * 1e6:.1f} us)")
    except Exception as e:
        print("  (could not derive GB/s:", e, ")")
