"""Per-launch summary of an `ncu --set full` report: duration, DRAM bytes, throughputs, occupancy, top stalls.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_<what>.txt
"""
import csv
import io
import subprocess
import sys


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def g(r, k):
        return r[idx[k]] if k in idx else ""
    print(f"# {rep}: ncu --set full --clock-control none (cold caches, serialised launches)")
    for r in rows[2:]:
        t_us = float(g(r, "gpu__time_duration.sum"))
        t_s = t_us * {"us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "ns": 1e-9}.get(units[idx["gpu__time_duration.sum"]], 1e-6)
        rd, wr = float(g(r, "dram__bytes_read.sum") or 0), float(g(r, "dram__bytes_write.sum") or 0)
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(units[idx["dram__bytes_read.sum"]], 1e6)
        scale_w = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(units[idx["dram__bytes_write.sum"]], 1e6)
        traffic = rd * scale + wr * scale_w
        stalls = sorted(((float(r[idx[h]] or 0), h.split("stalled_")[1].replace("_per_issue_active.ratio", ""))
                         for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")),
                        reverse=True)[:4]
        print(f"{g(r, 'Kernel Name')[:46]:46s} {t_s * 1e6:9.1f} us  dram {traffic / 1e6:8.1f} MB ({traffic / t_s / 1e9:7.1f} GB/s, "
              f"{float(g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed') or 0):5.1f}% of ncu peak)  "
              f"L1 {float(g(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed') or 0):5.1f}%  "
              f"issue {float(g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active') or 0):5.1f}%  "
              f"warps {float(g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active') or 0):5.1f}%  "
              f"regs {g(r, 'launch__registers_per_thread')}  grid {g(r, 'launch__grid_size')}  "
              f"stalls " + ",".join(f"{n}={v:.1f}" for v, n in stalls))


if __name__ == "__main__":
    main(sys.argv[1])
