"""Per-launch pipe utilisation and shared-memory bank conflicts of an `ncu --set full` report: which unit bounds a kernel
when neither DRAM nor the issue slots are saturated (this is how the ADU bound of the match.any radix pass was found).

    python profiles/summarize_pipes.py gpurun_out/prof.ncu-rep > profiles/rNN_pipes.txt
"""
import csv
import io
import subprocess
import sys

PIPES = ["adu", "cbu", "xu", "lsu", "alu", "fma", "fp64", "uniform"]


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    t_scale = {"us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "ns": 1e-3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}.get(
        units[ix["gpu__time_duration.sum"]], 1.0)

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (KeyError, ValueError):
            return float("nan")
    print(f"# {rep}: % of peak (sustained, active cycles) per pipe; L1 = l1tex data-pipe wavefronts % of peak; "
          "smem = shared-memory wavefronts (M) of which bank-conflict replays (M, loads/stores)")
    print(f"{'kernel':28s} {'us':>7s} " + " ".join(f"{p:>7s}" for p in PIPES) + f" {'L1':>6s} {'issue':>6s}   smem wavefronts / conflicts (ld / st)")
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].replace("void ", "")[:28]
        pipes = " ".join(f"{f(r, f'sm__inst_executed_pipe_{p}.avg.pct_of_peak_sustained_active'):7.1f}" for p in PIPES)
        print(f"{name:28s} {f(r, 'gpu__time_duration.sum') * t_scale:7.1f} {pipes} "
              f"{f(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):6.1f} "
              f"{f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f}   "
              f"{f(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / 1e6:6.1f} / "
              f"{f(r, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum') / 1e6:5.1f} "
              f"({f(r, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum') / 1e6:.1f} / "
              f"{f(r, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum') / 1e6:.1f})")


if __name__ == "__main__":
    main(sys.argv[1])
