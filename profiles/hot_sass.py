"""Top stall-sample SASS lines of one kernel launch from an .ncu-rep (source page).

    python profiles/hot_sass.py gpurun_out/prof.ncu-rep <kernel-regex> [launch-skip] [top-n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}",
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    print(lines[start - 1][:160])
    rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))) if (r["# Samples"] or "0").isdigit()]
    tot = sum(int(r["# Samples"] or 0) for r in rows)
    print(f"total samples {tot}, {len(rows)} SASS instructions")
    ranked = sorted(enumerate(rows), key=lambda ir: -int(ir[1]["# Samples"] or 0))[:top]
    for i, r in sorted(ranked):
        print(f"{i:5d} {int(r['# Samples']):7d} {int(r['# Samples']) / max(tot, 1):6.1%}  {r['Source'].strip()[:110]}")


if __name__ == "__main__":
    main()
