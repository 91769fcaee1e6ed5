"""Executed warp-instructions per CUDA source line of one kernel launch from an .ncu-rep (needs -lineinfo + --import-source on).

    python profiles/inst_lines.py gpurun_out/prof.ncu-rep <kernel-regex> [launch-skip] [top-n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name",
                          f"regex:{rx}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    agg, fname, i = {}, None, 0
    while i < len(lines):
        l = lines[i]
        if l.startswith('"File Path"'):
            fname = next(csv.reader([l]))[1].split("/")[-1]
        elif l.startswith('"Line No"'):
            hdr = next(csv.reader([l]))
            h = {k: idx for idx, k in enumerate(hdr)}
            j = i + 1
            while j < len(lines) and not lines[j].startswith('"File Path"'):
                j += 1
            for r in csv.reader(io.StringIO("\n".join(lines[i + 1:j]))):
                if len(r) > h["Instructions Executed"] and r[0].isdigit() and r[h["Instructions Executed"]].isdigit():
                    a = agg.setdefault((fname, int(r[0])), [r[1], 0, 0])
                    a[1] += int(r[h["Instructions Executed"]])
                    a[2] += int(r[h["Thread Instructions Executed"]])
            i = j
            continue
        i += 1
    tot = sum(a[1] for a in agg.values())
    print(f"total warp instructions {tot}")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{a[1] / max(tot, 1):6.1%} {a[1]:11d} thr/inst {a[2] / max(a[1], 1):5.1f}  {f}:{ln:<5d} {a[0].strip()[:100]}")


if __name__ == "__main__":
    main()
