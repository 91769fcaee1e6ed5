"""Stall samples per CUDA source line of one kernel launch from an .ncu-rep (needs -lineinfo + --import-source on).

    python profiles/hot_lines.py gpurun_out/prof.ncu-rep <kernel-regex> [launch-skip] [top-n]
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
    rows, fname, i = [], None, 0
    while i < len(lines):
        l = lines[i]
        if l.startswith('"File Path"'):
            fname = next(csv.reader([l]))[1].split("/")[-1]
            i += 1
            continue
        if l.startswith('"Line No"'):
            hdr = next(csv.reader([l]))
            j = i + 1
            while j < len(lines) and not lines[j].startswith('"File Path"'):
                j += 1
            for r in csv.reader(io.StringIO("\n".join(lines[i + 1:j]))):
                if len(r) >= 7 and r[0].isdigit():
                    d = dict(zip(range(len(r)), r))
                    rows.append((fname, int(r[0]), r[1], r, hdr))
            i = j
            continue
        i += 1
    agg = {}
    for fname, ln, src, r, hdr in rows:
        h = {k: idx for idx, k in enumerate(hdr)}
        smp = r[h["# Samples"]]
        if not smp.isdigit():
            continue
        key = (fname, ln)
        a = agg.setdefault(key, {"src": src, "n": 0, "st": {}})
        a["n"] += int(smp)
        for k, idx in h.items():
            if k.startswith("stall_") and "Not Issued" not in k and idx < len(r) and r[idx].isdigit():
                a["st"][k] = a["st"].get(k, 0) + int(r[idx])
    tot = sum(a["n"] for a in agg.values())
    print(f"total samples {tot}")
    for (fname, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:top]:
        st = sorted(a["st"].items(), key=lambda kv: -kv[1])[:3]
        sts = " ".join(f"{k[6:]}={v}" for k, v in st if v)
        print(f"{a['n'] / max(tot, 1):6.1%} {fname}:{ln:<5d} {a['src'].strip()[:90]:90s} {sts}")


if __name__ == "__main__":
    main()
