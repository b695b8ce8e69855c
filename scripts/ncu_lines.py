#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: share of warp instructions, average active
threads per instruction (SIMT efficiency) and share of stall samples.

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, lines = None, None, {}
    kernel = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            kernel = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            ie, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or r[0] == "":
            continue   # SASS rows are already summed into their source line row
        try:
            key = (cur_file, int(r[0]))
            e, t, s = int(r[ie]), int(r[it]), int(r[isamp])
        except (ValueError, IndexError):
            continue
        x = lines.setdefault(key, [0, 0, 0, r[1].strip()[:90]])
        x[0] += e; x[1] += t; x[2] += s
    te = sum(v[0] for v in lines.values()); tt = sum(v[1] for v in lines.values()); ts = sum(v[2] for v in lines.values())
    print(f"kernel: {kernel}")
    print(f"warp instructions {te:.4g}, thread instructions {tt:.4g}, avg active threads {tt / max(te, 1):.2f}, samples {ts}")
    print(f"{'file:line':28s} {'%inst':>6s} {'thr':>6s} {'%smp':>6s}  source")
    for (f, ln), (e, t, s, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f + ':' + str(ln):28s} {100 * e / te:6.2f} {t / max(e, 1):6.2f} {100 * s / max(ts, 1):6.2f}  {src}")


if __name__ == "__main__":
    main()
