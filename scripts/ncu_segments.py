#!/usr/bin/env python
"""Split the SASS of a profiled kernel into runs of instructions with similar execution
counts (= basic-block groups) and print each run's share of warp instructions, its average
active threads per instruction and its share of stall samples, with the CUDA source line
of the run's hottest instruction.

    python scripts/ncu_segments.py gpurun_out/prof.ncu-rep [min_share_percent]
"""
import csv, subprocess, sys


def load(rep, what):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", what, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
    rows = load(rep, "sass")
    hdr = next(r for r in rows if r and r[0] == "Address")
    ie, it, ism = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    data, base = [], None
    for r in rows:
        try:
            a = int(r[0], 16)
        except (ValueError, IndexError):
            continue
        base = a if base is None else base
        data.append((a - base, r[1].strip(), int(r[ie]), int(r[it]), int(r[ism])))
    te, tt, ts = sum(d[2] for d in data), sum(d[3] for d in data), sum(d[4] for d in data)
    print(f"{rows[0][1] if rows and len(rows[0]) > 1 else ''}")
    print(f"SASS instructions {len(data)} ({len(data) * 16 / 1024:.1f} KB)  warp instructions {te:.4g}  thread instructions {tt:.4g}  avg active threads {tt / te:.2f}")
    print(f"{'offset range':15s} {'n':>4s} {'exec/inst':>10s} {'%inst':>6s} {'thr':>5s} {'%smp':>6s}  first instruction")
    s = 0
    for i in range(1, len(data) + 1):
        if i == len(data) or abs(data[i][2] - data[s][2]) > 0.25 * max(data[s][2], 1):
            we = sum(d[2] for d in data[s:i]); wt = sum(d[3] for d in data[s:i]); sm = sum(d[4] for d in data[s:i])
            if 100 * we / te >= min_share:
                print(f"{data[s][0]:6x}-{data[i - 1][0]:6x}  {i - s:4d} {we / (i - s):10.3g} {100 * we / te:6.2f} {wt / max(we, 1):5.1f} {100 * sm / max(ts, 1):6.2f}  {data[s][1][:60]}")
            s = i


if __name__ == "__main__":
    main()
