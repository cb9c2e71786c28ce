"""Warp-stall samples of an ncu report by CUDA source line (needs -lineinfo and --import-source on).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [N]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname = None
    h = None
    total = 0
    recs = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Name":
            fname = r[1].split("/")[-1]
            h = None
            continue
        if r[0] == "Line No":
            h = r
            continue
        if h is None or len(r) != len(h) or "# Samples" not in h:
            continue
        n = int(r[h.index("# Samples")] or 0)
        if n == 0:
            continue
        total += n
        stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        st = sorted(((int(r[c] or 0), h[c]) for c in stall_cols), reverse=True)[:2]
        recs.append((n, fname, r[0], r[1].strip()[:100], st))
    print("total samples", total)
    for n, f, ln, src, st in sorted(recs, reverse=True)[:top]:
        print(f"{100.0 * n / max(total, 1):5.1f}%  {f}:{ln:>4s}  {src:100s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")


if __name__ == "__main__":
    main()
