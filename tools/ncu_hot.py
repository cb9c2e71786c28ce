"""Top stall locations of an ncu report's source page (SASS level, with the dominant stall reason).

    python tools/ncu_hot.py gpurun_out/prof.ncu-rep [N]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # one section per profiled launch: take the one asked for (third argument, default the first)
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    hi = heads[which]
    rows = rows[:heads[which + 1]] if which + 1 < len(heads) else rows
    h = rows[hi]
    si = h.index("# Samples")
    src = h.index("Source")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    body = [r for r in rows[hi + 1:] if len(r) == len(h)]
    total = sum(int(r[si] or 0) for r in body)
    print("total samples", total)
    ranked = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:top]
    for i in sorted(ranked):
        r = body[i]
        n = int(r[si] or 0)
        st = sorted(((int(r[c] or 0), h[c]) for c in stall_cols), reverse=True)[:2]
        print(f"{i:5d} {100.0 * n / max(total, 1):5.1f}%  {r[src][:90]:90s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")


if __name__ == "__main__":
    main()
