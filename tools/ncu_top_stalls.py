"""Top warp-stall sites of a kernel from an ncu report (the SASS page of `ncu --set full --import-source on`).

    python tools/ncu_top_stalls.py gpurun_out/prof_x.ncu-rep [N] [--by-line]

--by-line aggregates the samples per CUDA source line instead (inlined helpers appear under their own line numbers).

Prints the N instructions with the most stall samples, their executions and dominant stall reasons.  Reading aid for deciding
what to restructure next; the numbers come from a profiler run and are never bench values."""
import csv
import io
import subprocess
import sys


def by_line(rep, n):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    k = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    si = rows[k].index("# Samples")
    cur, agg, src = None, {}, {}
    for r in rows[k + 1:]:
        if len(r) <= si:
            continue
        if r[0] not in ("", "-"):
            try:
                cur = int(r[0])
                src[cur] = r[1]
            except ValueError:
                pass
        elif cur is not None and r[0] == "":
            try:
                agg[cur] = agg.get(cur, 0) + int(r[si] or 0)
            except ValueError:
                pass
    tot = sum(agg.values())
    print(f"total samples {tot}")
    for ln, v in sorted(agg.items(), key=lambda kv: -kv[1])[:n]:
        print(f"{v:6d} {100 * v / tot:5.1f}%  L{ln}: {src[ln].strip()[:110]}")


def main():
    rep = sys.argv[1]
    args = [a for a in sys.argv[2:] if not a.startswith("--")]
    n = int(args[0]) if args else 15
    if "--by-line" in sys.argv:
        return by_line(rep, n)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[k - 1][1] if k > 0 and len(rows[k - 1]) > 1 else "")
    hdr, data = rows[k], [r for r in rows[k + 1:] if len(r) > 5]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[idx["# Samples"]]) for r in data)
    print(f"total samples {tot}")
    agg = {h: sum(int(r[idx[h]]) for r in data) for h in stalls}
    print("by reason: " + ", ".join(f"{h[6:]}={100 * v / tot:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:n]:
        s = int(r[idx["# Samples"]])
        top = sorted(((int(r[idx[h]]), h[6:]) for h in stalls if int(r[idx[h]]) > 0), reverse=True)[:2]
        print(f"{s:6d} {100 * s / tot:5.1f}% {int(r[idx['Instructions Executed']]):9d}  {r[1].strip()[:72]:72s} {', '.join(f'{h}={v}' for v, h in top)}")


if __name__ == "__main__":
    main()
