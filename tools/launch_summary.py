"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv): count, total and typical duration."""
import collections
import csv
import sys


def main(path, width=56):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(list)
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
        agg[r[ix["Kernel Name"]][:width]].append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
        s = sorted(v)
        print(f"{k:{width}s} n={len(v):5d} total {sum(v):10.1f} us ({100 * sum(v) / tot:5.1f} %)  median {s[len(s) // 2]:8.1f}  max {s[-1]:8.1f}")
    print(f"total {tot:.1f} us over {sum(len(v) for v in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
