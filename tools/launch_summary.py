"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share."""
import collections, csv, sys

def main(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi or not r[vi].replace(",", "").replace(".", "").isdigit():
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
        k = r[ki].split("(")[0][-70:]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += float(r[vi].replace(",", "")) * scale
    tot = sum(v[1] for v in agg.values())
    print(f"{'total us':>12} {'n':>5} {'share':>6}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:12.1f} {v[0]:5d} {100 * v[1] / tot:5.1f}%  {k}")

if __name__ == "__main__":
    main(sys.argv[1])
