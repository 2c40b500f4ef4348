"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import collections
import csv
import sys


def main(path, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.OrderedDict()
    n_launch, total = 0, 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        name = row["Kernel Name"]
        ours = "<unnamed>::" in name and "at::" not in name
        key = ("OURS " if ours else "lib  ") + name.replace("void ", "").replace("<unnamed>::", "")[:90]
        c = tot.setdefault(key, [0, 0.0])
        c[0] += 1
        c[1] += v
        n_launch += 1
        total += v
    print(f"# {title}\n")
    print(f"{n_launch} launches, {total:.1f} us summed device time (ncu: cold-cache, serialised -- compare shares, not absolutes)\n")
    ours = sum(v for k, (n, v) in tot.items() if k.startswith("OURS"))
    print(f"hand-written kernels: {ours:.1f} us ({100 * ours / total:.1f} %); library kernels: {total - ours:.1f} us\n")
    print("| share | us | launches | kernel |\n|---:|---:|---:|---|")
    for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| {100 * v / total:.1f}% | {v:.1f} | {n} | `{k}` |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
