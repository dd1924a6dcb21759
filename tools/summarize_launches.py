"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown table of per-kernel totals.  usage: summarize_launches.py in.csv out.md "title" """
import collections, csv, re, sys
src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(l for l in open(src) if not l.startswith("==")):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    m = re.search(r"gemm_kernel<(\d+), *(\d+)>", row["Kernel Name"])
    name = f"gemm_kernel<{m.group(1)},{m.group(2)}>" if m else re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(t for _, t in agg.values())
out = [f"# {title}", "", "| kernel | launches | total ms | avg us | share |", "|---|---|---|---|---|"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {k} | {c} | {t/1e3:.2f} | {t/c:.1f} | {t/tot:.1%} |")
out += ["", f"total {tot/1e3:.2f} ms over {sum(c for c, _ in agg.values())} launches"]
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out))
