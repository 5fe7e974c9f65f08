"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import collections, csv, re, sys

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.defaultdict(lambda: [0, 0.0])
n = 0
for x in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", x["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    val = float(x["Metric Value"].replace(",", ""))
    unit = x["Metric Unit"]
    val = val / 1e3 if unit == "ns" else (val * 1e3 if unit == "ms" else val)
    tot[name][0] += 1
    tot[name][1] += val
    n += 1
allt = sum(v[1] for v in tot.values())
print(f"# {n} launches, {allt / 1e3:.2f} ms of kernel time in the capture ({steps:g} steps -> {allt / 1e3 / steps:.2f} ms/step, cold-cache serialised)")
print("| share | ms/step | launches/step | kernel |")
print("|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| {v[1] / allt * 100:5.1f}% | {v[1] / 1e3 / steps:7.3f} | {v[0] / steps:6.1f} | {k[:90]} |")
