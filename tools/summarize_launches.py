"""Summarise an ncu launch list (``--metrics gpu__time_duration.sum --csv``) of bench.py: shares of ONE eager train
step (from one embedding gather launch to the next).  usage: python tools/summarize_launches.py launches.csv"""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Metric Unit"]))
unit = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
starts = [i for i, r in enumerate(rows) if "gather_onehot_kernel<4" in r[0] or "gather_onehot_kernel<(int)4" in r[0]]
pairs = [(a, b) for a, b in zip(starts, starts[1:]) if b - a >= 20]  # a train step, not the kernel-timing loops after it
if not pairs:
    sys.exit("fewer than two steps in the list")
a, b = min(pairs, key=lambda ab: ab[1] - ab[0])  # (the last one also spans the set-up of the kernel timings)
step = rows[a:b]
tot = sum(t * unit[u] for _, t, u in step)
agg = collections.OrderedDict()
for k, t, u in step:
    name = k.split("(")[0][:110]
    e = agg.setdefault(name, [0.0, 0])
    e[0] += t * unit[u]
    e[1] += 1
mine = sum(v[0] for k, v in agg.items() if "ptrec::" in k)
nmine = sum(v[1] for k, v in agg.items() if "ptrec::" in k)
print(f"kernel time of the step: {tot:.1f} us over {len(step)} launches; this library: {nmine} launches = {100 * mine / tot:.1f} % of the kernel time")
print(" share%        us    n  kernel")
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{100 * t / tot:7.2f} {t:9.1f} {n:4d}  {k}")
