"""Turn an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X) into a per-kernel share table:
python tools/summarize_launches.py profiles/launches_r01_v3.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
H = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter()
cnt = collections.Counter()
for r in rows:
    if r[H["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[H["Kernel Name"]])
    name = re.sub(r"^void ", "", name)
    ns = float(r[H["Metric Value"]].replace(",", ""))
    if r[H["Metric Unit"]] in ("us", "usecond"):
        ns *= 1e3
    elif r[H["Metric Unit"]] in ("ms", "msecond"):
        ns *= 1e6
    tot[name] += ns
    cnt[name] += 1
setup = {k for k in tot if "gen_points" in k or "precompute" in k or "imad_" in k}
prover = sum(v for k, v in tot.items() if k not in setup)
print("| kernel | launches | total ms | share of prover kernel time |\n|---|---|---|---|")
for k, v in tot.most_common():
    if k in setup or v / prover < 0.002:
        continue
    print("| `%s` | %d | %.1f | %.1f%% |" % (k, cnt[k], v / 1e6, 100 * v / prover))
print("\nprover kernels total %.0f ms; outside every timed region: %s" % (
    prover / 1e6, ", ".join("`%s` %.0f ms" % (k, tot[k] / 1e6) for k in sorted(setup))))
