"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the
structure of one bench step.  usage: summarize_launches.py launches.csv > profiles/...txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
seq = []
for r in rows[1:]:
    v = float(r[ix["Metric Value"]])
    unit = r[ix["Metric Unit"]]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v * 1e3
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    seq.append((name, r[ix["Grid Size"]], ms))
tot = collections.defaultdict(lambda: [0, 0.0])
for n, g, ms in seq:
    tot[n][0] += 1
    tot[n][1] += ms
total = sum(v[1] for v in tot.values())
print(f"# {len(seq)} launches, {total:.3f} ms of kernel time (ncu: cold-cache, serialised; compare shares, not absolutes)")
for n, (c, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:12.3f} ms {100*ms/total:6.2f}%  x{c:<5d} {n}")
# one step of the HBM-resident digest = the run of launches from one big leaf launch to k_tree_top
big = [i for i, (n, g, ms) in enumerate(seq) if "k_sha256_lanes" in n and ms > 5.0]
if big:
    i = big[0]
    j = next(k for k in range(i, len(seq)) if ("k_tree_top" in seq[k][0] or "k_tree_root" in seq[k][0]))
    step = seq[i:j + 1]
    st = sum(ms for _, _, ms in step)
    print(f"\n# one step (launches {i}..{j}): {st:.3f} ms")
    for n, g, ms in step:
        print(f"{ms:12.4f} ms {100*ms/st:6.2f}%  grid {g:>14s}  {n}")
