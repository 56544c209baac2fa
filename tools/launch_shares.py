#!/usr/bin/env python
"""Per-kernel shares of ONE denoising step from an ncu launch list (``--metrics gpu__time_duration.sum --csv``).

    python tools/launch_shares.py gpurun_out/final/launches.csv > profiles/r01_launches_step.txt

A step is delimited by two consecutive ``p_sample_update`` launches (the last kernel of every step).  ncu times are
cold-cache and serialised: compare SHARES with bench.py's live CUDA-event numbers, not absolutes.
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mn, mv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
launches = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[hi + 1:] if len(r) > mv and r[mn] == "gpu__time_duration.sum"]
unit = rows[hi + 1][h.index("Metric Unit")]
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
ends = [i for i, (n, _) in enumerate(launches) if "p_sample_update" in n]
if len(ends) < 2:
    sys.exit("need at least two p_sample_update launches in the list")
a, b = ends[-2] + 1, ends[-1] + 1
step = launches[a:b]


def short(n):
    n = re.sub(r"^void ", "", n)
    n = n.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    n = re.sub(r"\(.*", "", n)
    n = re.sub(r"<.*", "", n)
    return n


agg = collections.OrderedDict()
for n, t in step:
    k = short(n)
    c, s = agg.get(k, (0, 0.0))
    agg[k] = (c + 1, s + t * scale)
total = sum(s for _, s in agg.values())
print(f"# launch list of ONE denoising step, rows {a}..{b - 1} of {sys.argv[1]} ({len(launches)} launches captured)")
print(f"# launches per step: {len(step)}; sum of kernel time: {total:.1f} us (cold-cache, serialised under ncu)")
print(f"{'kernel':74s}{'count':>6s}{'total_us':>11s}{'share':>8s}")
for k, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:72]:74s}{c:6d}{s:11.1f}{100 * s / total:7.1f}%")
ours = sum(s for k, (c, s) in agg.items() if k.startswith("dm::"))
print(f"# our kernels (dm::*): {ours:.1f} us = {100 * ours / total:.1f}% of the step")
