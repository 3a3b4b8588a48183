"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-step kernel table + shares.

    python profiles/summarize_launches.py gpurun_out/launches.csv [skip_steps]
"""
import csv
import re
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
kn, mv, gs, bs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
launches = [(r[kn], float(r[mv].replace(",", "")) / 1000.0, r[gs], r[bs]) for r in rows[h + 1:] if len(r) > mv and re.match(r"(void )?(dd|tc)::", r[kn])]
# a step starts at every synth kernel
starts = [i for i, l in enumerate(launches) if "synth_" in l[0] and "finalize" not in l[0]]
steps = [launches[a:b] for a, b in zip(starts, starts[1:] + [len(launches)])]
last = steps[-1]
tot = sum(l[1] for l in last)
print(f"# {len(steps)} steps, {len(last)} launches in the last step, sum of kernel durations {tot:.1f} us (cold-cache, serialised)")
print("| # | kernel | grid | block | us | share |")
print("|---|---|---|---|---|---|")
for i, l in enumerate(last):
    name = l[0].replace("void ", "").split("(")[0]
    print(f"| {i} | `{name}` | {l[2]} | {l[3]} | {l[1]:.1f} | {100 * l[1] / tot:.1f}% |")
