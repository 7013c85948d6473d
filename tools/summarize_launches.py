#!/usr/bin/env python
"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list of bench.py into per-phase / per-kernel tables
(profiles/*.md).  usage: summarize_launches.py launches.csv out.md"""
import collections
import csv
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
recs = []
for row in csv.DictReader(lines):
    try:
        recs.append((row["Kernel Name"], row["Grid Size"], row["Block Size"], float(row["Metric Value"]) / 1e3))
    except Exception:
        pass


def short(n):
    n = re.sub(r"void (<unnamed>::)?", "", n)
    n = re.sub(r"\(.*", "", n)
    return n[:70]


def table(seg, title, out):
    per = collections.OrderedDict()
    for name, grid, block, us in seg:
        k = (short(name), grid, block)
        d = per.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += us
    tot = sum(v[1] for v in per.values())
    out.append(f"\n### {title}: {len(seg)} launches, {tot / 1e3:.3f} ms summed kernel time\n")
    out.append("| kernel | grid | block | launches | avg us | total us | share |")
    out.append("|---|---|---|---:|---:|---:|---:|")
    for (k, grid, block), (c, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {grid} | {block} | {c} | {us / c:.1f} | {us:.0f} | {100 * us / tot:.1f}% |")


idx_prep = [i for i, x in enumerate(recs) if "llm_prep" in x[0]]
idx_arg = [i for i, x in enumerate(recs) if "argmax_step" in x[0]]
idx_stem = [i for i, x in enumerate(recs) if "stem_im2col" in x[0]]
out = [f"# ncu launch list summary ({src})", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 1 --warmup 1 --new-tokens 4` "
       "(B=32 per GPU, bf16).  Per-launch times are cold-cache and serialised: compare SHARES, not absolutes."]
# the last vision pass that is followed by a complete prefill + two decode steps inside the capture window
s0 = p0 = None
for cand in reversed(idx_stem):
    nxt = [i for i in idx_prep if i > cand]
    if nxt and len([i for i in idx_arg if i > nxt[0]]) >= 3:
        s0, p0 = cand, nxt[0]
        break
if s0 is None:
    sys.exit("no complete vision -> prefill -> decode sequence in the launch list")
a0 = [i for i in idx_arg if i > p0][0]
table(recs[s0:p0], "vision (ResNet-50 + Q-Former), B=32", out)
table(recs[p0:a0 + 1], "prefill, B=32 x T=64", out)
a1 = [i for i in idx_arg if i > a0][:2]
table(recs[a1[0] + 1:a1[1] + 1], "one decode step, B=32 (c~66)", out)
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[-14:]))
