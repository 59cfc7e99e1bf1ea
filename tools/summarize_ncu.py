"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel family."""
import csv
import sys
from collections import defaultdict

path, out = sys.argv[1], sys.argv[2]
rows = []
with open(path) as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    unit = r[iu]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    rows.append((r[ik], us))
fam = defaultdict(lambda: [0, 0.0])
for name, us in rows:
    key = name.split("(")[0]
    if "gemm_f64_kernel" in key:
        key = key[key.index("gemm_f64_kernel"):]
    fam[key][0] += 1
    fam[key][1] += us
tot = sum(v[1] for v in fam.values())
with open(out, "w") as fh:
    fh.write(f"# ncu launch list summary ({path})\n\n")
    fh.write("Per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes.\n\n")
    fh.write(f"{len(rows)} launches, {tot / 1e3:.2f} ms total\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
    for k, (n, us) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"| `{k}` | {n} | {us:.1f} | {100 * us / tot:.1f}% |\n")
print(open(out).read())
