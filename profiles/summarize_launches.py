"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list into a markdown table."""
import collections
import csv
import sys

src, title = sys.argv[1], sys.argv[2]
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = row["Kernel Name"].replace("nk::<unnamed>::", "").split("(")[0][:60]
    agg.setdefault(k, []).append(float(row["Metric Value"].replace(",", "")))
print(f"# {title}\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none` — cold-cache, serialised: compare SHARES.\n")
print("Calibration micro-kernels (int_peak, red_peak), the synthetic generator and torch's L2-flush fill run outside a step's "
      "timed region; they are listed but excluded from the step share.\n")
print("| kernel | launches | total us | avg us | share of step kernels |\n|---|---:|---:|---:|---:|")
step = {k: v for k, v in agg.items() if not any(x in k for x in ("peak", "synth", "Fill", "vectorized", "hashed_idx"))}
tot = sum(sum(v) for v in step.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    sh = f"{sum(v) / tot * 100:.1f}%" if k in step else "-"
    print(f"| `{k}` | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / len(v) / 1e3:.2f} | {sh} |")
