"""Per-kernel share of ONE denoising step from an `ncu --metrics gpu__time_duration.sum --csv` launch list of
`bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline`. A step starts at the last cast_f32_bf16_kernel launch (the
latent cast that opens Bagel._velocity) and ends with cfg_apply_kernel. Usage: python tools/launch_shares.py list.csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))
names = [r["Kernel Name"] for r in rows]
start = max(i for i, n in enumerate(names) if "cast_f32_bf16_kernel" in n)
end = max(i for i, n in enumerate(names) if "cfg_apply_kernel" in n)
step = rows[start:end + 1]
agg = OrderedDict()
for r in step:
    n = r["Kernel Name"]
    m = re.search(r"(\w+)(<[^>]*>)?\(", n)
    key = (m.group(1) + (m.group(2) or "")) if m else n[:60]
    key = key.replace("(int)", "").replace("(bool)", "")
    if "gemm_bf16_kernel" in key and int(re.search(r"\((\d+)", r["Grid Size"]).group(1)) < 148:
        key += " [und-expert rows, grid<148]"
    t = float(r["Metric Value"]) / 1e6
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"# launches in the step: {len(step)}; sum of kernel time {tot:.1f} ms")
print("kernel,launches,total_ms,share_pct")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'"{k}",{n},{t:.3f},{100 * t / tot:.2f}')
