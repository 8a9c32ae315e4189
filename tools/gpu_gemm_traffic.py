"""DRAM traffic + time of the gate|up GEMM per raster / L2-hint setting (ncu, 4 metrics only), then the same settings inside
the power-capped denoising step (bench.py --steps 8). One process per setting."""
import os, subprocess, sys, json
# (pair, group_m in M-tiles, group_n, hints, pair stages)
cfgs = [("1", "32", "0", "0", "4"), ("1", "32", "0", "8", "4"), ("1", "48", "0", "8", "4"), ("1", "64", "0", "8", "4"), ("1", "64", "0", "0", "4"),
        ("1", "32", "37", "9", "4"), ("1", "32", "0", "8", "5"), ("1", "32", "0", "12", "4"), ("1", "32", "0", "8", "4")]
M = "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
for pr, gm, gn, h, st in cfgs:
    env = dict(os.environ, BAGEL_GEMM_PAIR=pr, BAGEL_GEMM_GROUP_M=gm, BAGEL_GEMM_GROUP_N=gn, BAGEL_GEMM_HINTS=h, BAGEL_GEMM_PAIR_STAGES=st)
    r = subprocess.run(["ncu", "--metrics", M, "--clock-control", "none", "-k", "regex:gemm2?_bf16_kernel", "-s", "1", "-c", "1", "--csv",
                        sys.executable, "tools/gpu_ncu_targets.py", "gemm"], env=env, capture_output=True, text=True)
    vals = {}
    for ln in r.stdout.splitlines():
        f = [x.strip('"') for x in ln.split('","')]
        if len(f) > 3 and any(k in ln for k in ("dram__bytes", "gpu__time", "lts__t_sector")):
            vals[f[-3]] = (f[-1].strip('"'), f[-2])
    print(f"pair={pr} group_m={gm} group_n={gn} hints={h} stages={st}: " + " | ".join(f"{k} {v[0]} {v[1]}" for k, v in vals.items()), flush=True)
for pr, gm, gn, h, st in cfgs:
    env = dict(os.environ, BAGEL_GEMM_PAIR=pr, BAGEL_GEMM_GROUP_M=gm, BAGEL_GEMM_GROUP_N=gn, BAGEL_GEMM_HINTS=h, BAGEL_GEMM_PAIR_STAGES=st)
    r = subprocess.run([sys.executable, "bench.py", "--steps", "8", "--warmup", "3", "--no-extra", "--no-e2e", "--no-cpu-baseline"],
                       env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"in-step pair={pr} group_m={gm} group_n={gn} hints={h} stages={st}: {d['ms_per_step']:.1f} ms/step, swiglu {d['roofline']['achieved']:.0f} TF/s, "
              f"{d['clocks']['sm_mhz']} MHz", flush=True)
    except Exception as e:
        print("bench failed", pr, gm, gn, h, st, e, r.stderr[-300:], flush=True)
