"""DRAM traffic + time of the gate|up GEMM per raster / L2-hint setting (ncu, 4 metrics only), then the same settings inside
the power-capped denoising step (bench.py --steps 8). One process per setting."""
import os, subprocess, sys, json
# (pair, group_m in M-tiles, group_n, hints, pair stages)
# down_proj (M=65568, N=3584, K=18944): raster for large-K GEMMs: M-group pairs, N super-tile, L2 hint bits
cfgs = ["16,0,0", "4,7,1", "8,7,1", "16,7,1", "4,7,0", "8,7,5", "2,7,1", "8,5,1"]
M = "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
for kmb in cfgs:
    env = dict(os.environ, BAGEL_GEMM_BIGK=kmb)
    r = subprocess.run(["ncu", "--metrics", M, "--clock-control", "none", "-k", "regex:gemm2_bf16_kernel", "-s", "1", "-c", "1", "--csv",
                        sys.executable, "tools/gpu_ncu_targets.py", "down"], env=env, capture_output=True, text=True)
    vals = {}
    for ln in r.stdout.splitlines():
        f = [x.strip('"') for x in ln.split('","')]
        if len(f) > 3 and any(k in ln for k in ("dram__bytes", "gpu__time", "lts__t_sector")):
            vals[f[-3]] = (f[-1].strip('"'), f[-2])
    print(f"down_proj BIGK(group_pairs,group_n,hints)={kmb}: " + " | ".join(f"{k} {v[0]} {v[1]}" for k, v in vals.items()), flush=True)
for kmb in cfgs:
    env = dict(os.environ, BAGEL_GEMM_BIGK=kmb)
    r = subprocess.run([sys.executable, "bench.py", "--steps", "8", "--warmup", "3", "--no-extra", "--no-e2e", "--no-cpu-baseline"],
                       env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"in-step BIGK(group_pairs,group_n,hints)={kmb}: {d['ms_per_step']:.1f} ms/step, swiglu {d['roofline']['achieved']:.0f} TF/s, "
              f"{d['clocks']['sm_mhz']} MHz", flush=True)
    except Exception as e:
        print("bench failed", kmb, e, r.stderr[-300:], flush=True)
