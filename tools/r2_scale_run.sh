#!/bin/bash
# Round-2 multi-GPU legs on ONE 8-GPU box (gpurun --gpus 8 -- bash tools/r2_scale_run.sh):
#   weak scaling of the headline (80^3 cells per GPU) at N = 1, 4, 8 on the same box;
#   north-star size: linear box, degree 4, 125^3 cells per GPU -> 1.0e9 dofs on 8 GPUs, f64 and f32;
#   configs[3]: Westervelt bowl, 99^3 cells per GPU -> 4.99e8 dofs on 8 GPUs;
#   configs[2]: piston, degree 5, ~1e8 dofs on 4 GPUs (58^3 per GPU) and 2 GPUs (74^3 per GPU), side by side.
# Every N > 1 run starts with bench.py's multi-GPU parity leg (a small box on all ranks vs rank 0 alone)
# and aborts if it fails.  1-GPU denominators of the big sizes: tools/single_gpu_refs.sh.
mkdir -p gpurun_out/r2_scale
O=gpurun_out/r2_scale
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
X="--steps 20 --warmup 3 --no-cpu --no-affine --no-extras --watchdog 400"
python bench.py $X > $O/n1_box80.json 2> $O/n1_box80.err
timeout 300 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 $X > $O/n8_box80.json 2> $O/n8_box80.err
timeout 300 $TR --nproc-per-node 4 --master-port 29512 bench.py --gpus 4 $X > $O/n4_box80.json 2> $O/n4_box80.err
timeout 600 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 $X --n-per-gpu 125 > $O/n8_box125_f64.json 2> $O/n8_box125_f64.err
timeout 600 $TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 $X --n-per-gpu 125 --dtype f32 > $O/n8_box125_f32.json 2> $O/n8_box125_f32.err
timeout 400 $TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 $X --workload nonlinear_bowl > $O/n8_bowl99.json 2> $O/n8_bowl99.err
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 400 $TR --nproc-per-node 4 --master-port 29516 bench.py --gpus 4 $X --workload linear_piston --n-per-gpu 58 > $O/n4_piston58.json 2> $O/n4_piston58.err &
CUDA_VISIBLE_DEVICES=4,5 timeout 400 $TR --nproc-per-node 2 --master-port 29517 bench.py --gpus 2 $X --workload linear_piston --n-per-gpu 74 > $O/n2_piston74.json 2> $O/n2_piston74.err &
wait
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_scale/*.json")):
    try:
        d = json.load(open(f))
        p = d.get("multi_gpu_parity") or {}
        print(f.split("/")[-1], "GDoF/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 3), "dofs", d["config"]["global_dofs"],
              "parity", p.get("ok"), p.get("rel_l2_u"), "stage_frac", round(d["stage_roofline"]["frac"], 3), "clk", d["clocks"].get("per_rank_sm_mhz") or d["clocks"].get("sm_mhz"))
    except Exception as e:
        print(f, "ERR", repr(e)[:200])
PY
tail -n 2 gpurun_out/r2_scale/*.err | cut -c1-300
