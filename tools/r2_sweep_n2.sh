#!/bin/bash
# configs[4] on 2 GPUs: degree 2 / 4 / 7 x f64 / f32 at ~33 M dofs per GPU (weak scaling against
# tools/r2_single_gpu_refs.sh), and the leapfrog integrator on 2 GPUs.  Every run starts with the
# multi-GPU parity leg at its degree / precision.
mkdir -p gpurun_out/r2_scale
O=gpurun_out/r2_scale
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
X="--gpus 2 --steps 20 --warmup 3 --no-cpu --no-affine --no-extras"
port=29600
for P in 2 4 7; do for T in f64 f32; do N=$(python -c "print(round(320/$P))"); port=$((port+1))
timeout 300 $TR --master-port $port bench.py $X --degree $P --dtype $T --n-per-gpu $N > $O/n2_sweep_P${P}_$T.json 2> $O/n2_sweep_P${P}_$T.err; done; done
timeout 300 $TR --master-port 29620 bench.py $X --integrator leapfrog > $O/n2_leapfrog.json 2> $O/n2_leapfrog.err
python bench.py --steps 20 --warmup 3 --no-cpu --no-affine --no-extras --integrator leapfrog > $O/n1_leapfrog.json 2> $O/n1_leapfrog.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_scale/n2_*.json") + glob.glob("gpurun_out/r2_scale/n1_leapfrog.json")):
    try:
        d = json.load(open(f))
        p = d.get("multi_gpu_parity") or {}
        print(f.split("/")[-1], "GDoF/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 3), "dofs", d["config"]["global_dofs"],
              "parity", p.get("ok"), p.get("rel_l2_u"), "stage_frac", round(d["stage_roofline"]["frac"], 3))
    except Exception as e:
        print(f, "ERR", repr(e)[:200])
PY
