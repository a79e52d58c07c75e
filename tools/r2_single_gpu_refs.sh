#!/bin/bash
# 1-GPU denominators of tools/r2_scale_run.sh and tools/r2_sweep_n2.sh: the same per-GPU sizes at N = 1.
mkdir -p gpurun_out/r2_scale
O=gpurun_out/r2_scale
X="--steps 20 --warmup 3 --no-cpu --no-affine --no-extras"
python bench.py $X --n-per-gpu 125 > $O/n1_box125_f64.json 2> $O/n1_box125_f64.err
python bench.py $X --n-per-gpu 125 --dtype f32 > $O/n1_box125_f32.json 2> $O/n1_box125_f32.err
python bench.py $X --workload nonlinear_bowl > $O/n1_bowl99.json 2> $O/n1_bowl99.err
python bench.py $X --workload linear_piston --n-per-gpu 58 > $O/n1_piston58.json 2> $O/n1_piston58.err
python bench.py $X --workload linear_piston --n-per-gpu 74 > $O/n1_piston74.json 2> $O/n1_piston74.err
for P in 2 4 7; do for T in f64 f32; do N=$(python -c "print(round(320/$P))")
python bench.py $X --degree $P --dtype $T --n-per-gpu $N > $O/n1_sweep_P${P}_$T.json 2> $O/n1_sweep_P${P}_$T.err; done; done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_scale/n1_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "GDoF/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 3), "dofs", d["config"]["global_dofs"],
              "stage_frac", round(d["stage_roofline"]["frac"], 3), "kernel_frac", round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "ERR", repr(e)[:200])
PY
