#!/bin/bash
# 1-GPU denominators of the weak-scaling runs (tools/scale_run.sh, tools/scale_run_auto.sh): same per-GPU size, N=1.
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --no-cpu --no-affine --n-per-gpu 125 -v > gpurun_out/n1_linear_box_125.json 2> gpurun_out/n1_linear_box_125.err
python bench.py --steps 20 --no-cpu --no-affine --workload nonlinear_bowl -v > gpurun_out/n1_bowl_99.json 2> gpurun_out/n1_bowl_99.err
python bench.py --steps 20 --no-cpu --no-affine --workload linear_piston --n-per-gpu 58 -v > gpurun_out/n1_piston_58.json 2> gpurun_out/n1_piston_58.err
python bench.py --steps 20 --no-cpu --no-affine --workload linear_piston --n-per-gpu 74 -v > gpurun_out/n1_piston_74.json 2> gpurun_out/n1_piston_74.err
python bench.py --steps 20 --no-cpu --n-per-gpu 125 --geometry auto -v > gpurun_out/n1_auto_linear_box_125.json 2> gpurun_out/n1_auto_linear_box_125.err
python bench.py --steps 20 --no-cpu --workload nonlinear_bowl --geometry auto -v > gpurun_out/n1_auto_bowl_99.json 2> gpurun_out/n1_auto_bowl_99.err
set +x
for f in gpurun_out/n1_*.json; do echo "== $f"; cut -c1-330 $f; done
