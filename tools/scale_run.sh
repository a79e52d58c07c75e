#!/bin/bash
# The multi-GPU legs of BASELINE.json's configs on ONE 8-GPU box (gpurun --gpus 8):
#   configs[1]/north-star target: linear box, degree 4, 125^3 cells per GPU -> 1.0e9 dofs on 8 GPUs
#   configs[3]: Westervelt bowl, degree 4, 99^3 cells per GPU -> 4.99e8 dofs on 8 GPUs
#   configs[2]: piston, degree 5, ~1e8 dofs on 4 GPUs (58^3 per GPU) and on 2 GPUs (74^3 per GPU),
#               run side by side on disjoint GPUs
# The 1-GPU denominators come from tools/single_gpu_refs.sh.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
set -x
timeout 600 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 20 --n-per-gpu 125 -v --watchdog 500 \
    > gpurun_out/n8_linear_box_125.json 2> gpurun_out/n8_linear_box_125.err
timeout 400 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 20 --workload nonlinear_bowl -v --watchdog 350 \
    > gpurun_out/n8_bowl_99.json 2> gpurun_out/n8_bowl_99.err
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 400 $TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 20 \
    --workload linear_piston --n-per-gpu 58 -v --watchdog 350 > gpurun_out/n4_piston_58.json 2> gpurun_out/n4_piston_58.err &
CUDA_VISIBLE_DEVICES=4,5 timeout 400 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 20 \
    --workload linear_piston --n-per-gpu 74 -v --watchdog 350 > gpurun_out/n2_piston_74.json 2> gpurun_out/n2_piston_74.err &
wait
set +x
for f in gpurun_out/n8_*.json gpurun_out/n4_*.json gpurun_out/n2_*.json; do echo "== $f"; cut -c1-330 $f; done
tail -n 4 gpurun_out/n8_*.err gpurun_out/n4_*.err gpurun_out/n2_*.err
