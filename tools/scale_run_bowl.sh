#!/bin/bash
# The Westervelt bowl (configs[3], 99^3 cells per GPU -> 4.99e8 dofs on 8 GPUs), streamed and geometry=auto.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 20 --workload nonlinear_bowl -v --watchdog 250 \
    > gpurun_out/n8_bowl_99.json 2> gpurun_out/n8_bowl_99.err
timeout 300 $TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --steps 20 --workload nonlinear_bowl --geometry auto -v --watchdog 250 \
    > gpurun_out/n8_auto_bowl_99.json 2> gpurun_out/n8_auto_bowl_99.err
for f in gpurun_out/n8_bowl_99.json gpurun_out/n8_auto_bowl_99.json; do cut -c1-300 $f; done
