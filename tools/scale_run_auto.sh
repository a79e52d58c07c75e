#!/bin/bash
# tools/scale_run.sh with --geometry auto (rectilinear cells keep 6 factors): the 8-GPU legs at the target sizes.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
set -x
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --n-per-gpu 125 --geometry auto -v --watchdog 500 \
    > gpurun_out/n8_auto_linear_box_125.json 2> gpurun_out/n8_auto_linear_box_125.err
timeout 400 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 20 --workload nonlinear_bowl --geometry auto -v --watchdog 350 \
    > gpurun_out/n8_auto_bowl_99.json 2> gpurun_out/n8_auto_bowl_99.err
set +x
for f in gpurun_out/n8_auto_*.json; do echo "== $f"; cut -c1-330 $f; done
tail -n 3 gpurun_out/n8_auto_*.err
