#!/bin/bash
# Round-end evidence on one GPU: parity tests, the default bench line, and the ncu launch list of
# the SAME bench command (per-launch gpu__time_duration; cold-cache and serialised - shares only).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || tail -5 gpurun_out/bench_final.err
cut -c1-400 gpurun_out/bench_final.json
python bench.py --impl reference > gpurun_out/bench_final_reference.json 2>> gpurun_out/bench_final.err
cut -c1-300 gpurun_out/bench_final_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r01_bench_launches.csv \
    python bench.py --no-cpu > gpurun_out/bench_ncu.json 2> gpurun_out/bench_ncu.err
wc -l gpurun_out/r01_bench_launches.csv
bash tools/single_gpu_refs.sh 2>&1 | tail -14
