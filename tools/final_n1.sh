#!/bin/bash
# Round-end evidence on one GPU: parity tests, the default bench line, the two reference arms, the
# ncu launch list of the SAME bench command (per-launch gpu__time_duration; cold-cache and
# serialised - shares only) and one `ncu --set full` capture of the stiffness kernel.
#   gpurun --timeout 900 -- bash tools/final_n1.sh        (outputs under gpurun_out/r2_final/)
O=gpurun_out/r2_final
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err || tail -5 $O/bench_n1.err
cut -c1-400 $O/bench_n1.json
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
cut -c1-300 $O/bench_reference.json
python bench.py --impl reference-cuda > $O/bench_reference_cuda.json 2>> $O/bench_reference.err
python bench.py --integrator leapfrog --no-cpu --no-affine > $O/bench_leapfrog.json 2> $O/bench_leapfrog.err
# profiler passes, each only after the plain command above exited 0
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-affine --no-extras > $O/ncu_list.log 2>&1
wc -l $O/launches.csv
ncu --set full --clock-control none --import-source on -k regex:stiffness_kernel -s 4 -c 2 -o $O/stiffness_p4_f64 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-affine --no-extras > $O/ncu_full.log 2>&1
ls -la $O
