#!/bin/bash
# A/B of the shared-last local renumbering on 2 GPUs: ranks split along z (the shared face is the
# scattered one in the index map's numbering) and along x (already contiguous), with and without.
# Usage: [BENCH_EXTRA="--dtype f32" TAG=_f32] bash tools/r2_renumber_ab.sh [case ...]   (default: all four)
mkdir -p gpurun_out/r2d
run() {  # name, extra flags
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port $((29620 + RANDOM % 50)) \
    bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu --no-affine --no-extras ${BENCH_EXTRA:-} $2 > gpurun_out/r2d/$1${TAG:-}.json 2> gpurun_out/r2d/$1${TAG:-}.err
  python - "$1" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2d/{n}{__import__('os').environ.get('TAG', '')}.json").read().strip().splitlines()[-1])
    p = d.get("multi_gpu_parity") or {}
    print(n, "GDoF/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 4), "parity", p.get("ok"), (p.get("unstructured_like") or {}).get("ok"))
except Exception as e:
    print(n, "FAILED", e)
    print(open(f"gpurun_out/r2d/{n}{__import__('os').environ.get('TAG', '')}.err").read()[-1500:])
PY
}
want=${@:-z_renumber z_plain x_renumber x_plain}
for c in $want; do
  case $c in
    z_renumber) run z_renumber "--rank-grid 1x1x2" ;;
    z_plain) run z_plain "--rank-grid 1x1x2 --no-renumber" ;;
    x_renumber) run x_renumber "" ;;
    x_plain) run x_plain "--no-renumber" ;;
  esac
done
