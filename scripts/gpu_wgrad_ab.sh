#!/usr/bin/env bash
# A/B of the weight-gradient kernel options on one box: warps per block, heavy-atom path with dense features.
mkdir -p gpurun_out
run() { # label, env..., -- bench args
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e ring $EXTRA 2>>gpurun_out/x.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$label', round(d['ms_per_step'],4), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()})"
}
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run "default" SAEV_B200_WGRAD_WPB=1
EXTRA="--dense-features 8"
run "dense8 heavy=1" SAEV_B200_WGRAD_HEAVY=1
EXTRA="--dense-features 64"
run "dense64 heavy=1" SAEV_B200_WGRAD_HEAVY=1
tail -3 gpurun_out/x.err
