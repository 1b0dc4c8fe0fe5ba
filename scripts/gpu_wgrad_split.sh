#!/usr/bin/env bash
# Two-pass weight gradients (SAEV_B200_WGRAD_SPLIT=1): parity, then the c3 step with either form.
mkdir -p gpurun_out
echo "== tests (split)"; SAEV_B200_WGRAD_SPLIT=1 timeout 400 python -m pytest tests/test_gpu_golden.py -m gpu -x -q -k "topk or midsize or dense_features or staged or batchtopk" 2>&1 | tail -6
for sp in 0 1; do
  echo "== c3 split=$sp"
  SAEV_B200_WGRAD_SPLIT=$sp timeout 300 python bench.py --steps 20 --warmup 5 --e2e ring --no-cpu-baseline --no-torch-gpu-baseline --no-disk-leg 2>gpurun_out/ws_c3_$sp.err | tee gpurun_out/ws_c3_$sp.json | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],3),'ms', round(d['value']), {k:round(v,3) for k,v in d.get('stage_ms_per_step',{}).items()}, d['final'].get('check'))"
  tail -2 gpurun_out/ws_c3_$sp.err
done
