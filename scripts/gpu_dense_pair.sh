#!/usr/bin/env bash
# Dense (ReLU) path on the CTA-pair kernel: parity tests, then the c5 step on one GPU with either kernel and either split.
mkdir -p gpurun_out
echo "== tests (pair kernel)"; timeout 400 python -m pytest tests/test_gpu_golden.py tests/test_gpu_dropin.py tests/test_gpu_logblock.py -m gpu -x -q -k "relu or split or matryoshka or logblock or log_" 2>&1 | tail -12
for terms in 6 3; do for pair in 0 1; do
  echo "== c5 terms=$terms pair=$pair"
  SAEV_B200_DENSE_TERMS=$terms SAEV_B200_DENSE_PAIR=$pair timeout 300 python bench.py --workload c5 --steps 5 --warmup 3 --preheat-s 0 --e2e ring --no-cpu-baseline --no-torch-gpu-baseline --no-disk-leg 2>gpurun_out/dp_c5_${terms}_${pair}.err | tee gpurun_out/dp_c5_${terms}_${pair}.json | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],2),'ms', round(d['value']), {k:round(v,2) for k,v in d.get('stage_ms_per_step',{}).items()}, d['final'].get('check'))"
  tail -2 gpurun_out/dp_c5_${terms}_${pair}.err
done; done
