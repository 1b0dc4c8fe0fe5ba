#!/usr/bin/env bash
# Full GPU suite, then the c3 step under a few settings of the screen's environment knobs (A/B, same box).
mkdir -p gpurun_out
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2g_tests.log
run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --preheat-s 1 --e2e ring --no-cpu-baseline --no-torch-gpu-baseline --no-disk-leg 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d['stage_ms_per_step']; print(round(d['ms_per_step'],3),'ms enc',round(s['encode_gemm'],3),'resc',round(s['rescore'],3), d['final']['screen'])"; }
run X=1
run SAEV_B200_TRIGGER=128
run SAEV_B200_TRIGGER=256
run SAEV_B200_GUESS_Q=0.02
run SAEV_B200_GUESS_S=0.98
run SAEV_B200_RESCORE_WPB=2
