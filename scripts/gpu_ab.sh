#!/usr/bin/env bash
# usage: gpu_ab.sh "LABEL ENV=V ..." ...   -- one bench.py (c3, 30 steps) per argument, stage times printed
mkdir -p gpurun_out
for spec in "$@"; do
  set -- $spec; label=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e ring $EXTRA 2>>gpurun_out/x.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$label', round(d['ms_per_step'],4), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()})"
done
tail -3 gpurun_out/x.err
