#!/usr/bin/env bash
# Quick GPU visit for a kernel change: parity tests, full-size top-k check, c3/c2 bench (all under timeouts).
mkdir -p gpurun_out
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== topk check (c3 size)"; timeout 300 python scripts/gpu_check_topk.py 16384 2>&1 | tail -4
echo "== bench c3"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/try_c3.err | tee gpurun_out/try_c3.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['stage_ms_per_step'], d['final'], d['e2e']['value'])"
tail -3 gpurun_out/try_c3.err
echo "== bench c2"; timeout 600 python bench.py --workload c2 --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/try_c2.err | tee gpurun_out/try_c2.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['stage_ms_per_step'], d['final'])"
tail -3 gpurun_out/try_c2.err
