#!/usr/bin/env bash
# BASELINE.json configs[4]: d_model 1536, d_sae 131072, ReLU + L1, batch 8192 per GPU, N GPUs data parallel, with the
# dead-feature AuxK sweep of SURVEY 8d (dead_threshold_tokens x k_aux).  8192 atoms are made unable to fire
# (--dead-atoms) so that AuxK has dead latents to work on once the threshold has passed; 20 warm-up steps put the
# 100k and 1M thresholds inside the timed region (10M is never reached in a short run: AuxK stays idle, as it would).
#   usage: scripts/run_c5_sweep.sh N out.jsonl ["thr:k thr:k ..."]      (default: the full 3 x 3 grid)
set -euo pipefail
N=${1:-8}; OUT=${2:-gpurun_out/c5_sweep_n$N.jsonl}
GRID=${3:-"10000000:256 10000000:512 10000000:1024 1000000:256 1000000:512 1000000:1024 100000:256 100000:512 100000:1024"}
: > "$OUT"
for cell in $GRID; do
  thr=${cell%%:*}; k=${cell##*:}
  for _once in 1; do
    if [ "$N" -gt 1 ]; then
      L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
    else
      L="python"
    fi
    $L bench.py --gpus "$N" --workload c5 --steps 8 --warmup 20 --preheat-s 0 --e2e ring --no-cpu-baseline \
       --no-torch-gpu-baseline --dead-atoms 8192 --dead-threshold $thr --k-aux $k 2>/dev/null | grep '^{' >> "$OUT" || echo "{\"error\": \"thr=$thr k=$k\"}" >> "$OUT"
    tail -1 "$OUT" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('thr=$thr k_aux=$k', d.get('value'), d.get('ms_per_step'), d.get('final',{}).get('n_dead'))"
  done
done
