#!/usr/bin/env bash
# Stage the UNMODIFIED reference package where a gpurun box can import it: baseline/_ref/ is git-ignored (never part
# of this repository's history) but travels with the gpurun snapshot.  Used by the opt-in end-to-end drop-in test:
#     bash scripts/stage_reference.sh
#     gpurun -- 'SAEV_B200_REF_SRC=baseline/_ref python -m pytest tests/test_gpu_real_saev.py -m gpu -x -q'
set -euo pipefail
cd "$(dirname "$0")/.."
mkdir -p baseline/_ref
rm -rf baseline/_ref/saev
cp -r /root/reference/src/saev baseline/_ref/saev
echo "staged $(find baseline/_ref/saev -name '*.py' | wc -l) files under baseline/_ref/saev"
