#!/usr/bin/env bash
# One GPU-box visit: parity tests, smoke, bench lines (c3 with both baselines, c2), ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r2f_tests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench c3"; timeout 900 python bench.py --steps 30 --warmup 5 2>gpurun_out/r2f_c3.err | tee gpurun_out/r2f_c3.json | cut -c1-400; tail -3 gpurun_out/r2f_c3.err
echo "== bench c2"; timeout 300 python bench.py --workload c2 --steps 30 --warmup 5 --no-cpu-baseline --no-torch-gpu-baseline --no-disk-leg 2>gpurun_out/r2f_c2.err | tee gpurun_out/r2f_c2.json | cut -c1-300
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 3 --warmup 3 --preheat-s 0 --e2e ring --no-cpu-baseline --no-torch-gpu-baseline --no-disk-leg > gpurun_out/r2f_ncu_bench.log 2>&1
tail -2 gpurun_out/r2f_ncu_bench.log | cut -c1-200; wc -l gpurun_out/r2f_launches.csv
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"encode_gemm2_kernel|wgrad_dec_kernel|wgrad_enc_kernel|rescore_topk_kernel|decode_kernel|adam_rows_kernel" -s 18 -c 6 -o gpurun_out/r2f_prof_step -f python bench.py --steps 3 --warmup 3 --preheat-s 0 --e2e ring --no-cpu-baseline --no-torch-gpu-baseline --no-disk-leg > gpurun_out/r2f_ncu_full.log 2>&1
grep -E "Profiling|Report" gpurun_out/r2f_ncu_full.log | cut -c1-120
