#!/usr/bin/env bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list (+ optional full capture of the top kernel).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== bench c3"; timeout 900 python bench.py --steps 30 --warmup 5 --torch-gpu-baseline 2>gpurun_out/bench_c3.err | tee gpurun_out/bench_c3.json | tail -2; tail -5 gpurun_out/bench_c3.err
echo "== bench c2"; timeout 600 python bench.py --workload c2 --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_c2.err | tee gpurun_out/bench_c2.json | tail -2
if [ "${NCU:-1}" = "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log | cut -c1-300
fi
if [ "${NCUFULL:-0}" = "1" ]; then
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"encode_gemm2_kernel|wgrad_kernel|rescore_topk_kernel|decode_kernel|adam_rows_kernel" -s 15 -c 5 -o gpurun_out/prof_step_final -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
fi
