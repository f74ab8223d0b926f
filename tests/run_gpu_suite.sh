#!/bin/bash
# GPU-box driver script: kernel tests, parity tests, smoke, short bench. Logs under gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solver_kernels.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/kernels.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/parity.log
if [ "${SKIP_CONFIGS:-0}" != "1" ]; then
timeout 1800 python -m pytest tests/test_gpu_parity_configs.py -q -m gpu --tb=short -p no:cacheprovider --durations=10 > gpurun_out/parity_configs.log 2>&1
echo "parity_configs rc=$?" >> gpurun_out/parity_configs.log
fi
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1200 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 ${BENCH_FLAGS} > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.log
timeout 600 python bench.py --steps 2 --warmup 3 --workload dpm20_b64_t864_bf16 --no-cpu-baseline --no-gpu-eager --no-strong --no-vocoder --no-units --no-train-loss > gpurun_out/bench_bf16.log 2>&1
echo "bench_bf16 rc=$?" >> gpurun_out/bench_bf16.log
tail -5 gpurun_out/kernels.log; tail -5 gpurun_out/parity.log; tail -15 gpurun_out/parity_configs.log; tail -3 gpurun_out/smoke.log; tail -c 1500 gpurun_out/bench.log; tail -c 600 gpurun_out/bench.err; tail -c 800 gpurun_out/bench_bf16.log
