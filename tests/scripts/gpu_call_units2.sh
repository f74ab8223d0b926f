# gpurun recipe: GEMM shape bench (producer-warp change), kernel + units + parity tests, smoke, and the default bench as the driver runs it
mkdir -p gpurun_out
python tests/gpu_gemm_bench.py _uniform_producer > gpurun_out/gemm_bench_uniform_producer.log 2>&1; echo gemm_bench rc=$? $(tail -1 gpurun_out/gemm_bench_uniform_producer.log | cut -c1-200)
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_units.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/units_tests.log 2>&1; echo kernels+units rc=$?; tail -4 gpurun_out/units_tests.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_configs.py tests/test_gpu_stress.py -q -m gpu -x -p no:cacheprovider --tb=short > gpurun_out/parity_tests.log 2>&1; echo parity rc=$?; tail -3 gpurun_out/parity_tests.log
timeout 400 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; grep smoke gpurun_out/smoke.log
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; echo bench rc=$? wall $(( $(date +%s) - T0 )) s; tail -1 gpurun_out/bench.log | cut -c1-200
timeout 400 python bench.py --workload dpm20_b64_t864_bf16 --no-cpu-baseline --no-gpu-eager --no-strong --no-vocoder --no-units > gpurun_out/bench_bf16.log 2>&1; echo bench_bf16 rc=$?; tail -1 gpurun_out/bench_bf16.log | cut -c1-200
