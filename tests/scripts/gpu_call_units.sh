# gpurun recipe: units front-end GPU tests + the GEMM kernel tests (GELU epilogue change) + the headline bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/units_tests.log 2>&1; echo units rc=$?; tail -25 gpurun_out/units_tests.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -p no:cacheprovider --tb=short > gpurun_out/kernel_tests.log 2>&1; echo kernels rc=$?; tail -3 gpurun_out/kernel_tests.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_configs.py -q -m gpu -x -p no:cacheprovider --tb=short > gpurun_out/parity_tests.log 2>&1; echo parity rc=$?; tail -3 gpurun_out/parity_tests.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-strong --no-vocoder --no-units > gpurun_out/bench.log 2>gpurun_out/bench.err; echo bench rc=$?; tail -1 gpurun_out/bench.log | cut -c1-200
