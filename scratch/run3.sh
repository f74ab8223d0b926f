mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/kernels.log 2>&1; echo kernels rc=$?; tail -3 gpurun_out/kernels.log
for prec in fp32 bf16; do timeout 200 python tests/gpu_nfe_once.py $prec 64 864 3; done 2>&1 | tee gpurun_out/nfe_ab.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/parity.log 2>&1; echo parity rc=$?; tail -3 gpurun_out/parity.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.log 2>&1; echo bench rc=$?; cat gpurun_out/bench.log | tail -1
timeout 600 python bench.py --steps 2 --warmup 3 --workload dpm20_b64_t864_bf16 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; echo bench_bf16 rc=$?; tail -1 gpurun_out/bench_bf16.log
