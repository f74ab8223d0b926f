# gpurun recipe: state check of the committed kernels — full GPU test suite, smoke, default bench (as the driver runs it), bf16 bench,
# latency / long-sequence workloads, one-evaluation launch lists with DRAM bytes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo gpu_tests rc=$?; tail -2 gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; tail -3 gpurun_out/smoke.log
timeout 700 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; echo bench rc=$?; tail -1 gpurun_out/bench.log | cut -c1-300
timeout 400 python bench.py --workload dpm20_b64_t864_bf16 --no-cpu-baseline --no-gpu-eager --no-strong --no-vocoder --no-units > gpurun_out/bench_bf16.log 2>&1; echo bench_bf16 rc=$?; tail -1 gpurun_out/bench_bf16.log | cut -c1-200
for w in shallow_dpm20_b32_t2584_fp32 dpm20_b1_t432_fp32; do
timeout 400 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-strong --no-vocoder --no-units > gpurun_out/bench_$w.log 2>&1; echo $w rc=$?; tail -1 gpurun_out/bench_$w.log | cut -c1-160
done
for prec in fp32 bf16; do
python tests/gpu_nfe_once.py $prec 64 864 0 > gpurun_out/nfe_plain_$prec.log 2>&1 && \
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/nfe_launches_$prec.csv python tests/gpu_nfe_once.py $prec 64 864 0 > gpurun_out/nfe_ncu_$prec.log 2>&1; echo nfe $prec rc=$?
done
