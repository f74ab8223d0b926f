# gpurun recipe of the round-end check: GPU tests, smoke, headline + bf16 bench, the other BASELINE workloads.
# usage: gpurun --timeout 1500 -- "bash tests/scripts/gpu_round_check.sh"   (results land in gpurun_out/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo gpu_tests rc=$?; tail -2 gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo bench rc=$?; tail -1 gpurun_out/bench.log | cut -c1-400
timeout 600 python bench.py --workload dpm20_b64_t864_bf16 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; echo bench_bf16 rc=$?; tail -1 gpurun_out/bench_bf16.log | cut -c1-200
for w in unipc10_b64_t864_bf16 shallow_dpm20_b32_t2584_bf16 shallow_dpm20_b32_t2584_fp32 ddim20_b64_t864_fp32 pndm20_b64_t864_fp32 dpm20_b1_t432_fp32 ddpm1000_b32_t864_bf16; do
timeout 400 python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; echo $w rc=$?; tail -1 gpurun_out/bench_$w.log | cut -c1-160
done
