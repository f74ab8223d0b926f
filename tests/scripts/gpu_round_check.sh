# gpurun recipe of the round-end check: every GPU test, smoke, the default bench exactly as the driver runs it, the reference arm.
# usage: gpurun --timeout 1500 -- "bash tests/scripts/gpu_round_check.sh"   (results land in gpurun_out/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo gpu_tests rc=$?; tail -2 gpurun_out/gpu_tests.log
timeout 400 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; grep "^smoke" gpurun_out/smoke.log
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; echo bench rc=$? wall $(( $(date +%s) - T0 )) s; tail -1 gpurun_out/bench.log | cut -c1-200
T0=$(date +%s); timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo bench_reference rc=$? wall $(( $(date +%s) - T0 )) s; tail -1 gpurun_out/bench_reference.log | cut -c1-200
timeout 400 python bench.py --workload dpm20_b64_t864_bf16 --no-cpu-baseline --no-gpu-eager --no-strong --no-vocoder --no-units --no-train-loss > gpurun_out/bench_bf16.log 2>&1; echo bench_bf16 rc=$?; tail -1 gpurun_out/bench_bf16.log | cut -c1-160
python tests/gpu_frontend_once.py vocoder > gpurun_out/vocoder_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/vocoder_launches.csv python tests/gpu_frontend_once.py vocoder > gpurun_out/vocoder_ncu.log 2>&1; echo vocoder launches rc=$?
