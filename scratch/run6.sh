mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention" > gpurun_out/attn_tests.log 2>&1; rc=$?; echo attn_tests rc=$rc; tail -3 gpurun_out/attn_tests.log
if [ $rc -ne 0 ]; then exit 1; fi
for pa in 1 0 1 0; do echo "PA128=$pa"; AB=64 LDS_ATT_PA128=$pa timeout 120 python tests/gpu_bench_attention.py 2>&1 | grep "parts=3"; done | tee gpurun_out/attn_ab.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pa128.log 2>&1; echo bench rc=$?
LDS_ATT_PA128=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pa64.log 2>&1; echo bench rc=$?
grep -o '"value": [0-9.]*\|"attention": {[^}]*}' gpurun_out/bench_pa128.log gpurun_out/bench_pa64.log
