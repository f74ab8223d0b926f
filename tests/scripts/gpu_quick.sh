# gpurun recipe: kernel tests + parity suite + the three bench lines used while iterating on a kernel (no context legs)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solver_kernels.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-strong --no-vocoder --no-units --no-train-loss > gpurun_out/bench.log 2>gpurun_out/bench.err
python bench.py --steps 2 --warmup 3 --workload dpm20_b64_t864_bf16 --no-cpu-baseline --no-gpu-eager --no-strong --no-vocoder --no-units --no-train-loss > gpurun_out/bench_bf16.log 2>&1
python bench.py --steps 2 --warmup 2 --workload shallow_dpm20_b32_t2584_fp32 --no-cpu-baseline --no-gpu-eager --no-strong --no-vocoder --no-units --no-train-loss > gpurun_out/bench_t2584.log 2>&1
python bench.py --steps 3 --warmup 2 --workload dpm20_b1_t432_fp32 --no-cpu-baseline --no-gpu-eager --no-strong --no-vocoder --no-units --no-train-loss > gpurun_out/bench_b1.log 2>&1
for f in bench bench_bf16 bench_t2584 bench_b1; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$f.log") if l.startswith("{")][-1])
    print("$f", round(d["value"]), round(d["ms_per_step"],1), "gemm_frac", round(d["roofline"]["frac"],4), {k:round(v["ms_per_step"],1) for k,v in d["kernel_classes"].items()})
except Exception as e:
    print("$f FAILED", e); print(open("gpurun_out/$f.log").read()[-1500:])
PY
done
