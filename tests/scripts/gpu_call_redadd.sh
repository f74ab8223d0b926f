# gpurun recipe: in-place transformers (proj_out through TMA reduce-add) — headline + parity + stress
mkdir -p gpurun_out
F="--steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-strong --no-vocoder --no-units --no-train-loss"
python bench.py $F > gpurun_out/bench_inplace.log 2>&1; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_inplace.log") if l.startswith("{")][-1])
print("in-place", round(d["value"]), round(d["ms_per_step"],1), {k:round(x["ms_per_step"],1) for k,x in d["kernel_classes"].items() if x["ms_per_step"]>1})
PY
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_parity_configs.py tests/test_gpu_stress.py tests/test_gpu_train_loss.py -q -m gpu -x -p no:cacheprovider --tb=short 2>&1 | tail -3
