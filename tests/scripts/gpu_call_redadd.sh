# gpurun recipe: in-place residual GEMMs through TMA reduce-add — A/B on the headline + parity
mkdir -p gpurun_out
F="--steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-strong --no-vocoder --no-units --no-train-loss"
for v in 0 1; do LDS_RED_ADD=$v python bench.py $F > gpurun_out/bench_redadd$v.log 2>&1; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_redadd$v.log") if l.startswith("{")][-1])
print("red_add=$v", round(d["value"]), round(d["ms_per_step"],1), {k:round(x["ms_per_step"],1) for k,x in d["kernel_classes"].items() if x["ms_per_step"]>1})
PY
done
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_units.py tests/test_gpu_stress.py -q -m gpu -x -p no:cacheprovider --tb=short 2>&1 | tail -3
