# gpurun recipe: A/B of the d=64 split-f16 attention configurations (LDS_ATT_DUAL64 = 0 single CTA / 1 two CTAs per SM)
mkdir -p gpurun_out
for v in 1 0; do
  echo "== LDS_ATT_DUAL64=$v kernel tests"
  LDS_ATT_DUAL64=$v timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -p no:cacheprovider -k "attention" 2>&1 | tail -2
  echo "== LDS_ATT_DUAL64=$v attention micro-bench"
  LDS_ATT_DUAL64=$v AB=64 timeout 300 python tests/gpu_bench_attention.py 2>&1 | grep "parts=2"
done
for v in 1 0; do
  LDS_ATT_DUAL64=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-strong --no-vocoder --no-units --no-train-loss > gpurun_out/bench_dual64_$v.log 2>gpurun_out/bench_dual64_$v.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_dual64_$v.log") if l.startswith("{")][-1])
print("dual64=$v", round(d["value"]), round(d["ms_per_step"],1), {k:round(x["ms_per_step"],1) for k,x in d["kernel_classes"].items()}, d["clocks"])
PY
done
