# gpurun recipe: vocoder tensor-core levels — kernel + vocoder tests, smoke, the bench's vocoder leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vocoder.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/voc_tests.log 2>&1; echo voc rc=$?; tail -25 gpurun_out/voc_tests.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_units.py -q -m gpu -x -p no:cacheprovider --tb=short 2>&1 | tail -2
python - <<'PY' > gpurun_out/voc_leg.log 2>&1
import json, torch, bench
print(json.dumps(bench.vocoder_leg(torch.device("cuda:0"))))
PY
tail -2 gpurun_out/voc_leg.log | cut -c1-600
