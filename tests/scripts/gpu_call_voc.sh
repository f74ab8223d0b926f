# gpurun recipe: vocoder — kernel + vocoder tests, the bench's vocoder leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vocoder.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/voc_tests.log 2>&1; echo voc rc=$?; tail -25 gpurun_out/voc_tests.log | grep -v Warn
python - <<'PY' > gpurun_out/voc_leg.log 2>&1
import json, torch, bench
print(json.dumps(bench.vocoder_leg(torch.device("cuda:0"))))
PY
tail -1 gpurun_out/voc_leg.log | cut -c1-330
