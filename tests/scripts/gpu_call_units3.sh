mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py -q -m gpu -p no:cacheprovider --tb=short 2>&1 | grep -v Warn | tail -12
