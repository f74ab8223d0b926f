mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_loss.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/loss_tests.log 2>&1; echo loss rc=$?; tail -30 gpurun_out/loss_tests.log
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider --tb=short 2>&1 | tail -2
