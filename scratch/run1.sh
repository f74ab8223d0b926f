set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k groupnorm > gpurun_out/gn_tests.log 2>&1; echo gn_tests rc=$?
tail -5 gpurun_out/gn_tests.log
for mode in 2 1 2 1; do for prec in fp32 bf16; do echo "GN_MODE=$mode"; LDS_GN_MODE=$mode timeout 200 python tests/gpu_nfe_once.py $prec 64 864 3; done; done 2>&1 | tee gpurun_out/gn_ab.log
for mode in 2 0; do echo "GN_MODE=$mode"; LDS_GN_MODE=$mode timeout 200 python tests/gpu_nfe_once.py fp32 32 2584 2; LDS_GN_MODE=$mode timeout 200 python tests/gpu_nfe_once.py fp32 1 864 5; done 2>&1 | tee -a gpurun_out/gn_ab.log
