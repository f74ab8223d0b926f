mkdir -p gpurun_out
set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel --launch-skip 12 --launch-count 1 -o gpurun_out/prof_ln_t0 -f python tests/gpu_nfe_once.py fp32 64 864 1 > gpurun_out/ncu_ln.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_cluster_kernel --launch-skip 5 --launch-count 2 -o gpurun_out/prof_gn_t0 -f python tests/gpu_nfe_once.py fp32 64 864 1 > gpurun_out/ncu_gn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_cluster_kernel --launch-skip 5 --launch-count 2 -o gpurun_out/prof_gn_t0_bf16 -f python tests/gpu_nfe_once.py bf16 64 864 1 > gpurun_out/ncu_gn_bf16.log 2>&1
tail -3 gpurun_out/ncu_ln.log gpurun_out/ncu_gn.log
