mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_cluster_kernel --launch-skip 55 --launch-count 3 -o gpurun_out/prof_gn_v4_bf16 -f python tests/gpu_nfe_once.py bf16 64 864 0 > gpurun_out/ncu_gn_v4.log 2>&1; echo rc=$?
