# gpurun recipe: where the gemm_tc time goes per shape — product library, then the debug-knob build (liblds_b200_dbg.so:
# LDS_TC_DEBUG=2 main loop only, =1 epilogue without global stores), then ncu --set full captures of single shapes.
mkdir -p gpurun_out
python tests/gpu_gemm_bench.py _product > gpurun_out/gemm_bench_product.log 2>&1; echo product rc=$?
DBG=$PWD/latent_diffusion_speech_b200/liblds_b200_dbg.so
LDS_B200_LIB=$DBG LDS_TC_DEBUG=2 python tests/gpu_gemm_bench.py _mainloop > gpurun_out/gemm_bench_mainloop.log 2>&1; echo mainloop rc=$?
LDS_B200_LIB=$DBG LDS_TC_DEBUG=1 python tests/gpu_gemm_bench.py _nostore > gpurun_out/gemm_bench_nostore.log 2>&1; echo nostore rc=$?
cap() { name=$1; shift
  python tests/gpu_gemm_one.py "$@" > /dev/null 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o gpurun_out/$name python tests/gpu_gemm_one.py "$@" > gpurun_out/$name.log 2>&1; echo $name rc=$?; }
cap ncu_ff2_L0_split 1 55296 1024 256 1 0 2 2 R
cap ncu_geglu_L0_split 1 55296 256 2048 1 2 2 2
cap ncu_lin256_split 1 55296 256 256 1 0 2 0 R
cap ncu_conv3_L0_split 64 864 256 256 3 0 2 0 R
cap ncu_geglu_L0_bf16 1 55296 256 2048 1 2 1 1
cap ncu_ff2_L0_bf16 1 55296 1024 256 1 0 1 1 R
ls -la gpurun_out/*.ncu-rep
