# gpurun recipe: gemm_tc ring-depth / L2-prefetch / epilogue-form experiments with the debug-knob build
mkdir -p gpurun_out
DBG=$PWD/latent_diffusion_speech_b200/liblds_b200_dbg.so
run() { tag=$1; shift; env LDS_B200_LIB=$DBG "$@" python tests/gpu_gemm_bench.py _$tag > gpurun_out/gemm_bench_$tag.log 2>&1; echo $tag rc=$? $(tail -1 gpurun_out/gemm_bench_$tag.log | cut -c1-200); }
run base LDS_TC_PF=0
run pf4 LDS_TC_PF=4
run pf8 LDS_TC_PF=8
run staged LDS_TC_TMAEPI_MAXKB=8
run staged_pf4 LDS_TC_TMAEPI_MAXKB=8 LDS_TC_PF=4
run na4 LDS_TC_NA=4
run na4_pf4 LDS_TC_NA=4 LDS_TC_PF=4
run main_pf4 LDS_TC_DEBUG=2 LDS_TC_PF=4
run main_pf8 LDS_TC_DEBUG=2 LDS_TC_PF=8
