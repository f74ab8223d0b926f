# gpurun recipe: ncu --set full captures of the final (split-f16) attention kernels at B=64 and of the vocoder's N=64 / folded GEMMs
mkdir -p gpurun_out
AB=64 python tests/gpu_bench_attention.py > gpurun_out/attn_plain.log 2>&1; cat gpurun_out/attn_plain.log
# launches per (T, parts) block: 13 qkv GEMMs + 13 attention kernels (+ casts); the split-f16 d=32 kernel is the first attention_tc instance
AB=64 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_attn_t0_split_final python tests/gpu_bench_attention.py > gpurun_out/ncu_attn_t0.log 2>&1; echo attn_t0 rc=$?
AB=64 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 29 -c 1 -f -o gpurun_out/prof_attn_t1_split_final python tests/gpu_bench_attention.py > gpurun_out/ncu_attn_t1.log 2>&1; echo attn_t1 rc=$?
python tests/gpu_frontend_once.py vocoder > /dev/null 2>&1 && timeout 400 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 260 -c 3 -f -o gpurun_out/prof_voc_gemm python tests/gpu_frontend_once.py vocoder > gpurun_out/ncu_voc.log 2>&1; echo voc rc=$?
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
