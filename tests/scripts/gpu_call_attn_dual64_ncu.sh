# gpurun recipe: ncu --set full captures of the split-f16 attention kernel at head dim 48 (T=432) and 64 (T=216), B=64, final configuration
mkdir -p gpurun_out
AB=64 python tests/gpu_bench_attention.py > gpurun_out/attn_plain.log 2>&1; cat gpurun_out/attn_plain.log
AB=64 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 29 -c 1 -f -o gpurun_out/prof_attn_t1_dual64 python tests/gpu_bench_attention.py > gpurun_out/ncu_attn_t1.log 2>&1; echo attn_t1 rc=$?
AB=64 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 55 -c 1 -f -o gpurun_out/prof_attn_t2_dual64 python tests/gpu_bench_attention.py > gpurun_out/ncu_attn_t2.log 2>&1; echo attn_t2 rc=$?
ls -la gpurun_out/*dual64.ncu-rep | awk '{print $5, $9}'
