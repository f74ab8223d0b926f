# gpurun recipe of the round-end evidence: every GPU test, smoke, the default bench (as the driver runs it), bf16 + other workloads,
# one-evaluation launch lists with DRAM bytes, the bench launch list, launch lists of the units front-end and the vocoder, and one
# ncu --set full capture of gemm_tc inside the bench.  Each ncu run follows a plain run of the same command in the same call.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo gpu_tests rc=$?; tail -2 gpurun_out/gpu_tests.log
timeout 400 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; grep "^smoke" gpurun_out/smoke.log
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; echo bench rc=$? wall $(( $(date +%s) - T0 )) s; tail -1 gpurun_out/bench.log | cut -c1-200
T0=$(date +%s); timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo bench_reference rc=$? wall $(( $(date +%s) - T0 )) s; tail -1 gpurun_out/bench_reference.log | cut -c1-300
F="--no-cpu-baseline --no-gpu-eager --no-strong --no-vocoder --no-units --no-train-loss"
for w in dpm20_b64_t864_bf16 unipc10_b64_t864_bf16 shallow_dpm20_b32_t2584_fp32 shallow_dpm20_b32_t2584_bf16 dpm20_b1_t432_fp32 ddim20_b64_t864_fp32 pndm20_b64_t864_fp32; do
timeout 400 python bench.py --workload $w --steps 2 --warmup 3 $F > gpurun_out/bench_$w.log 2>&1; echo $w rc=$?; tail -1 gpurun_out/bench_$w.log | cut -c1-140
done
timeout 600 python bench.py --workload ddpm1000_b32_t864_bf16 --steps 1 --warmup 1 $F > gpurun_out/bench_ddpm1000_b32_t864_bf16.log 2>&1; echo ddpm rc=$?; tail -1 gpurun_out/bench_ddpm1000_b32_t864_bf16.log | cut -c1-140
for prec in fp32 bf16; do
python tests/gpu_nfe_once.py $prec 64 864 0 > gpurun_out/nfe_plain_$prec.log 2>&1 && \
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/nfe_launches_$prec.csv python tests/gpu_nfe_once.py $prec 64 864 0 > gpurun_out/nfe_ncu_$prec.log 2>&1; echo nfe $prec rc=$?
done
for what in units vocoder; do
python tests/gpu_frontend_once.py $what > gpurun_out/${what}_plain.log 2>&1 && \
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${what}_launches.csv python tests/gpu_frontend_once.py $what > gpurun_out/${what}_ncu.log 2>&1; echo $what launches rc=$?; tail -1 gpurun_out/${what}_plain.log
done
BENCH="python bench.py --steps 1 --warmup 3 $F"
$BENCH > gpurun_out/bench_plain.log 2>&1 && \
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 20700 -c 6900 --csv --log-file gpurun_out/bench_launches_fp32.csv $BENCH > gpurun_out/bench_ncu.log 2>&1; echo bench-ncu rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip 11700 -c 4 -f -o gpurun_out/prof_gemm_tc_bench $BENCH > gpurun_out/ncu_gemm.log 2>&1; echo ncu-gemm rc=$?
ls -la gpurun_out/*.ncu-rep gpurun_out/*launches*.csv 2>/dev/null | awk '{print $5, $9}'
