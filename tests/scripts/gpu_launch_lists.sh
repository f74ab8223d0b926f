# gpurun recipe of the ncu launch lists committed under profiles/ (one denoiser evaluation per mode with DRAM bytes, and ~6900
# launches = one timed step of the bench command).  Each ncu run follows a plain run of the same command in the same call.
mkdir -p gpurun_out
for prec in fp32 bf16; do
python tests/gpu_nfe_once.py $prec 64 864 0 > gpurun_out/nfe_plain_$prec.log 2>&1 && \
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/nfe_launches_$prec.csv python tests/gpu_nfe_once.py $prec 64 864 0 > gpurun_out/nfe_ncu_$prec.log 2>&1; echo nfe $prec rc=$?
done
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-strong --no-vocoder --no-units --no-train-loss"
$BENCH > gpurun_out/bench_plain.log 2>&1 && \
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${SKIP:-20800} -c ${COUNT:-6900} --csv --log-file gpurun_out/bench_launches_fp32.csv $BENCH > gpurun_out/bench_ncu.log 2>&1; echo bench-ncu rc=$?
tail -1 gpurun_out/bench_plain.log | cut -c1-200
