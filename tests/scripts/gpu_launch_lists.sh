# gpurun recipe of the ncu launch lists committed under profiles/ (one denoiser evaluation per mode with DRAM bytes,
# and 1400 launches inside the timed step of the bench command).  The bench list takes ~5 min: ncu also intercepts the skipped launches.
mkdir -p gpurun_out
for prec in fp32 bf16; do
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/nfe_launches_$prec.csv python tests/gpu_nfe_once.py $prec 64 864 0 > gpurun_out/nfe_ncu_$prec.log 2>&1; echo nfe $prec rc=$?
done
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 21500 -c 1400 --csv --log-file gpurun_out/bench_launches_fp32.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo bench-ncu rc=$?
