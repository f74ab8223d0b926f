# gpurun recipe: ncu launch lists (duration + DRAM bytes per launch) of ONE denoiser evaluation per mode, B=64 x T=864 — the source of
# profiles/r02_gemm_traffic.json (tests/scripts/make_gemm_traffic.py).  Each ncu run follows a plain run of the same command.
mkdir -p gpurun_out
for prec in fp32 bf16; do
python tests/gpu_nfe_once.py $prec 64 864 0 > gpurun_out/nfe_plain_$prec.log 2>&1 && \
timeout 140 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/nfe_launches_$prec.csv python tests/gpu_nfe_once.py $prec 64 864 0 > gpurun_out/nfe_ncu_$prec.log 2>&1; echo nfe $prec rc=$?
done
wc -l gpurun_out/nfe_launches_*.csv
