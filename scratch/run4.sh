mkdir -p gpurun_out
for prec in fp32 bf16; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/nfe_launches_$prec.csv python tests/gpu_nfe_once.py $prec 64 864 1 > gpurun_out/nfe_ncu_$prec.log 2>&1; echo rc=$?
done
