mkdir -p gpurun_out
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; echo bench rc=$? wall $(( $(date +%s) - T0 )) s; tail -1 gpurun_out/bench.log | cut -c1-200
