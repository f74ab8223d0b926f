# gpurun recipe: compute-sanitizer over the kernel-level GPU tests (small shapes).  ONE tool per gpurun call (B200_PROFILING.md):
#   gpurun --timeout 1500 -- 'bash tests/scripts/gpu_sanitizer.sh memcheck'      (then racecheck, synccheck in separate calls)
# The summary lands in gpurun_out/sanitizer_<tool>.log; copy it to profiles/ (r02_sanitizer_<tool>.txt).
TOOL=${1:-memcheck}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solver_kernels.py -q -m gpu -x -p no:cacheprovider > gpurun_out/sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitizer_plain.log; exit 1; }
timeout 1300 compute-sanitizer --tool $TOOL --print-limit 20 --error-exitcode 0 \
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solver_kernels.py -q -m gpu -p no:cacheprovider -k "${SAN_K:-tc or solver or bit_exact or groupnorm or layernorm}" \
  > gpurun_out/sanitizer_$TOOL.log 2>&1
echo "sanitizer $TOOL rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" gpurun_out/sanitizer_$TOOL.log | sort | uniq -c | sort -rn | head -20
tail -5 gpurun_out/sanitizer_$TOOL.log
