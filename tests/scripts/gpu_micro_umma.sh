# gpurun recipe: tcgen05.mma issue-cost microbenchmarks (tests/micro/bench_umma.cu; build it first with
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/micro/bench_umma tests/micro/bench_umma.cu
# — the binary is git-ignored but travels with the gpurun snapshot).  Results land in gpurun_out/bench_umma.log;
# copy them to profiles/ when they inform a design decision.
mkdir -p gpurun_out
timeout 120 tests/micro/bench_umma > gpurun_out/bench_umma.log 2>&1; echo all rc=$?
timeout 60 tests/micro/bench_umma x >> gpurun_out/bench_umma.log 2>&1; echo probes rc=$?
tail -12 gpurun_out/bench_umma.log
