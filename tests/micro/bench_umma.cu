// Micro-benchmark: cycles per tcgen05.mma (kind::f16, cta_group::1, M=128, K=16) as a function of N and swizzle width,
// operands in shared memory (SS).  One CTA per SM, one issuing thread; operands are whatever is in smem (zeros).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../latent_diffusion_speech_b200/csrc/tc_ptx.cuh"
using namespace lds::ptx;

template <int N, int SWZ, int NACC = 1>
__global__ void __launch_bounds__(128) k(long long* out, int iters, int ksteps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar = base + 96 * 1024;
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 96 * 1024 + 16);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(slot), 256);
  fence_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    const uint64_t ad = umma_desc_kmajor(base, SWZ), bd = umma_desc_kmajor(base + 32 * 1024, SWZ);
    const uint32_t idesc = umma_idesc_bf16(128, N);
    long long t0 = clock64();
    // NACC independent accumulators (disjoint TMEM column ranges), used round-robin: separates the latency of a dependent
    // accumulate chain from the issue rate of the tensor pipe
    for (int i = 0; i < iters; ++i)
      for (int ks = 0; ks < ksteps; ++ks) umma_bf16(tm + (uint32_t)(((i * ksteps + ks) % NACC) * N), ad + 2 * ks, bd + 2 * ks, idesc, 1u);
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 256); }
}

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}

// A operand from TMEM (TS): A tile lives in TMEM columns [256, 256+8*ksteps)
template <int N, int SWZ>
__global__ void __launch_bounds__(128) kts(long long* out, int iters, int ksteps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar = base + 96 * 1024;
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 96 * 1024 + 16);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(slot), 512);
  fence_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    const uint64_t bd = umma_desc_kmajor(base + 32 * 1024, SWZ);
    const uint32_t idesc = umma_idesc_bf16(128, N);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
      for (int ks = 0; ks < ksteps; ++ks) umma_bf16_ts(tm, tm + 256 + 8 * ks, bd + 2 * ks, idesc, 1u);
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, int SWZ>
void run_ts(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 2000, ksteps = SWZ / 32;
  cudaFuncSetAttribute(kts<N, SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  kts<N, SWZ><<<grid, 128, 100 * 1024>>>(d, iters, ksteps);
  kts<N, SWZ><<<grid, 128, 100 * 1024>>>(d, iters, ksteps);
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("TS %-23s grid=%3d  %.1f cycles/MMA  (ideal math %d)  %s\n", name, grid, (double)h / (iters * ksteps), 128 * N / 256, cudaGetErrorString(e));
  cudaFree(d);
}


// CTA pair: tcgen05.mma.cta_group::2 (M = 256 over two SMs), issued by the leader; cycles per instruction vs N.
template <int N, int SWZ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) kpair(long long* out, int iters, int ksteps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar = base + 96 * 1024;
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 96 * 1024 + 16);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc_pair(smem_u32(slot), 256);
  fence_async_smem();
  tc_fence_before(); __syncthreads(); cluster_sync(); tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0 && cluster_ctarank() == 0) {
    const uint64_t ad = umma_desc_kmajor(base, SWZ), bd = umma_desc_kmajor(base + 32 * 1024, SWZ);
    const uint32_t idesc = umma_idesc_bf16(256, N);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
      for (int ks = 0; ks < ksteps; ++ks) umma_bf16_pair(tm, ad + 2 * ks, bd + 2 * ks, idesc, 1u);
    umma_commit_pair(bar, 1);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads(); cluster_sync();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc_pair(tm, 256); }
}

template <int N, int SWZ>
void run_pair(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 2000, ksteps = SWZ / 32;
  cudaFuncSetAttribute(kpair<N, SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  kpair<N, SWZ><<<grid, 128, 100 * 1024>>>(d, iters, ksteps);
  kpair<N, SWZ><<<grid, 128, 100 * 1024>>>(d, iters, ksteps);
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("PAIR %-21s grid=%3d  %.1f cycles/MMA  (ideal math %d)  %s\n", name, grid, (double)h / (iters * ksteps), 256 * N / 512, cudaGetErrorString(e));
  cudaFree(d);
}

template <int N, int SWZ, int NACC = 1>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 2000, ksteps = SWZ / 32;
  cudaFuncSetAttribute(k<N, SWZ, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<N, SWZ, NACC><<<grid, 128, 100 * 1024>>>(d, iters, ksteps);
  k<N, SWZ, NACC><<<grid, 128, 100 * 1024>>>(d, iters, ksteps);
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("%-26s grid=%3d  %.1f cycles/MMA  (ideal math %d)  %s\n", name, grid, (double)h / (iters * ksteps), 128 * N / 256, cudaGetErrorString(e));
  cudaFree(d);
}

// Generic shape / operand-major probe for the attention redesign: M = 64 or 128, any N, B operand K-major
// or MN-major (instruction-descriptor bit 16; the shared-memory descriptor then strides 16 rows of 128 bytes per k-step).
// Operands are zeros: only the issue cost is measured.
template <int M, int N, int BMN>
__global__ void __launch_bounds__(128) kx(long long* out, int iters, int ksteps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar = base + 96 * 1024;
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 96 * 1024 + 16);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(slot), 256);
  fence_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    const uint64_t ad = umma_desc_kmajor(base, 128);
    uint64_t bd = umma_desc_kmajor(base + 32 * 1024, 128);
    if (BMN) bd |= (uint64_t)(1024 >> 4) << 16;            // leading byte offset between 64-element column blocks (MN-major)
    const uint32_t idesc = umma_idesc_bf16(M, N) | ((uint32_t)BMN << 16);
    const uint64_t bstep = BMN ? 128 : 2;                  // 16-byte units per k-step of 16 elements
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
      for (int ks = 0; ks < ksteps; ++ks) umma_bf16(tm, ad + 2 * ks, bd + bstep * ks, idesc, 1u);
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 256); }
}

template <int M, int N, int BMN>
void run_x(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 2000, ksteps = 4;
  cudaFuncSetAttribute(kx<M, N, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  kx<M, N, BMN><<<grid, 128, 100 * 1024>>>(d, iters, ksteps);
  kx<M, N, BMN><<<grid, 128, 100 * 1024>>>(d, iters, ksteps);
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("X %-26s grid=%3d  %.1f cycles/MMA  (ideal math %d)  %s\n", name, grid, (double)h / (iters * ksteps), M * N / 256, cudaGetErrorString(e));
  cudaFree(d);
}

int main(int argc, char** argv) {
  if (argc > 1 && argv[1][0] == 'a') {                     // bench_umma a : independent accumulators
    const int grid = 148;
    run<32, 128, 1>("M128 N32  1 accum", grid);
    run<32, 128, 2>("M128 N32  2 accum", grid);
    run<32, 128, 4>("M128 N32  4 accum", grid);
    run<32, 128, 8>("M128 N32  8 accum", grid);
    run<64, 128, 1>("M128 N64  1 accum", grid);
    run<64, 128, 2>("M128 N64  2 accum", grid);
    run<64, 128, 4>("M128 N64  4 accum", grid);
    run<128, 128, 1>("M128 N128 1 accum", grid);
    run<128, 128, 2>("M128 N128 2 accum", grid);
    run<64, 64, 2>("M128 N64 SW64 2 accum", grid);
    run<64, 64, 4>("M128 N64 SW64 4 accum", grid);
    return 0;
  }
  if (argc > 1 && argv[1][0] == 's') {                     // bench_umma s : K-major M = 64 probes only
    const int grid = 148;
    run_x<128, 256, 0>("M128 N256 B:K-major", grid);
    run_x<64, 64, 0>("M64  N64  B:K-major", grid);
    run_x<64, 128, 0>("M64  N128 B:K-major", grid);
    run_x<64, 256, 0>("M64  N256 B:K-major", grid);
    return 0;
  }
  if (argc > 1 && argv[1][0] == 'x') {                     // bench_umma x : only the probes of the attention redesign
    const int grid = 148;
    run_x<128, 128, 0>("M128 N128 B:K-major", grid);
    run_x<128, 256, 0>("M128 N256 B:K-major", grid);
    run_x<128, 128, 1>("M128 N128 B:MN-major", grid);
    run_x<128, 256, 1>("M128 N256 B:MN-major", grid);
    run_x<64, 64, 0>("M64  N64  B:K-major", grid);
    run_x<64, 128, 0>("M64  N128 B:K-major", grid);
    run_x<64, 256, 0>("M64  N256 B:K-major", grid);
    run_x<64, 256, 1>("M64  N256 B:MN-major", grid);
    return 0;
  }
  for (int grid : {148}) {
    run<32, 128>("M128 N32  SW128", grid);
    run<64, 128>("M128 N64  SW128", grid);
    run<64, 64>("M128 N64  SW64", grid);
    run<128, 128>("M128 N128 SW128", grid);
    run<256, 128>("M128 N256 SW128", grid);
    run<64, 128, 2>("M128 N64 SW128 2 accum", grid);
    run<64, 128, 4>("M128 N64 SW128 4 accum", grid);
    run<32, 128, 4>("M128 N32 SW128 4 accum", grid);
    run<128, 128, 2>("M128 N128 SW128 2 accum", grid);
    run<96, 128>("M128 N96  SW128", grid);
    run<192, 128>("M128 N192 SW128", grid);
    run_ts<32, 128>("M128 N32  B:SW128", grid);
    run_ts<64, 128>("M128 N64  B:SW128", grid);
    run_ts<64, 64>("M128 N64  B:SW64", grid);
    run_ts<96, 128>("M128 N96  B:SW128", grid);
    run_ts<128, 128>("M128 N128 B:SW128", grid);
    run_ts<192, 128>("M128 N192 B:SW128", grid);
    run_ts<256, 128>("M128 N256 B:SW128", grid);
    run_pair<32, 128>("M256 N32  SW128", 144);
    run_pair<64, 128>("M256 N64  SW128", 144);
    run_pair<64, 64>("M256 N64  SW64", 144);
    run_pair<96, 128>("M256 N96  SW128", 144);
    run_pair<128, 128>("M256 N128 SW128", 144);
    run_pair<128, 64>("M256 N128 SW64", 144);
    run_pair<192, 128>("M256 N192 SW128", 144);
    run_pair<256, 128>("M256 N256 SW128", 144);
  }
  return 0;
}
