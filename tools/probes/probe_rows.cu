// Micro-benchmark for the row-streaming conv's inner pipeline: cycles per N=96 MMA (M=128, K=16, fp16, SS) when
//   MODE 0: MMAs only, A plane stride (LBO) 2048 B (128-byte aligned)
//   MODE 1: MMAs only, LBO 2080 B (the 130-pixel rows of conv3x3_rows.cuh)
//   MODE 2: + a producer warp refilling 8-plane stages with 8 x 2080 B cp.async.bulk per 12 MMAs (source: 4 MB, L2 resident)
//   MODE 3: same, source streamed from a buffer far larger than L2 (HBM)
#include <cstdio>
#include <cuda_fp16.h>
#include "../../explorable-super-resolution_b200/csrc/ptx.cuh"
using namespace esr;
constexpr int STAGES = 8;
constexpr uint32_t STAGE_BYTES = 8 * 2080;
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) rate(long long* out, int iters, const uint8_t* src, size_t src_bytes) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* sp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_full = base, bar_empty = base + 128, bar_done = base + 256, slot = base + 264;
  const uint32_t wres = base + 1024, stage0 = wres + 3 * 8 * N * 16;   // weights of one chunk: [dx][8 planes][N][8]
  for (int i = threadIdx.x; i < (int)((3 * 8 * N * 16 + STAGES * STAGE_BYTES) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(sp + 1024)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(sp + 264);
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t lbo = MODE == 0 ? 2048u : 2080u;
  long long t0 = 0, t1 = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (elect_one()) {
      const uint64_t adesc_t = make_smem_desc(0u, lbo, 128u);
      const uint64_t bdesc_t = make_smem_desc(wres, N * 16, 128u);
      int s = 0; uint32_t ph = 0;
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (MODE >= 2) { mbar_wait(bar_full + 8 * s, ph); tc_fence_after(); }
        const uint64_t ad0 = adesc_t + (uint64_t)((stage0 + s * STAGE_BYTES) >> 4);
        const uint32_t d = tmem + (uint32_t)((it & 3) * N);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
            umma_f16(d, ad0 + (uint64_t)(j * 2 * (lbo >> 4) + dx), bdesc_t + (uint64_t)((dx * 8 + 2 * j) * N), idesc, 1);
        if (MODE >= 2) umma_commit(bar_empty + 8 * s);
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      umma_commit(bar_done);
    }
  } else if (warp == 1 && MODE >= 2) {
    int s = 0; uint32_t ph = 0;
    size_t off = (((size_t)blockIdx.x * 2654435761u) % (src_bytes / 2)) & ~(size_t)4095;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(bar_empty + 8 * s, ph ^ 1u);
      if (lane == 0) mbar_expect_tx(bar_full + 8 * s, 8 * 2080);
      __syncwarp();
      if (lane < 8) bulk_load(stage0 + s * STAGE_BYTES + lane * 2080, src + off + (size_t)lane * 1048576, 2080, bar_full + 8 * s);
      off += 4096;
      if (off + 8 * 1048576 + 4096 > src_bytes) off = 0;
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
  }
  if (threadIdx.x == 0) {
    mbar_wait(bar_done, 0);
    t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
template <int N, int MODE> void run(const char* name, const uint8_t* src, size_t src_bytes) {
  long long* d; cudaMalloc(&d, 148 * 8);
  const int smem = 1024 + 128 + 3 * 8 * N * 16 + STAGES * STAGE_BYTES;
  cudaFuncSetAttribute(rate<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 3000;
  rate<N, MODE><<<148, 128, smem>>>(d, iters, src, src_bytes); cudaDeviceSynchronize();
  rate<N, MODE><<<148, 128, smem>>>(d, iters, src, src_bytes);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  printf("%-44s N=%3d: %6.1f cycles per MMA (tensor %d)  %s\n", name, N, (double)mx / (iters * 12.0), N / 2, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  uint8_t* small; uint8_t* big;
  const size_t sb = 16ull << 20, bb = 2048ull << 20;
  cudaMalloc(&small, sb); cudaMemset(small, 0, sb);
  cudaMalloc(&big, bb); cudaMemset(big, 0, bb);
  run<96, 0>("MMA only, LBO 2048", small, sb);
  run<96, 1>("MMA only, LBO 2080", small, sb);
  run<96, 2>("MMA + bulk-copy producer (L2 source)", small, sb);
  run<96, 3>("MMA + bulk-copy producer (HBM source)", big, bb);
  run<192, 1>("MMA only, LBO 2080", small, sb);
  run<192, 2>("MMA + bulk-copy producer (L2 source)", small, sb);
  run<192, 3>("MMA + bulk-copy producer (HBM source)", big, bb);
  return 0;
}
