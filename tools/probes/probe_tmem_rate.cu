// Micro-benchmark: cycles per MMA (M=128, K=16, fp16) for
//   SS: A and B from shared memory (what conv3x3_tc_kernel does today), 9 MMAs per "row step"
//   TS: 3 tcgen05.cp (A -> TMEM) + 9 MMAs with A from TMEM (the strip-kernel plan)
#include <cstdio>
#include <cuda_fp16.h>
#include "../../explorable-super-resolution_b200/csrc/ptx.cuh"
using namespace esr;
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_cp(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
template <int N, int MODE>
__global__ void rate(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* sp = smem_raw + (base - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < 120000 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sp + 1024)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(base, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (threadIdx.x < 32) { tmem_alloc(base + 64, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(sp + 64);
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (threadIdx.x < 32) {
   if (elect_one()) {
    const uint32_t a_off = 1024, a_chunk = 9216, b_off = 1024 + 2 * 9216;
    const uint64_t ad0 = make_smem_desc(base + a_off, a_chunk, 128);
    const uint64_t bd0 = make_smem_desc(base + b_off, N * 16, 128);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint64_t adr = ad0 + (uint64_t)((it & 7) * 32);
      const uint32_t aslot = tmem + 256 + (it & 1) * 24;
      if (MODE == 1) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) tmem_cp(aslot + dx * 8, adr + dx);
      }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const uint64_t bd = bd0 + (uint64_t)(tap * N * 2);
        const uint32_t d = tmem + (3 * N <= 256 ? (tap / 3) * N : 0);
        if (MODE == 0) umma_f16(d, adr + (uint64_t)((tap / 3) * 32 + tap % 3), bd, idesc, 1);
        else umma_f16_ts(d, aslot + (tap % 3) * 8, bd, idesc, 1);
      }
    }
    umma_commit(base);
    t1 = clock64();
   }
  }
  mbar_wait(base, 0);
  if (t0 != 0) { t1 = clock64(); out[blockIdx.x] = t1 - t0; }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
template <int N, int MODE> void run(const char* name) {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  const int iters = 4000;
  rate<N, MODE><<<148, 128, 131072>>>(d, iters); cudaDeviceSynchronize();
  rate<N, MODE><<<148, 128, 131072>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  printf("%-28s N=%2d: %6.1f cycles per MMA (ideal tensor %d)  %s\n", name, N, (double)mx / (iters * 9.0), N / 2, cudaGetErrorString(e));
}
int main() {
  run<16, 0>("SS"); run<32, 0>("SS"); run<48, 0>("SS"); run<64, 0>("SS"); run<96, 0>("SS"); run<128, 0>("SS"); run<160, 0>("SS");
  run<192, 0>("SS"); run<256, 0>("SS");
  return 0;
}
