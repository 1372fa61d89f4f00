// Micro-benchmark of the planned "strip" pipeline: 4 loader warps copy the A operand smem -> registers -> TMEM
// (3 dx-shifted copies per k-step via 3 LDS.128 pairs + tcgen05.st), one thread issues 9 MMAs (A from TMEM) per k-step.
#include <cstdio>
#include <cuda_fp16.h>
#include "../../explorable-super-resolution_b200/csrc/ptx.cuh"
using namespace esr;
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4& a, const uint4& b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
constexpr int SLOTS = 8;
template <int N, int LW>
__global__ void __launch_bounds__(32 * LW + 32, 1) rate(long long* out, int iters, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* sp = smem_raw + (base - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < 60000 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sp + 1024)[i] = 0x3c003c00u;
  const uint32_t bar_full = base, bar_empty = base + 8 * SLOTS, bar_done = base + 16 * SLOTS;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SLOTS; ++s) { mbar_init(bar_full + 8 * s, 4); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_done, 1); fence_barrier_init();
  }
  fence_proxy_async();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == LW) { tmem_alloc(base + 512, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(sp + 512);
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t a_off = 1024, plane = 132 * 16 * 8, b_off = 1024 + 2 * plane;   // 8 rows of 132 px per plane
  long long t0 = clock64();
  if (warp < LW) {            // loaders: lane quarter = warp & 3; the LW/4 warps of a quarter take iterations round-robin
    const int wq = warp & 3;
    for (int it = warp >> 2; it < iters; it += LW / 4) {
      const int s = it % SLOTS; const uint32_t ph = (it / SLOTS) & 1;
      mbar_wait(bar_empty + 8 * s, ph ^ 1u);
      tc_fence_after();
      const uint32_t rowa = base + a_off + (it & 7) * 132 * 16 + (wq * 32 + lane) * 16;
      uint4 v[3][2];
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[dx][0].x), "=r"(v[dx][0].y), "=r"(v[dx][0].z), "=r"(v[dx][0].w) : "r"(rowa + dx * 16));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[dx][1].x), "=r"(v[dx][1].y), "=r"(v[dx][1].z), "=r"(v[dx][1].w) : "r"(rowa + plane + dx * 16));
      }
      const uint32_t ta = tmem + ((uint32_t)(wq * 32) << 16) + 256 + s * 24;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) tmem_st8(ta + dx * 8, v[dx][0], v[dx][1]);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * s);
    }
  } else if (warp == LW) {    // MMA issuer
    const uint64_t bd0 = make_smem_desc(base + b_off, N * 16, 128);
    for (int it = 0; it < iters; ++it) {
      const int s = it % SLOTS; const uint32_t ph = (it / SLOTS) & 1;
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t aslot = tmem + 256 + s * 24;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
          umma_f16_ts(tmem + (tap / 3) * N, aslot + (tap % 3) * 8, bd0 + (uint64_t)(tap * N * 2), idesc, 1);
        umma_commit(bar_empty + 8 * s);
        if (it == iters - 1) umma_commit(bar_done);
      }
      __syncwarp();
    }
  }
  mbar_wait(bar_done, 0);
  long long t1 = clock64();
  if (threadIdx.x == 32 * LW) out[blockIdx.x] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == LW) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
template <int N, int LW> void run() {
  long long* d; cudaMalloc(&d, 148 * 8); float* sink; cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(rate<N, LW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  const int iters = 4000;
  rate<N, LW><<<148, 32 * LW + 32, 100000>>>(d, iters, sink); cudaDeviceSynchronize();
  rate<N, LW><<<148, 32 * LW + 32, 100000>>>(d, iters, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  printf("LDS->STTM loaders (%d warps) + TS MMA   N=%2d: %6.1f cycles per MMA (ideal tensor %d)  %s\n", LW, N, (double)mx / (iters * 9.0), N / 2, cudaGetErrorString(e));
}
int main() { run<32, 4>(); run<32, 8>(); run<32, 12>(); run<32, 16>(); run<64, 8>(); run<64, 12>(); run<16, 12>(); return 0; }
