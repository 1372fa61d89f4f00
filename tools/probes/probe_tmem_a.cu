// Probe: does  tcgen05.cp.128x256b (smem -> TMEM)  +  tcgen05.mma with the A operand in TMEM  reproduce the SS result?
// One CTA, M=128, N=32, K=16 (one MMA), fp16 operands in the no-swizzle K-major canonical layout.
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "../../explorable-super-resolution_b200/csrc/ptx.cuh"
using namespace esr;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

__global__ void probe(const __half* A, const __half* B, float* D_ss, float* D_ts, int N, int shift_rows) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* sp = smem_raw + (base - smem_u32(smem_raw));
  // A: [2 k-chunks][160 rows][8 halfs] (row pitch 16 B, chunk stride 160*16); B: [2][N][8]
  const uint32_t a_off = 1024, a_chunk = 160 * 16, b_off = 1024 + 2 * a_chunk, bar_off = 0, slot_off = 64;
  for (int i = threadIdx.x; i < 160 * 16; i += blockDim.x) {  // A given as [160][16] row-major
    int r = i / 16, k = i % 16;
    *reinterpret_cast<__half*>(sp + a_off + (k / 8) * a_chunk + r * 16 + (k % 8) * 2) = A[i];
  }
  for (int i = threadIdx.x; i < N * 16; i += blockDim.x) {
    int r = i / 16, k = i % 16;
    *reinterpret_cast<__half*>(sp + b_off + (k / 8) * (N * 16) + r * 16 + (k % 8) * 2) = B[i];
  }
  if (threadIdx.x == 0) { mbar_init(base + bar_off, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (threadIdx.x < 32) { tmem_alloc(base + slot_off, 128); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(sp + slot_off);
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (threadIdx.x == 0) {
    const uint64_t ad = make_smem_desc(base + a_off + shift_rows * 16, a_chunk, 128);
    const uint64_t bd = make_smem_desc(base + b_off, N * 16, 128);
    umma_f16(tmem + 0, ad, bd, idesc, 0);            // SS -> columns [0, N)
    tmem_cp_128x256b(tmem + 96, ad);                 // A tile -> columns [96, 104)
    umma_f16_ts(tmem + 32, tmem + 96, bd, idesc, 0); // TS -> columns [32, 32+N)
    umma_commit(base + bar_off);
  }
  mbar_wait(base + bar_off, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 4) {
    for (int cb = 0; cb < N; cb += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + cb, r); tc_wait_ld();
      for (int k = 0; k < 16; ++k) D_ss[(warp * 32 + lane) * N + cb + k] = __uint_as_float(r[k]);
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + 32 + cb, r); tc_wait_ld();
      for (int k = 0; k < 16; ++k) D_ts[(warp * 32 + lane) * N + cb + k] = __uint_as_float(r[k]);
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

int main() {
  const int N = 32;
  __half hA[160 * 16], hB[N * 16];
  srand(1);
  for (auto& v : hA) v = __float2half((rand() % 17 - 8) / 8.f);
  for (auto& v : hB) v = __float2half((rand() % 13 - 6) / 4.f);
  __half *dA, *dB; float *dss, *dts;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dss, 128 * N * 4); cudaMalloc(&dts, 128 * N * 4);
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  for (int shift : {0, 1, 2, 33}) {
    cudaMemset(dss, 0, 128 * N * 4); cudaMemset(dts, 0, 128 * N * 4);
    probe<<<1, 128, 16384>>>(dA, dB, dss, dts, N, shift);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shift %d: CUDA error %s\n", shift, cudaGetErrorString(e)); return 1; }
    static float ss[128 * N], ts[128 * N];
    cudaMemcpy(ss, dss, sizeof(ss), cudaMemcpyDeviceToHost); cudaMemcpy(ts, dts, sizeof(ts), cudaMemcpyDeviceToHost);
    double e_ss = 0, e_ts = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
      float ref = 0; for (int k = 0; k < 16; ++k) ref += __half2float(hA[(m + shift) * 16 + k]) * __half2float(hB[n * 16 + k]);
      e_ss = fmax(e_ss, fabs(ss[m * N + n] - ref)); e_ts = fmax(e_ts, fabs(ts[m * N + n] - ref));
    }
    printf("shift %2d: max err SS %.3g   TS(A via tcgen05.cp) %.3g   sample ts[5][3]=%g ss[5][3]=%g\n", shift, e_ss, e_ts, ts[5 * N + 3], ss[5 * N + 3]);
  }
  return 0;
}
