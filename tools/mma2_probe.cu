// Issue-rate probe for tcgen05.mma kind::f16 (bf16 -> fp32) with shared-memory operands: cycles per instruction for
//   cta_group::1, M = 128   (what conv_halo.cu / conv_wgrad_halo.cu issue today) and
//   cta_group::2, M = 256   (one instruction for a CTA pair: the planned next step)
// at N = 16 ... 256, K = 16, all SMs busy.  Operands are zero-filled swizzle-128B K-major tiles; only timing matters.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I byo-gan_b200/csrc -o tools/_build/mma2_probe tools/mma2_probe.cu
//   timeout 60 tools/_build/mma2_probe
#include "common.cuh"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

using namespace bg;

namespace {

constexpr int kThreads = 160;           // warp 0 allocates TMEM, warps 1..4 can issue
constexpr int kMaxIssuers = 4;
constexpr uint32_t kABytes = 44 * 1024;   // room for an 18 x 18 halo of 128-byte pixel rows (conv_halo.cu's A stage)
constexpr uint32_t kBBytes = 256 * 128;   // up to 256 rows x 64 bf16

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

template <int CTAS>
__global__ void __launch_bounds__(kThreads, 1) probe_kernel(int N, int iters, int halo, int issuers, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* a_tile = smem;
  uint8_t* b_tile = smem + kABytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kABytes + kBBytes);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + kMaxIssuers);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = CTAS == 2 ? cluster_rank() : 0u;

  for (uint32_t i = threadIdx.x; i < (kABytes + kBBytes) / 16; i += kThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxIssuers; ++i) mbar_init(bar + i, 1);
    fence_barrier_init();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy zero fill -> visible to the MMA unit
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();
  if (warp == 0) {
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *slot;

  // issuer w (lane 0 of warp 1 + w) owns accumulator columns [w * N, (w + 1) * N) and barrier w: `issuers` threads feed
  // the tensor pipe concurrently (conv_halo.cu issues everything from ONE thread today)
  const int w = warp - 1;
  if (w >= 0 && w < issuers && (threadIdx.x & 31) == 0) {
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(CTAS == 2 ? 256 : 128, N, 0, 0);
      // halo = 0: one aligned 128-row tile (SBO 1024).  halo = 1: conv_halo.cu's operand: tap (ky, kx) is the 18-pixel-pitch
      // halo tile viewed from pixel row ky * 18 + kx (SBO = one halo row = 2304 bytes), the 9 taps unrolled like the kernel
      const bool views = (halo & 1) != 0;
      const uint32_t alt = (halo & 2) ? (uint32_t)N : 0u;      // 2: alternate two accumulators from ONE thread (what conv_halo.cu does)
      const uint64_t adesc0 = umma_desc(smem_u32(a_tile), 0, views ? 18 * 128 : 1024, 2);
      const uint64_t bdesc = umma_desc(smem_u32(b_tile), 0, 1024, 2);
      const uint32_t d_tmem = tmem_base + (uint32_t)(w * N);
      const long long t0 = clock64();
      for (int it = 0; it < iters; it += 9) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint64_t adesc = adesc0 + (uint64_t)(views ? (((tap / 3) * 18 + tap % 3) * 128) >> 4 : 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t acc = (it | tap | (k >> 1)) != 0 ? 1u : 0u;
            if (CTAS == 2) {
              asm volatile(
                  "{\n\t"
                  ".reg .pred p;\n\t"
                  "setp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
                  "}\n" ::"r"(d_tmem),
                  "l"(adesc + 2u * k), "l"(bdesc + 2u * k), "r"(idesc), "r"(acc)
                  : "memory");
            } else {
              tc_mma_bf16(d_tmem + (k & 1) * alt, adesc + 2u * k, bdesc + 2u * k, idesc, acc);
            }
          }
        }
      }
      if (CTAS == 2) {
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                         "r"(smem_u32(bar + w)), "h"((uint16_t)3)
                     : "memory");
      } else {
        tc_commit(bar + w);
      }
      mbar_wait(bar + w, 0);
      const long long t1 = clock64();
      out[(blockIdx.x / CTAS) * kMaxIssuers + w] = (unsigned long long)(t1 - t0);
    } else {
      mbar_wait(bar + w, 0);   // the peer CTA must not tear down before the pair's MMAs have retired
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (CTAS == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      tmem_dealloc(tmem_base, 512);
  }
}

template <int CTAS>
double run(int N, int iters, int halo, int issuers, int sms, unsigned long long* d_out) {
  const size_t smem = kABytes + kBBytes + 1024 + 128;
  cudaFuncSetAttribute(probe_kernel<CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((sms / CTAS) * CTAS);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const int groups = sms / CTAS;
  cudaMemset(d_out, 0, sizeof(unsigned long long) * 1024);
  for (int rep = 0; rep < 2; ++rep) {   // first launch warms up
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe_kernel<CTAS>, N, iters, halo, issuers, d_out);
    if (e != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) {
      printf("launch failed (CTAS %d N %d): %s\n", CTAS, N, cudaGetErrorString(e));
      exit(1);
    }
  }
  unsigned long long h[1024];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  unsigned long long mx = 0;
  for (int i = 0; i < groups * kMaxIssuers; ++i) mx = h[i] > mx ? h[i] : mx;
  return (double)mx / (4.0 * iters * issuers);       // clk per instruction retired by the SM (pair)
}

}  // namespace

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* d_out = nullptr;
  cudaMalloc(&d_out, sizeof(unsigned long long) * 1024);
  const int iters = 1998;   // multiple of the 9 unrolled taps
  printf("tcgen05.mma kind::f16 bf16, K=16, smem operands, %d SMs busy, %d instructions per issuer\n", sms, 4 * iters);
  printf("clk per instruction retired (per SM, or per SM pair for cta_group::2)\n");
  printf("%5s | %9s %9s %9s | %9s %9s | %9s %9s | %s\n", "N", "1cta x1", "1cta x2", "1cta x4", "x1 alt2", "halo alt2", "2cta x1",
         "2cta x2", "TFLOP/s @1.965 GHz: 1cta x1 / best 1cta / best 2cta");
  const int ns[] = {16, 32, 64, 96, 128, 256};
  for (int N : ns) {
    const int max_iss = 512 / N < kMaxIssuers ? 512 / N : kMaxIssuers;
    const double a1 = run<1>(N, iters, 0, 1, sms, d_out);
    const double a2 = run<1>(N, iters, 0, 2, sms, d_out);
    const double a4 = max_iss >= 4 ? run<1>(N, iters, 0, 4, sms, d_out) : 0.0;
    const double h1 = 2 * N <= 512 ? run<1>(N, iters, 2, 1, sms, d_out) : 0.0;   // one thread, two accumulators
    const double h2 = 2 * N <= 512 ? run<1>(N, iters, 3, 1, sms, d_out) : 0.0;   // same, on the shifted halo views
    const double p1 = run<2>(N, iters, 0, 1, sms, d_out);
    const double p2 = run<2>(N, iters, 0, 2, sms, d_out);
    double b1 = a1 < a2 ? a1 : a2;
    if (a4 > 0 && a4 < b1) b1 = a4;
    const double b2 = p1 < p2 ? p1 : p2;
    const double k1 = 2.0 * 128 * N * 16 * 1.965e9 * sms / 1e12, k2 = 2.0 * 256 * N * 16 * 1.965e9 * (sms / 2) / 1e12;
    printf("%5d | %9.1f %9.1f %9.1f | %9.1f %9.1f | %9.1f %9.1f | %.0f / %.0f / %.0f\n", N, a1, a2, a4, h1, h2, p1, p2, k1 / a1,
           k1 / b1, k2 / b2);
  }
  cudaFree(d_out);
  return 0;
}
