// Weight-gradient GEMM of the 3x3 pad-1 convolution on tcgen05 tensor cores.
//
// Replaces autograd's convolution_backward w.r.t. the weight for EqualizedConv2d (gan.py:29-38) and the
// weight half of _convolution_double_backward that the R1 penalty (gan.py:398-410) triggers.
//
//   dWp[tap][co][ci] = sum_{n,h,w} G[n,h,w,co] * X[n, h+ky-1, w+kx-1, ci]
//
// GEMM view: M = co (128 per CTA), N = ci (one <=64-wide slab), K = pixels.  Both operands are stored
// pixel-major with channels contiguous (NHWC), i.e. they are MN-major for this GEMM; UMMA consumes them
// directly through MN-major shared-memory descriptors, so no transpose pass exists.  A K block is the
// same 128-pixel (bw x bh x bn) TMA box the forward kernel uses: G is loaded un-shifted, X is loaded
// once per tap shifted by the tap offset with TMA zero fill as the padding.  One CTA owns one
// (co tile, ci slab, ky row of 3 taps, K split): the G tile is reused by the 3 taps, the 3 accumulators
// (3 x 64 fp32 columns) live in TMEM, and the split-K partials are reduced with fp32 atomics into dWp.
#include "common.cuh"

#include <stdlib.h>

namespace bg {

namespace {

constexpr int kEpiWarps = 4;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kStages = 2;             // classic mode; the single-tap mode runs 3 stages of 64 KB
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kTapStride = 64;     // TMEM columns reserved per tap accumulator
constexpr uint32_t kARegion = 32768;    // 128 pixels x 128 co x 2 B
constexpr uint32_t kBSlab = 16384;      // 128 pixels x 64 ci x 2 B
constexpr uint32_t kStageBytes = kARegion + 3 * kBSlab;

struct WgradParams {
  int N, H, W, Cin, Cout;
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n;
  int total_kblocks, splits, kblocks_per_split;
  int co_tiles, ci_slabs;
  int co_slab, co_nslabs, ci_slab;
  uint32_t a_row_bytes, b_row_bytes;
  uint32_t a_layout, b_layout;
  uint32_t a_slab_bytes;   // bytes of one co slab (128 rows)
  float* dw;
  // Single-tap, output-stationary mode (small maps, Cin and Cout multiples of 128): a CTA owns ONE tap and a 128 x 128
  // (co x ci) tile and runs the WHOLE pixel range, so there is no split-K, no atomics and no zero fill: 9 * (Cout/128) *
  // (Cin/128) CTAs (144 for the 512 x 512 layers) each stream G (32 KB) + one shifted X tile (2 x 16 KB) per K block —
  // the classic decomposition streams 80 KB per K block from 288 CTAs and then adds 28 MB of fp32 partials atomically
  // (68 us for the 8 x 8 layers whose operands are 4 MB).  stages = 3 here; the result is stored, or added in place
  // when `accumulate` (each element has exactly one owner: bit-reproducible).
  int single_tap, stages, accumulate;
  uint32_t stage_bytes;
};

__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x,
                  const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint8_t* aux = smem + (size_t)p.stages * p.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* done_bar = empty_bar + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work decode
  const int split = blockIdx.x % p.splits;
  const int unit = blockIdx.x / p.splits;
  const int ntg = p.single_tap ? 9 : 3;
  const int tg = unit % ntg;                     // ky, or the tap itself in single-tap mode
  const int cis = (unit / ntg) % p.ci_slabs;
  const int cot = unit / (ntg * p.ci_slabs);
  const int co0 = cot * 128;
  const int ci0 = cis * p.ci_slab;
  const int kb_begin = split * p.kblocks_per_split;
  int kb_end = kb_begin + p.kblocks_per_split;
  if (kb_end > p.total_kblocks) kb_end = p.total_kblocks;
  const int my_kblocks = kb_end > kb_begin ? kb_end - kb_begin : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_g);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
    // only now may the next kernel of the stream become resident: released at kernel entry, a dependent CTA that landed
    // on this SM could take the TMEM columns first and then wait for this grid, which would be waiting for the columns
    pdl_launch_dependents();
  }
  pdl_wait();        // everything above overlaps the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (my_kblocks > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx = (uint32_t)p.co_nslabs * p.a_slab_bytes + (p.single_tap ? 2u : 3u) * (128u * p.b_row_bytes);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          const int tw = kb % p.tiles_w;
          const int th = (kb / p.tiles_w) % p.tiles_h;
          const int tn = kb / (p.tiles_w * p.tiles_h);
          const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
          uint8_t* sb = sa + kARegion;
          mbar_expect_tx(&full_bar[stage], tx);
          for (int s = 0; s < p.co_nslabs; ++s)
            tma_load_4d(&tmap_g, &full_bar[stage], sa + (size_t)s * p.a_slab_bytes, co0 + s * p.co_slab, w0, h0, n0);
          if (p.single_tap) {
            // one tap (ky, kx) = (tg / 3, tg % 3), two 64-channel sub-slabs of the 128-wide ci tile
            for (int s2 = 0; s2 < 2; ++s2)
              tma_load_4d(&tmap_x, &full_bar[stage], sb + (size_t)s2 * kBSlab, ci0 + 64 * s2, w0 + tg % 3 - 1, h0 + tg / 3 - 1, n0);
          } else {
            for (int kx = 0; kx < 3; ++kx)
              tma_load_4d(&tmap_x, &full_bar[stage], sb + (size_t)kx * kBSlab, ci0, w0 + kx - 1, h0 + tg - 1, n0);
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, p.single_tap ? 128 : p.ci_slab, 1, 1);
        const uint32_t a_hi = umma_desc_hi(8u * p.a_row_bytes, p.a_layout);
        const uint32_t b_hi = umma_desc_hi(8u * p.b_row_bytes, p.b_layout);
        const uint32_t a_kstep = (16u * p.a_row_bytes) >> 4;   // 16 pixels per UMMA K step, in 16-byte units
        const uint32_t b_kstep = (16u * p.b_row_bytes) >> 4;
        int stage = 0;
        uint32_t phase = 0;
        uint32_t accum = 0u;
        for (int i = 0; i < my_kblocks; ++i) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
          const uint32_t a_lo = umma_desc_lo(sa, p.a_slab_bytes);
          const uint32_t b_lo = umma_desc_lo(sa + kARegion, kBSlab);
          if (p.single_tap) {
            // N = 128: the two ci sub-slabs are the two MN slabs of the B operand, LBO (= kBSlab) apart
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              tc_mma_bf16_lohi(tmem_base, a_lo + ks * a_kstep, a_hi, b_lo + ks * b_kstep, b_hi, idesc, ks == 0 ? accum : 1u);
          } else {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                tc_mma_bf16_lohi(tmem_base + kx * kTapStride, a_lo + ks * a_kstep, a_hi,
                                 b_lo + kx * (kBSlab >> 4) + ks * b_kstep, b_hi, idesc, ks == 0 ? accum : 1u);
              }
            }
          }
          accum = 1u;
          tc_commit(&empty_bar[stage]);
          if (i == my_kblocks - 1) tc_commit(done_bar);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    } else {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const int co = co0 + row;
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      if (p.single_tap) {
        // the only owner of dw[tg][co0 + row][ci0 .. ci0 + 127]: plain (read-modify-)write, no atomics
        float* drow = p.dw + ((size_t)tg * p.Cout + co) * p.Cin + ci0;
        for (int c = 0; c < 128; c += 16) {
          uint32_t v[16];
          tmem_ld_x16(taddr + c, v);
          tmem_ld_wait();
          if (co < p.Cout) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                     __uint_as_float(v[j + 3]));
              float4* d4 = reinterpret_cast<float4*>(drow + c + j);
              if (p.accumulate) {
                const float4 old = *d4;
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
              }
              *d4 = o;
            }
          }
        }
      }
      for (int kx = 0; kx < (p.single_tap ? 0 : 3); ++kx) {
        const int tap = tg * 3 + kx;
        float* drow = p.dw + ((size_t)tap * p.Cout + co) * p.Cin + ci0;
        for (int c = 0; c < p.ci_slab; c += 16) {
          uint32_t v[16];
          tmem_ld_x16(taddr + kx * kTapStride + c, v);
          tmem_ld_wait();
          if (row < p.co_slab * p.co_nslabs && co < p.Cout) {
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(drow + c + j, __uint_as_float(v[j]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

// x: (N,H,W,Cin) bf16, g: (N,H,W,Cout) bf16, dw: [9][Cout][Cin] fp32 (overwritten, or += if accumulate).
int launch_conv_wgrad(const void* x, const void* g, float* dw, int N, int H, int W, int Cin, int Cout,
                      int accumulate, cudaStream_t stream) {
  BG_REQUIRE(Cin % 16 == 0 && Cout % 16 == 0, "conv_wgrad: channels must be multiples of 16 (Cin %d Cout %d)", Cin,
             Cout);
  BG_REQUIRE(Cin < 64 ? (Cin == 16 || Cin == 32) : Cin % 64 == 0, "conv_wgrad: unsupported Cin %d", Cin);
  BG_REQUIRE(Cout < 64 ? (Cout == 16 || Cout == 32) : Cout % 64 == 0, "conv_wgrad: unsupported Cout %d", Cout);
  BG_REQUIRE(N > 0 && H > 0 && W > 0, "conv_wgrad: empty tensor");
  BG_REQUIRE((W & (W - 1)) == 0 && (H & (H - 1)) == 0, "conv_wgrad: H and W must be powers of two (%d x %d)", H, W);

  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.bw = W < 16 ? W : 16;
  p.bh = H < (128 / p.bw) ? H : (128 / p.bw);
  p.bn = 128 / (p.bw * p.bh);
  p.tiles_w = W / p.bw;
  p.tiles_h = H / p.bh;
  p.tiles_n = (N + p.bn - 1) / p.bn;
  p.total_kblocks = p.tiles_w * p.tiles_h * p.tiles_n;
  p.co_slab = Cout < 64 ? Cout : 64;
  p.co_nslabs = Cout >= 128 ? 2 : 1;
  p.ci_slab = Cin < 64 ? Cin : 64;
  p.co_tiles = (Cout + 127) / 128;
  p.ci_slabs = Cin / p.ci_slab;
  p.a_row_bytes = p.co_slab * 2;
  p.b_row_bytes = p.ci_slab * 2;
  p.a_layout = p.a_row_bytes == 128 ? 2u : (p.a_row_bytes == 64 ? 4u : 6u);
  p.b_layout = p.b_row_bytes == 128 ? 2u : (p.b_row_bytes == 64 ? 4u : 6u);
  p.a_slab_bytes = 128u * p.a_row_bytes;
  p.dw = dw;
  {
    static int st_on = -1;              // BG_WGRAD_SINGLE_TAP=0: always the classic decomposition
    if (st_on < 0) { const char* e = getenv("BG_WGRAD_SINGLE_TAP"); st_on = (e && e[0] == '0') ? 0 : 1; }
    p.single_tap = (st_on && Cin % 128 == 0 && Cout % 128 == 0 && p.total_kblocks <= 32) ? 1 : 0;
  }
  p.accumulate = accumulate ? 1 : 0;
  p.stages = p.single_tap ? 3 : kStages;
  p.stage_bytes = p.single_tap ? (kARegion + 2 * kBSlab) : kStageBytes;
  if (p.single_tap) {
    p.ci_slab = 128;
    p.ci_slabs = Cin / 128;
  }
  const int units = p.co_tiles * p.ci_slabs * (p.single_tap ? 9 : 3);
  int splits = (2 * num_sms() + units - 1) / units;
  if (p.single_tap) splits = 1;
  if (splits < 1) splits = 1;
  if (splits > p.total_kblocks) splits = p.total_kblocks;
  p.kblocks_per_split = (p.total_kblocks + splits - 1) / splits;
  splits = (p.total_kblocks + p.kblocks_per_split - 1) / p.kblocks_per_split;
  p.splits = splits;

  CUtensorMap tmg, tmx;
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
    uint32_t box[4] = {(uint32_t)p.co_slab, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    if (make_tmap_bf16(&tmg, g, 4, dims, str, box, (int)p.a_row_bytes) != 0) return 1;
  }
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {(uint32_t)(p.b_row_bytes / 2), (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};   // one <= 64-wide sub-slab
    if (make_tmap_bf16(&tmx, x, 4, dims, str, box, (int)p.b_row_bytes) != 0) return 1;
  }

  if (!accumulate && !p.single_tap) if (launch_zero(dw, (size_t)9 * Cout * Cin * sizeof(float), stream) != 0) return 1;
  const size_t smem_bytes = (size_t)p.stages * p.stage_bytes + 256 + 1024;
  BG_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  BG_CHECK_CUDA(launch_pdl(conv_wgrad_kernel, units * splits, kThreads, smem_bytes, stream, tmg, tmx, p));
  return 0;
}

}  // namespace bg
