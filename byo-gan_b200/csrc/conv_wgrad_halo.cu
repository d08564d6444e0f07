// Weight-gradient GEMM of the 3x3 pad-1 convolution for maps >= 16x16: shifted-view variant of conv_wgrad.cu.
//
//   dWp[tap][co][ci] = sum_{n,h,w} G[n,h,w,co] * X[n, h+ky-1, w+kx-1, ci]
//
// Same GEMM view as conv_wgrad.cu (M = co, N = ci, K = pixels, both operands MN-major straight from NHWC), but
// the three kx taps of a kernel row no longer get three separately loaded X tiles.  A K block is a 16x8 pixel
// tile; X is loaded ONCE per K block as an (16+2) x 8 halo box (TMA, zero fill = padding) and tap kx of image
// row r is just the same shared-memory tile read from pixel row  r*18 + kx  on: UMMA's swizzle is a function of
// the absolute shared-memory address, so a descriptor whose start is shifted by whole 128/64/32-byte pixel rows
// stays consistent with what TMA wrote (verified on B200 against torch for all three swizzle widths).
// That cuts the L2->SMEM fill per K block from G + 3 X tiles to G + 1.125 X tiles, which was the binding limit
// (profiles/r1_ncu_full_conv_wgrad.csv).  The ci slab grows to 128 (two 64-wide sub-slabs, LBO apart) so the
// N=128 MMAs run at full tensor rate; 3 tap accumulators x 128 columns live in TMEM; 3-stage TMA pipeline; the
// issue loop is warp-uniform (descriptors in uniform registers); split-K partials are reduced with vector fp32
// reductions (red.global.add.v4.f32).
//
// Round 2: (1) CTAs are numbered with the unit index fastest and the grid is one resident wave, so the CTAs that share a
// K range run together and find G / X in L2 (DRAM read 10.2 -> 5.84 GB per iteration).  (2) For Cout <= 64 the M = 128
// rows of the instruction carry two or three KERNEL ROWS: the operands are MN-major, the M-atoms of one instruction are
// LBO bytes apart, and with LBO = one tile row atom a reads the G tile a image rows further down (see ky_stack below):
// 256^2 64->32 204 -> 84 us.
#include "common.cuh"

#include <stdlib.h>

namespace bg {

namespace {

constexpr int kEpiWarps = 4;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kStages = 3;
constexpr int kBw = 16, kBh = 8;                 // K block = 16 x 8 pixels of one image
constexpr int kHaloW = kBw + 2;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTapStride = 128;             // TMEM columns reserved per tap accumulator
constexpr uint32_t kARegion = 32768;             // 128 pixels x (2 x 64 co) x 2 B
constexpr uint32_t kBSub = 18432;                // 18 x 8 pixels x 64 ci x 2 B (one 64-wide ci sub-slab)
constexpr uint32_t kStageBytes = kARegion + 2 * kBSub;

struct WgradHaloParams {
  int N, H, W, Cin, Cout;
  int tiles_w, tiles_h;
  int total_kblocks, splits, kblocks_per_split, units;
  int co_tiles, ci_slabs;
  int co_slab, co_nslabs;          // G is loaded as co_nslabs boxes of co_slab channels (<= 64 each)
  int ci_sub, ci_nsub, ci_slab;    // X slab = ci_nsub sub-slabs of ci_sub channels (<= 64 each)
  uint32_t a_row_bytes, b_row_bytes;
  uint32_t a_layout, b_layout;
  uint32_t a_slab_bytes;
  int stack_taps;                  // 1: the three kx taps are ONE MMA (N = 3 * ci_sub, N-atoms one pixel apart)
  // ky_stack (Cout <= 64, plain 3x3): the M = 128 rows of the instruction, of which Cout <= 64 are channels, carry
  // SEVERAL kernel rows: M-atom a (co_slab channels) reads the G tile a image rows further down (LBO = one 16-pixel tile
  // row), the X halo stays at row offset 0, so atom a holds kernel row ky = 2 - abase - a.  The G box starts one image row
  // above the K block and has 10 rows.  Every (G row, X row) pair a shifted atom misses at the top / bottom of the map
  // has its X row outside the image, i.e. contributes zero.  Units per (ci slab): 2 for Cout = 64 (ky {2,1}, ky {0}),
  // 1 for Cout <= 32 (ky {2,1,0}).
  int ky_stack, ky_units;
  // ky_dual (Cout = 64, taps stacked along N): the third kernel row comes from a SECOND MMA per K step on the same
  // shared-memory tiles (A start two tile rows further down, its own TMEM columns) instead of from a second CTA that would
  // fetch G and X again: the narrow layers had become L2 -> shared-memory bound (29 KB per 448 clk and SM).
  int ky_dual;
  // pool4: weight gradient of conv3x3 -> AvgPool2d(2) taken on the 4x4 stride-2 form: G is the POOLED gradient (H, W
  // below are its dims), X the full-resolution conv input (2H x 2W).  A CTA owns one tap row a (0..3); per K block it
  // loads two column-parity tiles of X (TMA boxes with element stride 2 along W and H): parity 1 serves taps b = 0, 2
  // (pixel shifts 0, 1), parity 0 serves b = 1, 3 — each pair is one N-stacked MMA.  dw is then [16][Cout][Cin].
  int pool4;
  float* dw;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x,
                       const WgradHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint8_t* aux = smem + (size_t)kStages * kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* done_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  // work decode: blockIdx -> (K split, co tile, ci slab, kernel row), UNIT FASTEST: the CTAs that share a K range (the
  // three kernel rows and all co / ci tiles of it) are neighbours in the launch order and run at the same time, so G
  // and the X halos they all need come out of L2 after the first of them has touched them.  (With the split index
  // fastest — round 1 — kernel row 0 streamed the whole map in the first wave and row 2 in the last: every operand was
  // read from DRAM three times on the 256x256 layers, 1.83x over the whole family, profiles/r1_ncu_dram_*.)
  const int unit = blockIdx.x % p.units;
  const int split = blockIdx.x / p.units;
  const int ntg = p.ky_stack ? p.ky_units : (p.pool4 ? 4 : 3);
  const int tg = unit % ntg;                     // ky (pool4: tap row a; ky_stack: index of the kernel-row group)
  // ky_stack: first G tile row of M-atom 0, relative to the box (which starts one image row above the K block).
  // pool4 + ky_stack (Cout = 64): pooled row i of G meets X row 2i + s, so an atom shifted by sigma pooled rows holds tap
  // row a = s - 2 sigma + 1: group 0 (s = 0, atoms sigma = -1, 0) -> a = 3, 1; group 1 (s = 1, sigma = 0, +1) -> a = 2, 0.
  const int abase = p.ky_stack ? (p.pool4 ? tg : 2 * tg) : 0;
  const int xrow = p.pool4 ? (p.ky_stack ? tg : tg - 1) : (p.ky_stack ? 0 : tg - 1);     // X row offset of the K block
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
    for (int s = 0; s < kStages; ++s) {
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
      // ------------------------------ TMA producer ------------------------------
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx = (p.ky_stack ? (uint32_t)(kBw * (kBh + 2)) * p.a_row_bytes : (uint32_t)p.co_nslabs * p.a_slab_bytes) +
                            (p.pool4 ? 2u * (uint32_t)((kBw + 1) * kBh) * p.b_row_bytes
                                     : (uint32_t)p.ci_nsub * (uint32_t)(kHaloW * kBh) * p.b_row_bytes);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          const int tw = kb % p.tiles_w;
          const int th = (kb / p.tiles_w) % p.tiles_h;
          const int n = kb / (p.tiles_w * p.tiles_h);
          const int w0 = tw * kBw, h0 = th * kBh;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + (size_t)stage * kStageBytes;
          uint8_t* sb = sa + kARegion;
          mbar_expect_tx(&full_bar[stage], tx);
          if (p.ky_stack) tma_load_4d(&tmap_g, &full_bar[stage], sa, co0, w0, h0 - 1, n);     // 10 rows from h0 - 1
          else
            for (int s = 0; s < p.co_nslabs; ++s)
              tma_load_4d(&tmap_g, &full_bar[stage], sa + (size_t)s * p.a_slab_bytes, co0 + s * p.co_slab, w0, h0, n);
          if (p.pool4) {
            // sub-region 0: odd full-resolution columns 2j-1 (j = w0..w0+16), sub-region 1: even columns 2j; rows
            // 2i + a - 1 for the 8 pooled rows i of the block (both with element stride 2)
            tma_load_4d(&tmap_x, &full_bar[stage], sb, ci0, 2 * w0 - 1, 2 * h0 + xrow, n);
            tma_load_4d(&tmap_x, &full_bar[stage], sb + kBSub, ci0, 2 * w0, 2 * h0 + xrow, n);
          } else {
            for (int s = 0; s < p.ci_nsub; ++s)
              tma_load_4d(&tmap_x, &full_bar[stage], sb + (size_t)s * kBSub, ci0 + s * p.ci_sub, w0 - 1, h0 + xrow, n);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    } else if (warp == 1) {
      // ------------------------------ MMA issuer (whole warp converged, one elected lane issues) ----------------
      // stack_taps (Cin <= 64): tap kx of X is the same tile one pixel row (b_row_bytes) further on, so the three
      // taps are three N-atoms with LBO = one pixel row and run as ONE MMA with N = 3 * ci: a small-N tcgen05.mma
      // costs ~60 clk no matter how narrow it is, so fewer, wider instructions are what speeds the narrow layers up.
      const uint32_t idesc = umma_idesc_bf16(128, p.stack_taps ? 3 * p.ci_slab : p.ci_slab, 1, 1);
      const uint64_t a_desc0 =
          umma_desc(smem_u32(smem) + (uint32_t)abase * (uint32_t)kBw * p.a_row_bytes,
                    p.ky_stack ? (uint32_t)kBw * p.a_row_bytes : p.a_slab_bytes, 8u * p.a_row_bytes, p.a_layout);
      const uint64_t b_desc0 = umma_desc(smem_u32(smem) + kARegion, (p.stack_taps || p.pool4) ? p.b_row_bytes : kBSub,
                                         8u * p.b_row_bytes, p.b_layout);
      const uint32_t a_kstep = (16u * p.a_row_bytes) >> 4;          // 16 pixels (one image row of the tile) per K step
      const uint32_t b_kstep = ((uint32_t)kHaloW * p.b_row_bytes) >> 4;   // ... which is 18 halo pixels further in X
      const uint32_t b_tap = p.b_row_bytes >> 4;                     // one pixel to the right = next tap
      const uint32_t idesc4 = umma_idesc_bf16(128, 2 * p.ci_slab, 1, 1);      // pool4: two taps per MMA
      const uint32_t b_kstep4 = ((uint32_t)(kBw + 1) * p.b_row_bytes) >> 4;   // pool4: 17-pixel tile rows
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accum = 0u;
      for (int i = 0; i < my_kblocks; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader) {
          const uint64_t a_st = a_desc0 + (uint64_t)((uint32_t)stage * (kStageBytes >> 4));
          const uint64_t b_st = b_desc0 + (uint64_t)((uint32_t)stage * (kStageBytes >> 4));
          if (p.pool4) {
            // two N-stacked MMAs per K step: taps (b=0, b=2) from the odd-column tile, (b=1, b=3) from the even one
#pragma unroll
            for (int ks = 0; ks < kBh; ++ks) {
#pragma unroll
              for (int pc = 0; pc < 2; ++pc)
                tc_mma_bf16(tmem_base + (uint32_t)pc * 2u * (uint32_t)p.ci_slab, a_st + (uint64_t)(ks * a_kstep),
                            b_st + (uint64_t)(ks * b_kstep4 + pc * (kBSub >> 4)), idesc4, ks == 0 ? accum : 1u);
            }
          } else if (p.stack_taps) {
#pragma unroll
            for (int ks = 0; ks < kBh; ++ks) {
              tc_mma_bf16(tmem_base, a_st + (uint64_t)(ks * a_kstep), b_st + (uint64_t)(ks * b_kstep), idesc,
                          ks == 0 ? accum : 1u);
              if (p.ky_dual)
                tc_mma_bf16(tmem_base + 3u * (uint32_t)p.ci_slab, a_st + (uint64_t)((ks + 2) * a_kstep),
                            b_st + (uint64_t)(ks * b_kstep), idesc, ks == 0 ? accum : 1u);
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < kBh; ++ks) {
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                tc_mma_bf16(tmem_base + kx * kTapStride, a_st + (uint64_t)(ks * a_kstep),
                            b_st + (uint64_t)(ks * b_kstep + kx * b_tap), idesc, ks == 0 ? accum : 1u);
              }
            }
          }
          tc_commit(&empty_bar[stage]);
          if (i == my_kblocks - 1) tc_commit(done_bar);
        }
        accum = 1u;
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    } else {
      // ------------------------------ epilogue: TMEM -> vector reductions into dWp ------------------------------
      const int q = warp & 3;
      const int row = q * 32 + lane;
      // ky_stack: row = (M-atom a, channel); atom a holds kernel row 2 - abase - a (dead when that is not 0..2)
      const int atom = p.ky_stack ? row / p.co_slab : 0;
      const int co = p.ky_stack ? co0 + row - atom * p.co_slab : co0 + row;
      mbar_wait(done_bar, 0);
      tc_fence_after();
      for (int dual = 0; dual < (p.ky_dual ? 2 : 1); ++dual) {
      const int ab = abase + 2 * dual;               // ky_dual: the second accumulator started two tile rows further down
      // kernel row of this accumulator row (pool4: tap row of the 4x4 form)
      const int ky = p.ky_stack ? (p.pool4 ? xrow - 2 * (ab + atom - 1) + 1 : 2 - ab - atom) : tg;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)dual * 3u * (uint32_t)p.ci_slab;
      const bool live = p.ky_stack ? (ky >= 0 && ky <= (p.pool4 ? 3 : 2) && co < p.Cout)
                                   : (row < p.co_slab * p.co_nslabs && co < p.Cout);
      const int ntaps = p.pool4 ? 4 : 3;
      for (int kx = 0; kx < ntaps; ++kx) {
        // pool4: TMEM holds the taps in the order b = 0, 2, 1, 3 (ci_slab columns each)
        const int tap = p.pool4 ? ky * 4 + ((kx & 1) * 2 + (kx >> 1)) : ky * 3 + kx;
        float* drow = p.dw + ((size_t)tap * p.Cout + co) * p.Cin + ci0;
        for (int c = 0; c < p.ci_slab; c += 16) {
          uint32_t v[16];
          tmem_ld_x16(taddr + kx * ((p.stack_taps || p.pool4) ? (uint32_t)p.ci_slab : kTapStride) + c, v);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              red_add_v4(drow + c + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                         __uint_as_float(v[j + 3]));
          }
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

bool conv_wgrad_halo_supported(int N, int H, int W, int Cin, int Cout) {
  const bool cin_ok = Cin < 64 ? (Cin == 16 || Cin == 32) : Cin % 64 == 0;
  const bool cout_ok = Cout < 64 ? (Cout == 16 || Cout == 32) : Cout % 64 == 0;
  return N > 0 && W >= 16 && H >= 8 && W % kBw == 0 && H % kBh == 0 && cin_ok && cout_ok;
}

// x: (N,H,W,Cin) bf16, g: (N,H,W,Cout) bf16, dw: [9][Cout][Cin] fp32 (overwritten, or += if accumulate).
// pool4: g is the POOLED gradient (N,H,W,Cout) of conv3x3 -> AvgPool2d(2), x the conv input (N,2H,2W,Cin), dw the
// 16-tap gradient [16][Cout][Cin] of the equivalent 4x4 stride-2 kernel (bg_unpack_wgrad_pool4 folds it back to 3x3).
int launch_conv_wgrad_halo(const void* x, const void* g, float* dw, int N, int H, int W, int Cin, int Cout,
                           int accumulate, int pool4, cudaStream_t stream) {
  BG_REQUIRE(conv_wgrad_halo_supported(N, H, W, Cin, Cout), "conv_wgrad_halo: unsupported shape N %d H %d W %d Cin %d Cout %d",
             N, H, W, Cin, Cout);
  WgradHaloParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.tiles_w = W / kBw;
  p.tiles_h = H / kBh;
  p.total_kblocks = p.tiles_w * p.tiles_h * N;
  p.co_slab = Cout < 64 ? Cout : 64;
  p.co_nslabs = Cout >= 128 ? 2 : 1;
  p.co_tiles = (Cout + 127) / 128;
  p.ci_sub = Cin < 64 ? Cin : 64;
  p.ci_nsub = (Cin >= 128 && !pool4) ? 2 : 1;      // pool4 uses the two sub-regions for the two column parities
  p.pool4 = pool4 ? 1 : 0;
  p.ci_slab = p.ci_sub * p.ci_nsub;
  p.ci_slabs = Cin / p.ci_slab;
  p.a_row_bytes = p.co_slab * 2;
  p.b_row_bytes = p.ci_sub * 2;
  p.a_layout = p.a_row_bytes == 128 ? 2u : (p.a_row_bytes == 64 ? 4u : 6u);
  p.b_layout = p.b_row_bytes == 128 ? 2u : (p.b_row_bytes == 64 ? 4u : 6u);
  p.a_slab_bytes = 128u * p.a_row_bytes;
  p.dw = dw;
  {
    static int mode = -1;     // BG_WGRAD_STACK: 0 never, 1 Cin <= 32 only, 2 (default) whenever Cin <= 64
    if (mode < 0) { const char* e = getenv("BG_WGRAD_STACK"); mode = e ? atoi(e) : 2; }
    p.stack_taps = (p.ci_nsub == 1 && (mode == 2 || (mode == 1 && Cin <= 32))) ? 1 : 0;
  }
  if (p.pool4) p.stack_taps = 0;
  {
    static int ky_on = -1;    // BG_WGRAD_KYSTACK=0: one kernel row per CTA also for Cout <= 64 (A/B switch)
    if (ky_on < 0) { const char* e = getenv("BG_WGRAD_KYSTACK"); ky_on = (e && e[0] == '0') ? 0 : 1; }
    p.ky_stack = (ky_on && (p.pool4 ? Cout == 64 : Cout <= 64)) ? 1 : 0;
    p.ky_units = Cout == 64 ? 2 : 1;
    // BG_WGRAD_KYDUAL=1 turns it on.  Off by default: measured 256^2 32->64 130 -> 96 us, 64->64 182 -> 151 us and all
    // kernel tests pass, but one of three full model-suite runs with it had a style-mixing gradient test under its cosine
    // threshold (not reproduced in isolation; the suite never failed without it) and there was no GPU time left to chase it.
    static int dual_on = -1;
    if (dual_on < 0) { const char* e = getenv("BG_WGRAD_KYDUAL"); dual_on = (e && e[0] == '1') ? 1 : 0; }
    p.ky_dual = (dual_on && p.ky_stack && !p.pool4 && Cout == 64 && p.stack_taps) ? 1 : 0;
    if (p.ky_dual) p.ky_units = 1;
  }
  const int units = p.co_tiles * p.ci_slabs * (p.ky_stack ? p.ky_units : (p.pool4 ? 4 : 3));
  p.units = units;
  // one resident wave (1 CTA per SM: 3 x 68 KB stages): splits = floor(SMs / units), so that every CTA of the grid runs
  // concurrently with the others that read the same K range.  BG_WGRAD_WAVES=2 restores round 1's two waves.
  static int waves = 0;
  if (waves == 0) { const char* e = getenv("BG_WGRAD_WAVES"); waves = (e && atoi(e) > 0) ? atoi(e) : 1; }
  int splits = waves * num_sms() / units;
  if (splits < 1) splits = 1;
  if (splits > p.total_kblocks) splits = p.total_kblocks;
  p.kblocks_per_split = (p.total_kblocks + splits - 1) / splits;
  splits = (p.total_kblocks + p.kblocks_per_split - 1) / p.kblocks_per_split;
  p.splits = splits;

  CUtensorMap tmg, tmx;
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
    uint32_t box[4] = {(uint32_t)p.co_slab, (uint32_t)kBw, (uint32_t)(p.ky_stack ? kBh + 2 : kBh), 1u};
    if (make_tmap_bf16(&tmg, g, 4, dims, str, box, (int)p.a_row_bytes) != 0) return 1;
  }
  {
    const int Hx = pool4 ? 2 * H : H, Wx = pool4 ? 2 * W : W;
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Wx, (uint64_t)Hx, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)Wx * Cin * 2, (uint64_t)Hx * Wx * Cin * 2};
    uint32_t box[4] = {(uint32_t)p.ci_sub, (uint32_t)kHaloW, (uint32_t)kBh, 1u};
    if (pool4) {
      // every other pixel / row: a (16+1) x 8 pixel tile per column parity
      uint32_t box2[4] = {(uint32_t)p.ci_sub, 2u * (kBw + 1), 2u * kBh, 1u};
      uint32_t est[4] = {1u, 2u, 2u, 1u};
      if (make_tmap_bf16_strided(&tmx, x, 4, dims, str, box2, est, (int)p.b_row_bytes) != 0) return 1;
    } else if (make_tmap_bf16(&tmx, x, 4, dims, str, box, (int)p.b_row_bytes) != 0) return 1;
  }

  if (!accumulate)
    if (launch_zero(dw, (size_t)(pool4 ? 16 : 9) * Cout * Cin * sizeof(float), stream) != 0) return 1;
  const size_t smem_bytes = (size_t)kStages * kStageBytes + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    BG_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  BG_CHECK_CUDA(launch_pdl(conv_wgrad_halo_kernel, units * splits, kThreads, smem_bytes, stream, tmg, tmx, p));
  return 0;
}

}  // namespace bg
