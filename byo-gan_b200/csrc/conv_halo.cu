// 3x3 pad-1 implicit-GEMM convolution with a HALO-RESIDENT input tile (resolutions >= 16x16).
//
// Same contract as conv_fprop.cu (EqualizedConv2d.forward, gan.py:29-38; its input-gradient on the flipped
// pack; the R1 tangent pass) but a different operand feed.  The tap-wise TMA kernel re-reads every input
// pixel 9 times from L2 and is bound by the per-SM L2->SMEM fill rate (~42 B/clk, see
// profiles/r1_ncu_full_conv_fprop_256x256.csv).  Here a CTA owns a 16x16 output tile (two M=128 MMA halves)
// and loads the 18x18 input halo of a channel chunk ONCE; the nine taps are nine shifted VIEWS of it:
//
//   smem A stage = [18*18 halo pixels][kc channels], one pixel = one 128/64/32-byte row, written by ONE TMA box load
//   (kc x 18 x 18 x 1 of the NHWC input, out-of-bounds zero fill = the conv padding) with the matching 128/64/32-byte
//   swizzle.  That is exactly UMMA's swizzled K-major layout with SBO = one halo row (18 pixels), and because the
//   swizzle is a function of the absolute shared-memory address, the view of tap (ky,kx) for MMA half h is just
//   start address += (ky*18 + kx + 8h) pixel rows.  Every operand row stays inside its own 128-byte line, so the
//   shifted views cost no extra shared-memory wavefronts (an earlier no-swizzle plane layout needed cp.async
//   producers and paid ~2x on the A reads: its 16-byte-shifted core matrices straddled two lines).
//
// Weights stream by TMA (128/64/32-byte swizzle) per (chunk, tap), or stay RESIDENT in shared memory for the whole
// kernel when the layer's pack fits (all the high-resolution layers).  Accumulators: TMEM, 2 stages x 2 halves x 128
// columns.  Epilogue: 8 warps; bias + noise + LeakyReLU (+ gate from a saved activation) in registers, then the bf16
// tile is transposed through a swizzled shared-memory stage so that every global store (and gate load) instruction
// moves whole 128-byte lines; an optional 2x2 average pool (warp shuffles) runs before the activation.
//
// Warp roles: 0 = weight TMA producer, 1 = MMA issuer (first M=128 half) + TMEM allocator, 19 = MMA issuer of the second
// half (one thread tops out at one tcgen05.mma per ~53 clk; two reach the pipe's floor), 2..9 = epilogue set 0 (one 32-row TMEM
// quarter of one MMA half each), 10..17 = epilogue set 1 in plain mode (the epilogue of a tile is a latency chain
// — TMEM load, math, transpose, stores — that is longer than the tile's MMAs on most layers, so two sets alternate
// tiles, each owning one accumulator stage) or, in upsample mode, the 256 threads that build the bilinearly
// upsampled halo from a low-resolution patch; 18 = halo TMA issuer (plain mode).
#include "common.cuh"

#include <stdlib.h>

namespace bg {

namespace {

constexpr int kEpiWarps = 8;                     // one epilogue SET: 4 TMEM lane quarters x 2 MMA halves
constexpr int kEpiSets = 2;                      // plain mode: set s drains accumulator stage s (every other tile)
constexpr int kProdWarps = 8;                    // upsample mode: these warps build the halo instead of being set 1
constexpr int kProdThreads = 32 * kProdWarps;
constexpr int kIssuer2 = 2 + kEpiWarps * kEpiSets + 1;         // warp 19: second MMA issuer (see HaloParams::issuers)
constexpr int kThreads = 32 * (kIssuer2 + 1);                  // warp 18: halo TMA issuer (plain mode)
constexpr int kTile = 16;                       // output tile edge
constexpr int kHalo = kTile + 2;                // 18
constexpr int kHaloPix = kHalo * kHalo;         // 324
constexpr int kMaxBStages = 8;
constexpr int kMaxAStages = 8;
constexpr int kMaxN = 128;
constexpr uint32_t kTmemCols = 512;

struct HaloParams {
  int N, H, W, Cin, Cout;
  int tw_shift, th_shift, nb_shift;     // log2 of tiles per row / per column / n-blocks (all powers of two)
  int block_n;
  int kc, k_chunks, cpp_shift;
  int a_stages, b_stages, b_resident;
  uint32_t a_stage_bytes, a_tx_bytes, b_tile_bytes, b_tx_bytes;
  uint32_t a_layout, a_sbo, b_layout, b_sbo;
  uint32_t epi_row_bytes;               // bytes of one pixel row in the epilogue transpose stage (32 / 64 / 128)
  int num_tiles;
  const __nv_bfloat16* x;
  const float* bias;
  const float* noise;
  const float* noise_w;
  const __nv_bfloat16* gate_src;
  __nv_bfloat16* out;
  int act;
  int debug;      // bisecting aid (BG_HALO_DEBUG): 1 no global stores, 2 no MMAs, 4 no halo copies, 8 no epilogue math
  int pool;       // 1: 2x2 average pool before the activation; out / gate_src are (N, H/2, W/2, Cout)
  float slope;
  // fused per-channel reductions of the (bf16-rounded) output, accumulated with fp32 atomics into a zeroed buffer:
  //   stats_mode 1: stats[n][c][0] += sum_hw out, stats[n][c][1] += sum_hw out^2   (instance-norm statistics, gan.py:59)
  //   stats_mode 2: stats[c] += sum_{n,hw} out                                      (bias gradient of the layer below)
  float* stats;
  int stats_mode;
  int n_blocks;
  int contig;     // tile order, see tile_range()
  int epi_sets;   // 1 or 2 epilogue warp sets (2: set s drains accumulator stage s)
  // MMA-issuing threads.  One thread cannot issue tcgen05.mma faster than one per ~53 clk, two threads reach the tensor
  // pipe's own floor (39 clk for N <= 32, 48 clk for N = 64: tools/mma2_probe.cu); with 2, warp 1 issues the MMAs of the
  // tile's first M=128 half and warp 19 those of the second, and every barrier the issuers commit to expects 2 arrivals.
  int issuers;
  // StyleGAN generator forward (gan.py:89-98,118-127) with the layer's AdaIN folded into the operands:
  //   per_sample_w: the weight pack is [N][9][Cout][Cin] (instance-norm scale * style gamma folded in per sample);
  //   bias_tab:     fp32 [N][9][Cout] replaces bias: bias + the conv of the per-channel AdaIN shift, one row per border
  //                 class (3 row classes x 3 column classes: which taps fall into the zero padding);
  //   upsample:     x is (N, H/2, W/2, Cin); the bilinear x2 upsample (align_corners=False) is applied while the
  //                 halo tile is built (patch -> shared memory -> 18x18 halo), so the 4x larger map never exists.
  int per_sample_w;
  const float* bias_tab;
  int upsample;
  uint32_t patch_bytes;   // one low-resolution 10x10 patch stage (upsample mode)
  // pool4 mode: conv3x3 followed by AvgPool2d(2) (CriticBlock.conv_2, gan.py:258-260) computed as ONE 4x4 stride-2
  // convolution (16 taps, weights = quarter sums of the shifted 3x3 kernel): N, H, W are the POOLED output's, the
  // input is (N, 2H, 2W, Cin).  The 34x34 input region of a tile is loaded as four 17x17 parity-phase tiles (TMA boxes
  // with element stride 2 along W and H, one pipeline stage each), so that tap (a, b) is phase (a&1, b&1) shifted by
  // (a>>1, b>>1) pixel rows — 2.25x fewer MMAs than conv-then-pool.
  // tconv4 mode: the input-gradient of that 4x4 stride-2 conv (a transposed conv): x is the POOLED gradient map
  // (N, H, W, Cin), out the full-resolution map (N, 2H, 2W, Cout).  Output parity phase (py, px) is a 2x2-tap conv over
  // the pooled map (4 of the 16 taps), so a work item is (16x16 pooled tile, phase): views of the ordinary 18x18 halo
  // shifted by (py + 1 - ta, px + 1 - tb), results stored to pixels (2h + py, 2w + px).
  int tconv4;
  int ph_shift;   // 2 in tconv4 mode (phase = low bits of the tile index above the n-block), else 0
  int pool4;
  int pool4_tma;  // == pool4: the four phase tiles are loaded by strided TMA boxes, one pipeline stage each
  int taps;       // 9, or 16 in pool4 mode
  int Hin, Win;
  uint32_t phase_bytes;
  // CTA pair mode (wide layers, block_n = 128): the two CTAs of a 2-CTA cluster work on two horizontally adjacent
  // tiles with the same weights; every tcgen05.mma is ONE M=256 cta_group::2 instruction issued by the leader CTA
  // (rows 0..127 = this half of the leader's tile, rows 128..255 = the same half of the peer's), and each CTA keeps only
  // HALF of the weight tile (block_n / 2 rows) in shared memory: per instruction an SM reads A 4 KB + B 2 KB instead of
  // 4 + 4 KB, which takes the operand fetch off the 128 B/clk shared-memory limit that N = 128 sits on.
  // b_tx_bytes / b_tile_bytes describe the per-CTA half tile.  Tile indices handed to decode_tile are PAIR indices.
  int cta2;
  int up_packed;  // upsample producers blend on packed bf16x2 pairs (BG_UP_PACKED=0: fp32 blends, one rounding)
  int epi_templated;   // 1: epilogue specialised for the stage row width (default); BG_EPI_TEMPLATED=0: run-time widths
};

struct TileCoord {
  int w0, h0, n, co0, ph;
};

__device__ __forceinline__ TileCoord decode_tile(const HaloParams& p, int tile, uint32_t crank = 0) {
  TileCoord t;
  const int nb = tile & ((1 << p.nb_shift) - 1);
  int pt = tile >> p.nb_shift;
  t.ph = pt & ((1 << p.ph_shift) - 1);
  pt >>= p.ph_shift;
  if (p.cta2) {
    // pair index: the pair covers tile columns 2j (leader) and 2j + 1 (peer) of the same tile row, n-block and sample
    const int tws = p.tw_shift - 1;
    t.w0 = (((pt & ((1 << tws) - 1)) << 1) + (int)crank) * kTile;
    pt >>= tws;
  } else {
    t.w0 = (pt & ((1 << p.tw_shift) - 1)) * kTile;
    pt >>= p.tw_shift;
  }
  t.h0 = (pt & ((1 << p.th_shift) - 1)) * kTile;
  t.n = pt >> p.th_shift;
  t.co0 = nb * p.block_n;
  return t;
}

// Tile order.  contig = 1: a CTA owns a CONTIGUOUS range of tiles, so consecutive tiles share the sample index and
// the epilogue can keep per-sample reductions in registers across tiles.  contig = 0: tiles are dealt round-robin,
// so that at any moment the 148 CTAs work on 148 neighbouring tiles (best L2 / DRAM-page locality).
__device__ __forceinline__ void tile_range(const HaloParams& p, int& t0, int& t1, int& step) {
  // pair mode: the unit of work is a tile PAIR and the unit of the grid a CTA pair (both CTAs walk the same sequence)
  const int bid = p.cta2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nblk = p.cta2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int ntiles = p.cta2 ? p.num_tiles >> 1 : p.num_tiles;
  if (p.contig) {
    t0 = (int)(((long long)ntiles * (long long)bid) / (long long)nblk);
    t1 = (int)(((long long)ntiles * (long long)(bid + 1)) / (long long)nblk);
    step = 1;
  } else {
    t0 = bid;
    t1 = ntiles;
    step = nblk;
  }
}

__device__ __forceinline__ float4 lds128f(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async_16_full(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// debug aid (BG_HALO_DEBUG & 4, "no halo copies"): complete the expected transaction bytes without loading anything
__device__ __forceinline__ void mbar_arrive_tx_debug(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------
// MMA issue (warp 1).  For the narrow layers an M=128 x N<=64 MMA retires in 16-48 clk, so the issue loop itself
// is the critical path: the whole warp runs it converged (all values warp-uniform, so ptxas keeps descriptors in
// uniform registers), one elected lane executes the tcgen05 instructions, descriptors are 64-bit values advanced
// by compile-time constants, and the k-steps / taps are unrolled.
// ---------------------------------------------------------------------------------------------------------
// Descriptor offset (16-byte units) of the A view of one tap.  9 taps: the 18x18 halo shifted by (ky, kx) pixel rows.
// 16 taps (pool4): parity-phase tile (a&1, b&1) of the 34x34 region, shifted by (a>>1, b>>1) inside its 17x17 pixels.
// TAPS == 4 (tconv4): tap (ta, tb) of output phase ph = 2*py + px reads the halo shifted by (py + 1 - ta, px + 1 - tb);
// `phase_units` carries ph in that mode.
template <int TAPS, uint32_t RBU>
__device__ __forceinline__ uint64_t a_tap_offset(int tap, uint32_t phase_units) {
  if (TAPS == 9) return (uint64_t)(((tap / 3) * kHalo + (tap % 3)) * RBU);
  const int py = (int)(phase_units >> 1), px = (int)(phase_units & 1u);
  return (uint64_t)(((py + 1 - (tap >> 1)) * kHalo + (px + 1 - (tap & 1))) * RBU);
}

// pool4 with TMA feed: the A pipeline unit is ONE parity-phase tile (17x17 pixels x kc channels); phase ph = 2*py + px
// serves the four taps (a, b) = (py + 2i, px + 2j), i, j in {0, 1}, whose view is the tile shifted by (i, j) pixels.
__device__ __forceinline__ int pool4_tap(int ph, int t4) { return ((ph >> 1) + 2 * (t4 >> 1)) * 4 + (ph & 1) + 2 * (t4 & 1); }

// HSEL: -1 = this thread issues both M=128 halves of the tile, 0 / 1 = only that half (two issuing warps)
template <int KSTEPS, int HSEL, bool CTA2>
__device__ __forceinline__ void mma_issue_loop_pool4(const HaloParams& p, uint8_t* a_base, uint8_t* b_base,
                                                     uint64_t* b_full, uint64_t* b_empty, uint64_t* a_full,
                                                     uint64_t* a_empty, uint64_t* tmem_full, uint64_t* tmem_empty,
                                                     uint32_t tmem_base) {
  constexpr uint32_t kRBU = 2u * KSTEPS;
  const uint32_t idesc = umma_idesc_bf16(CTA2 ? 256 : 128, p.block_n, 0, 0);
  const uint64_t a_desc0 = umma_desc(smem_u32(a_base), 16u, p.a_sbo, p.a_layout);
  const uint64_t b_desc0 = umma_desc(smem_u32(b_base), 16u, p.b_sbo, p.b_layout);
  const uint32_t a_stage_step = p.a_stage_bytes >> 4;
  const uint32_t b_tile_step = p.b_tile_bytes >> 4;
  const bool leader = elect_one();
  int tile_lo, tile_hi, tile_step;
  tile_range(p, tile_lo, tile_hi, tile_step);
  int bstage = 0, astage = 0, acc = 0;
  uint32_t bphase = 0, aphase = 0, acc_phase = 0;
  for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
    mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
    uint32_t accum = 0u;
    for (int kcx = 0; kcx < p.k_chunks; ++kcx) {
      for (int ph = 0; ph < 4; ++ph) {
        mbar_wait(&a_full[astage], aphase);
        tc_fence_after();
        const uint64_t a_stage = a_desc0 + (uint64_t)((uint32_t)astage * a_stage_step);
#pragma unroll
        for (int t4 = 0; t4 < 4; ++t4) {
          mbar_wait(&b_full[bstage], bphase);
          tc_fence_after();
          if (leader) {
            const uint64_t a_tap = a_stage + (uint64_t)(((t4 >> 1) * (kTile + 1) + (t4 & 1)) * kRBU);
            const uint64_t b_tap = b_desc0 + (uint64_t)((uint32_t)bstage * b_tile_step);
            if (!(p.debug & 2)) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
#pragma unroll
                for (int half = (HSEL < 0 ? 0 : HSEL); half < (HSEL < 0 ? 2 : HSEL + 1); ++half)
                  tc_mma_g<CTA2>(d_tmem + (uint32_t)half * 128u, a_tap + (uint64_t)(half * 8 * kRBU + k * 2),
                              b_tap + (uint64_t)(k * 2), idesc, k == 0 ? accum : 1u);
              }
            }
            tc_commit_g<CTA2>(&b_empty[bstage]);
          }
          accum = 1u;
          if (++bstage == p.b_stages) {
            bstage = 0;
            bphase ^= 1u;
          }
        }
        if (leader) tc_commit_g<CTA2>(&a_empty[astage]);
        __syncwarp();
        if (++astage == p.a_stages) {
          astage = 0;
          aphase ^= 1u;
        }
      }
    }
    if (leader) tc_commit_g<CTA2>(&tmem_full[acc]);
    __syncwarp();
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1u;
  }
}

template <int KSTEPS, int TAPS, int HSEL, bool CTA2>
__device__ __forceinline__ void mma_issue_loop(const HaloParams& p, uint8_t* a_base, uint8_t* b_base, uint64_t* b_full,
                                               uint64_t* b_empty, uint64_t* a_full, uint64_t* a_empty,
                                               uint64_t* tmem_full, uint64_t* tmem_empty, uint32_t tmem_base) {
  const uint32_t idesc = umma_idesc_bf16(CTA2 ? 256 : 128, p.block_n, 0, 0);
  constexpr uint32_t kRBU = 2u * KSTEPS;                       // one pixel row (kc bf16) in 16-byte units
  const uint64_t a_desc0 = umma_desc(smem_u32(a_base), 16u, p.a_sbo, p.a_layout);
  const uint64_t b_desc0 = umma_desc(smem_u32(b_base), 16u, p.b_sbo, p.b_layout);
  const uint32_t a_stage_step = p.a_stage_bytes >> 4;
  const uint32_t phase_units = p.phase_bytes >> 4;
  const uint32_t b_tile_step = p.b_tile_bytes >> 4;
  const bool leader = elect_one();
  int tile_lo, tile_hi, tile_step;
  tile_range(p, tile_lo, tile_hi, tile_step);
  int bstage = 0;
  uint32_t bphase = 0;
  int astage = 0;
  uint32_t aphase = 0;
  int acc = 0;
  uint32_t acc_phase = 0;
  int w_cur = -1;
  uint32_t w_loads = 0;
  for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
    if (p.b_resident) {
      const int wn = p.per_sample_w ? decode_tile(p, tile).n : 0;
      if (wn != w_cur) {                      // first tile, or the tile range moved on to another sample's pack
        if (w_loads > 0 && leader) tc_commit_g<CTA2>(&b_empty[0]);
        __syncwarp();
        mbar_wait(&b_full[0], w_loads & 1u);
        tc_fence_after();
        w_cur = wn;
        ++w_loads;
      }
    }
    mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
    uint32_t accum = 0u;
    // tconv4 (TAPS == 4): this work item's output phase selects the 4 taps (of the 16 resident ones) and their views
    const uint32_t tap_sel = (TAPS == 4) ? (uint32_t)decode_tile(p, tile).ph : phase_units;
    for (int kcx = 0; kcx < p.k_chunks; ++kcx) {
      mbar_wait(&a_full[astage], aphase);
      tc_fence_after();
      const uint64_t a_stage = a_desc0 + (uint64_t)((uint32_t)astage * a_stage_step);
      if (p.b_resident) {
        const uint64_t b_chunk =
            b_desc0 + (uint64_t)((uint32_t)(TAPS == 4 ? kcx * 16 + (int)tap_sel * 4 : kcx * TAPS) * b_tile_step);
        if (leader) {
          if (!(p.debug & 2)) {
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
              const uint64_t a_tap = a_stage + a_tap_offset<TAPS, kRBU>(tap, tap_sel);
              const uint64_t b_tap = b_chunk + (uint64_t)((uint32_t)tap * b_tile_step);
              // alternate the two accumulators (MMA halves): back-to-back MMAs into the SAME accumulator serialise
              // on the tensor pipe's latency (~60 clk), which is longer than a small-N MMA itself
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
#pragma unroll
                for (int half = (HSEL < 0 ? 0 : HSEL); half < (HSEL < 0 ? 2 : HSEL + 1); ++half) {
                  tc_mma_g<CTA2>(d_tmem + (uint32_t)half * 128u,
                              a_tap + (uint64_t)(half * 8 * kRBU + k * 2), b_tap + (uint64_t)(k * 2),
                              idesc, (k == 0 && tap == 0) ? accum : 1u);
                }
              }
            }
          }
          tc_commit_g<CTA2>(&a_empty[astage]);
        }
        accum = 1u;
      } else {
        for (int tap = 0; tap < TAPS; ++tap) {
          mbar_wait(&b_full[bstage], bphase);
          tc_fence_after();
          if (leader) {
            const uint64_t a_tap = a_stage + a_tap_offset<TAPS, kRBU>(tap, tap_sel);
            const uint64_t b_tap = b_desc0 + (uint64_t)((uint32_t)bstage * b_tile_step);
            if (!(p.debug & 2)) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
#pragma unroll
                for (int half = (HSEL < 0 ? 0 : HSEL); half < (HSEL < 0 ? 2 : HSEL + 1); ++half) {
                  tc_mma_g<CTA2>(d_tmem + (uint32_t)half * 128u,
                              a_tap + (uint64_t)(half * 8 * kRBU + k * 2), b_tap + (uint64_t)(k * 2),
                              idesc, k == 0 ? accum : 1u);
                }
              }
            }
            tc_commit_g<CTA2>(&b_empty[bstage]);
          }
          accum = 1u;
          if (++bstage == p.b_stages) {
            bstage = 0;
            bphase ^= 1u;
          }
        }
        if (leader) tc_commit_g<CTA2>(&a_empty[astage]);
      }
      __syncwarp();
      if (++astage == p.a_stages) {
        astage = 0;
        aphase ^= 1u;
      }
    }
    if (leader) tc_commit_g<CTA2>(&tmem_full[acc]);
    __syncwarp();
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1u;
  }
}

// The issue loop specialised for the launch's chunk width / tap count, for the half (or both halves) this warp owns.
template <bool kStats, int kFeed, int HSEL, bool CTA2>
__device__ __forceinline__ void issue_dispatch(const HaloParams& p, uint8_t* a_base, uint8_t* b_base, uint64_t* b_full,
                                               uint64_t* b_empty, uint64_t* a_full, uint64_t* a_empty,
                                               uint64_t* tmem_full, uint64_t* tmem_empty, uint32_t tmem_base) {
  if (kFeed == 0 && !kStats && p.pool4) {
    if (p.kc == 64) mma_issue_loop_pool4<4, HSEL, CTA2>(p, a_base, b_base, b_full, b_empty, a_full, a_empty, tmem_full, tmem_empty, tmem_base);
    else mma_issue_loop_pool4<2, HSEL, CTA2>(p, a_base, b_base, b_full, b_empty, a_full, a_empty, tmem_full, tmem_empty, tmem_base);
  } else if (p.tconv4) {
    if (p.kc == 64) mma_issue_loop<4, 4, HSEL, CTA2>(p, a_base, b_base, b_full, b_empty, a_full, a_empty, tmem_full, tmem_empty, tmem_base);
    else if (p.kc == 32) mma_issue_loop<2, 4, HSEL, CTA2>(p, a_base, b_base, b_full, b_empty, a_full, a_empty, tmem_full, tmem_empty, tmem_base);
    else mma_issue_loop<1, 4, HSEL, CTA2>(p, a_base, b_base, b_full, b_empty, a_full, a_empty, tmem_full, tmem_empty, tmem_base);
  } else if (p.kc == 64) mma_issue_loop<4, 9, HSEL, CTA2>(p, a_base, b_base, b_full, b_empty, a_full, a_empty, tmem_full, tmem_empty, tmem_base);
  else if (p.kc == 32) mma_issue_loop<2, 9, HSEL, CTA2>(p, a_base, b_base, b_full, b_empty, a_full, a_empty, tmem_full, tmem_empty, tmem_base);
  else mma_issue_loop<1, 9, HSEL, CTA2>(p, a_base, b_base, b_full, b_empty, a_full, a_empty, tmem_full, tmem_empty, tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// Halo feed, plain mode: one thread issues one TMA box load (kc channels x 18 x 18 pixels x 1 image) per (tile,
// channel chunk); coordinates start at (w0-1, h0-1), the out-of-bounds part of the box is zero-filled by the TMA
// unit — that IS the convolution's padding.
// ---------------------------------------------------------------------------------------------------------
// Pair mode: both CTAs load their own tile's halo into their own shared memory; the bytes of BOTH loads are counted on
// the LEADER's a_full barrier (the leader's thread posts the expected total), because only the leader's MMA warps wait.
// (compile-time switch: a kernel that merely CONTAINS cta_group::2 instructions cannot be launched without a cluster)
template <bool CTA2>
__device__ __forceinline__ void halo_load(const HaloParams& p, const CUtensorMap* tmap_x, uint64_t* full, void* dst, int c0,
                                          int c1, int c2, int c3, uint32_t crank) {
  if (CTA2) {
    if (crank == 0) mbar_expect_tx(full, 2u * p.a_tx_bytes);
    tma_load_4d_2sm(tmap_x, mapa_cta(smem_u32(full), 0u), dst, c0, c1, c2, c3);
  } else {
    mbar_expect_tx(full, p.a_tx_bytes);
    if (p.debug & 4) mbar_arrive_tx_debug(full, p.a_tx_bytes);
    else tma_load_4d(tmap_x, full, dst, c0, c1, c2, c3);
  }
}

template <bool CTA2>
__device__ __forceinline__ void halo_tma_loop(const HaloParams& p, const CUtensorMap* tmap_x, uint8_t* a_base,
                                              uint64_t* a_full, uint64_t* a_empty, uint32_t crank) {
  int tile_lo, tile_hi, tile_step;
  tile_range(p, tile_lo, tile_hi, tile_step);
  int stage = 0;
  uint32_t phase = 0;
  for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
    const TileCoord t = decode_tile(p, tile, crank);
    for (int kcx = 0; kcx < p.k_chunks; ++kcx) {
      if (p.pool4) {
        // four pipeline units per chunk, one per parity-phase tile of the 34x34 input region: the tensor map walks W
        // and H with element stride 2, so each unit is a dense 17x17-pixel tile
        for (int ph = 0; ph < 4; ++ph) {
          mbar_wait(&a_empty[stage], phase ^ 1u);
          halo_load<CTA2>(p, tmap_x, &a_full[stage], a_base + (size_t)stage * p.a_stage_bytes, kcx * p.kc,
                    2 * t.w0 - 1 + (ph & 1), 2 * t.h0 - 1 + (ph >> 1), t.n, crank);
          if (++stage == p.a_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        continue;
      }
      mbar_wait(&a_empty[stage], phase ^ 1u);
      halo_load<CTA2>(p, tmap_x, &a_full[stage], a_base + (size_t)stage * p.a_stage_bytes, kcx * p.kc, t.w0 - 1, t.h0 - 1, t.n,
                crank);
      if (++stage == p.a_stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Halo producers, upsample mode: the conv input is nn.Upsample(scale_factor=2, mode='bilinear') of x (gan.py:112,123;
// align_corners=False: taps .75/.25, indices clamped at the edges).  Two phases per (tile, channel chunk):
//   1. cp.async the 10x10 low-resolution patch that the 18x18 halo depends on into a 4-deep patch ring (issued two
//      chunks ahead; indices are clamped at load time, so every copy is in bounds);
//   2. after a producer-wide named barrier, each thread builds 2x2 output quads from 2x2 patch pixels (separable
//      .75/.25 blends in fp32), writes them into the UMMA plane layout of the A stage (zeros outside the image: the
//      conv's padding applies to the UPSAMPLED map) and arrives on a_full.
// Thread -> (quad, 16-byte channel group): u = pt + 256 i, c8 = u % cpp, quad = u / cpp, quad = qk * 9 + ql covers
// halo rows 2qk, 2qk+1 and columns 2ql, 2ql+1 from patch rows qk, qk+1 and columns ql, ql+1.
// ---------------------------------------------------------------------------------------------------------
constexpr int kPatch = 10;
constexpr int kPatchStages = 4;

template <int CPP_SHIFT>
__device__ __forceinline__ void halo_producer_upsample(const HaloParams& p, uint8_t* a_base, uint8_t* patch_base,
                                                       uint64_t* a_full, uint64_t* a_empty, int pt) {
  constexpr int kCpp = 1 << CPP_SHIFT;
  constexpr int kPatchUnits = kPatch * kPatch * kCpp;
  constexpr int kLoadIters = (kPatchUnits + kProdThreads - 1) / kProdThreads;     // 4 / 2 / 1
  constexpr int kQuadUnits = 81 * kCpp;
  constexpr int kQuadIters = (kQuadUnits + kProdThreads - 1) / kProdThreads;      // 3 / 2 / 1
  constexpr uint32_t kPixBytes = 16u * kCpp;                                        // one pixel row (kc channels)
  constexpr uint32_t kSwzMask = kCpp - 1u;                                          // 7 / 3 / 1: SWIZZLE_128B / 64B / 32B
  const int HL = p.H >> 1, WL = p.W >> 1;
  int tile_lo, tile_hi, tile_step;
  tile_range(p, tile_lo, tile_hi, tile_step);
  const int total = ((tile_hi - tile_lo + tile_step - 1) / tile_step) * p.k_chunks;   // (tile, chunk) items of this CTA

  // ---- patch loader state (runs two items ahead of the builder)
  int l_item = 0, l_tile = tile_lo, l_kcx = 0;
  auto issue_patch = [&]() {
    if (l_item < total) {
      const TileCoord t = decode_tile(p, l_tile);
      const int k0 = (t.h0 >> 1) - 1, l0 = (t.w0 >> 1) - 1;
      const uint32_t sdst = smem_u32(patch_base + (size_t)(l_item & (kPatchStages - 1)) * p.patch_bytes);
      const __nv_bfloat16* src = p.x + (size_t)t.n * HL * WL * p.Cin + l_kcx * p.kc;
#pragma unroll
      for (int i = 0; i < kLoadIters; ++i) {
        const int u = pt + i * kProdThreads;
        if (i < kLoadIters - 1 || u < kPatchUnits) {
          const int px = u >> CPP_SHIFT, c8 = u & (kCpp - 1);
          const int pr = px / kPatch, pc = px - pr * kPatch;
          const int r = min(max(k0 + pr, 0), HL - 1), c = min(max(l0 + pc, 0), WL - 1);
          cp_async_16_full(sdst + (uint32_t)px * kPixBytes + (uint32_t)c8 * 16u, src + ((size_t)r * WL + c) * p.Cin + c8 * 8);
        }
      }
      if (++l_kcx == p.k_chunks) {
        l_kcx = 0;
        l_tile += tile_step;
      }
    }
    ++l_item;
    cp_async_commit();          // always commit (possibly empty) so that wait_group<2> counts items uniformly
  };
  issue_patch();
  issue_patch();

  // Everything about a thread's quads that does not depend on the tile — patch offset of the 2x2 sources, halo
  // coordinates, the four swizzled destination offsets — is worked out once (the per-quad index arithmetic was 140 of
  // the ~330 instructions a producer warp spent per quad, profiles/r2_sass_mix_style512.txt)
  uint32_t q_pa[kQuadIters], q_yx[kQuadIters], q_dst[kQuadIters][4];
#pragma unroll
  for (int i = 0; i < kQuadIters; ++i) {
    const int u = min(pt + i * kProdThreads, kQuadUnits - 1);
    const int q = u >> CPP_SHIFT, c8 = u & (kCpp - 1);
    const int qk = q / 9, ql = q - qk * 9;
    const int hy = 2 * qk, hx = 2 * ql;
    q_pa[i] = (uint32_t)(qk * kPatch + ql) * kPixBytes + (uint32_t)c8 * 16u;
    q_yx[i] = (uint32_t)hy | ((uint32_t)hx << 8);
    // A stage = [halo pixel][kc channels] with the 128/64/32-byte swizzle UMMA expects: the 16-byte chunk index
    // is XORed with address bits 7.. (stage bases are 1024-byte aligned, so offsets can stand in for addresses)
    const uint32_t l0 = (uint32_t)(hy * kHalo + hx) * kPixBytes + (uint32_t)c8 * 16u;
    const uint32_t l1 = l0 + kPixBytes, l2 = l0 + kHalo * kPixBytes, l3 = l2 + kPixBytes;
    q_dst[i][0] = l0 ^ (((l0 >> 7) & kSwzMask) << 4);
    q_dst[i][1] = l1 ^ (((l1 >> 7) & kSwzMask) << 4);
    q_dst[i][2] = l2 ^ (((l2 >> 7) & kSwzMask) << 4);
    q_dst[i][3] = l3 ^ (((l3 >> 7) & kSwzMask) << 4);
  }

  int stage = 0;
  uint32_t phase = 0;
  int item = 0;
  for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
    const TileCoord t = decode_tile(p, tile);
    const int hb = t.h0 - 1, wb = t.w0 - 1;
    for (int kcx = 0; kcx < p.k_chunks; ++kcx, ++item) {
      issue_patch();                                        // item + 2
      cp_async_wait<2>();                                   // this thread's copies of `item` have landed
      if (pt < 32) mbar_wait(&a_empty[stage], phase ^ 1u);  // ONE warp polls for the free stage (8 polling warps only burn issue slots)
      asm volatile("bar.sync 2, %0;" ::"n"(kProdThreads) : "memory");   // ... everybody's copies have landed, the stage is free
      const uint32_t patch = smem_u32(patch_base + (size_t)(item & (kPatchStages - 1)) * p.patch_bytes);
      const uint32_t sdst = smem_u32(a_base + (size_t)stage * p.a_stage_bytes);
#pragma unroll
      for (int i = 0; i < kQuadIters; ++i) {
        const int u = pt + i * kProdThreads;
        if (i < kQuadIters - 1 || u < kQuadUnits) {
          const uint32_t pa = patch + q_pa[i];
          const uint4 r00 = lds128(pa), r01 = lds128(pa + kPixBytes), r10 = lds128(pa + kPatch * kPixBytes),
                      r11 = lds128(pa + (kPatch + 1) * kPixBytes);
          const uint32_t w00[4] = {r00.x, r00.y, r00.z, r00.w}, w01[4] = {r01.x, r01.y, r01.z, r01.w};
          const uint32_t w10[4] = {r10.x, r10.y, r10.z, r10.w}, w11[4] = {r11.x, r11.y, r11.z, r11.w};
          uint32_t oAP[4], oAQ[4], oBP[4], oBQ[4];
          if (p.up_packed) {
            // the four outputs of a quad straight from its four sources on PACKED bf16x2 pairs:
            //   out = 9/16 near + 3/16 side + 3/16 other side + 1/16 far   (the product of the two .75/.25 blends),
            // one HMUL2 + three HFMA2 per pair and output, smallest term first (each fma rounds a partial sum that is
            // smaller than the result: ~1.1x the error of one final rounding) — 64 instead of ~176 instructions per
            // quad and channel group, no bf16 <-> fp32 conversions; the producer warps were the kernel's issue limit
            const __nv_bfloat162 k9 = __floats2bfloat162_rn(0.5625f, 0.5625f), k3 = __floats2bfloat162_rn(0.1875f, 0.1875f),
                                 k1 = __floats2bfloat162_rn(0.0625f, 0.0625f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&w00[j]);
              const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w01[j]);
              const __nv_bfloat162 c = *reinterpret_cast<const __nv_bfloat162*>(&w10[j]);
              const __nv_bfloat162 d = *reinterpret_cast<const __nv_bfloat162*>(&w11[j]);
              const __nv_bfloat162 ap = __hfma2(a, k9, __hfma2(b, k3, __hfma2(c, k3, __hmul2(d, k1))));
              const __nv_bfloat162 aq = __hfma2(b, k9, __hfma2(a, k3, __hfma2(d, k3, __hmul2(c, k1))));
              const __nv_bfloat162 bp = __hfma2(c, k9, __hfma2(a, k3, __hfma2(d, k3, __hmul2(b, k1))));
              const __nv_bfloat162 bq = __hfma2(d, k9, __hfma2(b, k3, __hfma2(c, k3, __hmul2(a, k1))));
              oAP[j] = *reinterpret_cast<const uint32_t*>(&ap);
              oAQ[j] = *reinterpret_cast<const uint32_t*>(&aq);
              oBP[j] = *reinterpret_cast<const uint32_t*>(&bp);
              oBQ[j] = *reinterpret_cast<const uint32_t*>(&bq);
            }
          } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 a = unpack_bf16x2(w00[j]), b = unpack_bf16x2(w01[j]);
            const float2 c = unpack_bf16x2(w10[j]), d = unpack_bf16x2(w11[j]);
            // columns: P = left output column (.75 left pixel), Q = right output column (.75 right pixel)
            const float p0x = 0.75f * a.x + 0.25f * b.x, p0y = 0.75f * a.y + 0.25f * b.y;
            const float q0x = 0.25f * a.x + 0.75f * b.x, q0y = 0.25f * a.y + 0.75f * b.y;
            const float p1x = 0.75f * c.x + 0.25f * d.x, p1y = 0.75f * c.y + 0.25f * d.y;
            const float q1x = 0.25f * c.x + 0.75f * d.x, q1y = 0.25f * c.y + 0.75f * d.y;
            // rows: A = upper output row (.75 upper pixel), B = lower output row (.75 lower pixel)
            oAP[j] = pack_bf16x2(0.75f * p0x + 0.25f * p1x, 0.75f * p0y + 0.25f * p1y);
            oAQ[j] = pack_bf16x2(0.75f * q0x + 0.25f * q1x, 0.75f * q0y + 0.25f * q1y);
            oBP[j] = pack_bf16x2(0.25f * p0x + 0.75f * p1x, 0.25f * p0y + 0.75f * p1y);
            oBQ[j] = pack_bf16x2(0.25f * q0x + 0.75f * q1x, 0.25f * q0y + 0.75f * q1y);
          }
          }
          const int hy = (int)(q_yx[i] & 0xffu), hx = (int)(q_yx[i] >> 8);
          const bool rA = (unsigned)(hb + hy) < (unsigned)p.H, rB = (unsigned)(hb + hy + 1) < (unsigned)p.H;
          const bool cP = (unsigned)(wb + hx) < (unsigned)p.W, cQ = (unsigned)(wb + hx + 1) < (unsigned)p.W;
          const uint4 z = make_uint4(0, 0, 0, 0);
          sts128(sdst + q_dst[i][0], (rA && cP) ? make_uint4(oAP[0], oAP[1], oAP[2], oAP[3]) : z);
          sts128(sdst + q_dst[i][1], (rA && cQ) ? make_uint4(oAQ[0], oAQ[1], oAQ[2], oAQ[3]) : z);
          sts128(sdst + q_dst[i][2], (rB && cP) ? make_uint4(oBP[0], oBP[1], oBP[2], oBP[3]) : z);
          sts128(sdst + q_dst[i][3], (rB && cQ) ? make_uint4(oBQ[0], oBQ[1], oBQ[2], oBQ[3]) : z);
        }
      }
      fence_proxy_async();
      mbar_arrive(&a_full[stage]);
      if (++stage == p.a_stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------------------
// Fused reductions of the output tile (epilogue warps).  During the line-wise store phase lane l owns the 16-byte
// channel chunk ch = l % cpr of rows l / cpr, l / cpr + 32 / cpr, ... and sums its 8 channels in registers; after the
// round a butterfly over the lanes that share a chunk leaves the warp totals in lanes 0..cpr-1, which add them to
// the warp's PRIVATE shared-memory slice [Cout][2] (plain load/add/store, no atomics).  The slices are combined and
// sent to global memory (red.global.add.f32) only when the CTA's tile sequence moves to another sample, or ends:
// a per-tile flush costs more than the convolution itself on the high-resolution layers.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stats_round(float* slice, float (&sa)[8], float (&sq)[8], int cpr_shift, int lane,
                                            int cbase) {
  const int cpr = 1 << cpr_shift;
  for (int o = 16; o >= cpr; o >>= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sa[j] += __shfl_xor_sync(0xffffffffu, sa[j], o);
      sq[j] += __shfl_xor_sync(0xffffffffu, sq[j], o);
    }
  }
  if (lane < cpr) {
    float4* d = reinterpret_cast<float4*>(slice + (size_t)(cbase + lane * 8) * 2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 v = d[j];
      v.x += sa[2 * j];
      v.y += sq[2 * j];
      v.z += sa[2 * j + 1];
      v.w += sq[2 * j + 1];
      d[j] = v;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sa[j] = sq[j] = 0.f;
}

// All 8 warps of an epilogue SET call this at the same point of their (identical) tile sequence; `slices` are the
// set's own 8 slices, `bar_id` its named barrier.
__device__ __forceinline__ void stats_flush(const HaloParams& p, float* slices, int et, int n, int bar_id) {
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * kEpiWarps) : "memory");
  const int len = p.Cout * 2;
  for (int i = et; i < len; i += 32 * kEpiWarps) {
    float v = 0.f;
#pragma unroll
    for (int e = 0; e < kEpiWarps; ++e) {
      v += slices[e * len + i];
      slices[e * len + i] = 0.f;
    }
    if (p.stats_mode == 1) atomicAdd(p.stats + (size_t)n * len + i, v);
    else if ((i & 1) == 0) atomicAdd(p.stats + (i >> 1), v);
  }
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * kEpiWarps) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Epilogue (one warp): TMEM -> registers -> bias / noise / activation / gate -> bf16 -> swizzled transpose stage ->
// 128-byte-line global stores (+ fused reductions).  Templated on the stage row width: an ncu capture of the narrow
// layers (profiles/r2_ncu_pool4_dgrad_stalls.txt) shows the SM issuing instructions 55 % of the time, ~30 thread
// instructions per output element, most of them the run-time swizzle / address arithmetic of these loops.
// ---------------------------------------------------------------------------------------------------------
template <bool kStats, bool kUp, bool kCta2, int CPRS>
__device__ __forceinline__ void epilogue_warp(const HaloParams& p, const int warp, const int lane, const uint32_t crank,
                                              const uint32_t tmem_base, uint64_t* tmem_full, uint64_t* tmem_empty,
                                              uint8_t* epi_base, float* bias_s, float* nw_s, float* stat_s, const int tile_lo,
                                              const int tile_hi, const int tile_step) {
    // ------------------------------ epilogue ------------------------------
    // Warp e = warp - 2: TMEM lane quarter q = warp & 3 (hardware rule), MMA half = e / 4.  Lane i holds MMA row
    // m = 32q + i = pixel (image row g = m / 8, column r = m % 8 of the half's 8-wide segment).
    const int q = warp & 3;
    const int eset = (warp - 2) >> 3;                  // 0, or 1 for the second set (plain mode)
    const int half = ((warp - 2) >> 2) & 1;
    const int g = (q * 32 + lane) >> 3, r = lane & 7;
    const bool pool_writer = ((lane & 1) == 0) && ((lane & 8) == 0);
    // transpose stage of this warp: 32 pixel rows x epi_row_bytes, 16-byte chunks XOR-swizzled by row so that both
    // the row-wise (lane = pixel) and the line-wise (8 lanes = 128 contiguous bytes) accesses are conflict-free
    // CPRS >= 0: the row width is a compile-time constant (16 << CPRS bytes), so the swizzle / address arithmetic of the
    // transpose loops folds and the loops unroll; CPRS < 0: taken from the launch parameters (BG_EPI_TEMPLATED=0)
    const uint32_t rb = CPRS >= 0 ? (16u << (CPRS >= 0 ? CPRS : 0)) : p.epi_row_bytes;
    const int cpr_shift = CPRS >= 0 ? CPRS : (rb == 128 ? 3 : (rb == 64 ? 2 : 1));       // log2(chunks per row)
    const int round_cols = (int)(rb >> 1);                          // output channels per transpose round
    const uint32_t stg = smem_u32(epi_base) + (uint32_t)(warp - 2) * 32u * rb;
    const uint32_t my_row = stg + (uint32_t)lane * rb;
    const uint32_t my_swz = (uint32_t)(lane >> (3 - cpr_shift)) & ((1u << cpr_shift) - 1u);
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr bool st_on = kStats;
    // one transpose round per tile and a fixed channel range: the lane partials stay in registers across tiles
    const bool st_regs = st_on && p.n_blocks == 1 && p.block_n <= round_cols;
    float* st_slice = stat_s + (size_t)(warp - 2) * p.Cout * 2;
    float* st_set = stat_s + (size_t)eset * kEpiWarps * p.Cout * 2;
    const int st_et = (int)threadIdx.x - 64 - eset * 32 * kEpiWarps;
    float sa[8], sq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sa[j] = sq[j] = 0.f;
    int st_n = -1;
    // plain mode: set `eset` takes every other tile of the CTA's sequence and always drains accumulator stage `eset`
    // (the MMA warp alternates stages tile by tile); upsample mode: the single set alternates stages itself.
    const int kSets = kUp ? 1 : p.epi_sets;
    if (kSets == 2) acc = eset;
    for (int tile = tile_lo + (kSets == 2 ? eset * tile_step : 0); tile < tile_hi; tile += tile_step * kSets) {
      const TileCoord t = decode_tile(p, tile, crank);
      if (st_on && p.stats_mode == 1 && st_n >= 0 && st_n != t.n) {
        if (st_regs) stats_round(st_slice, sa, sq, cpr_shift, lane, 0);
        stats_flush(p, st_set, st_et, st_n, 1 + eset);
      }
      st_n = t.n;
      const int h = t.h0 + g, w = t.w0 + half * 8 + r;
      const size_t pix = ((size_t)t.n * p.H + h) * p.W + w;
      const float nz = p.noise != nullptr ? p.noise[pix] : 0.f;
      // first pixel of this warp's 4 image rows x 8 columns block (row index 0 of the transpose stage); in tconv4 mode
      // the tile's pixels are every other pixel (phase py, px) of the twice-as-large output map
      const int os = p.tconv4 ? 2 : 1;
      const size_t out_w = (size_t)p.W * os;
      const size_t pix_q0 = ((size_t)t.n * p.H * os + (size_t)(t.h0 + q * 4) * os + (t.ph >> 1)) * out_w +
                            (size_t)(t.w0 + half * 8) * os + (t.ph & 1);
      const size_t opix_pool = ((size_t)t.n * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
      // per-sample, per-border-class bias row (AdaIN shift folded through the conv), else the plain bias in smem
      const int cls = (h == 0 ? 0 : (h == p.H - 1 ? 2 : 1)) * 3 + (w == 0 ? 0 : (w == p.W - 1 ? 2 : 1));
      const float* btab = p.bias_tab != nullptr ? p.bias_tab + ((size_t)t.n * 9 + cls) * p.Cout + t.co0 : nullptr;
      const bool staged = !p.pool;
      // line-wise load of the gate tile of round c0 into the (warp-private) stage: lane -> (row = lane / cpr + i * 32 / cpr,
      // chunk = lane % cpr)
      // Line-wise walks over the warp's 32 pixel rows (gate staging below, global stores further down): lane -> rows
      // r0 + i * step.  The byte offset of row r in the global map is (r >> 3) * row_bytes + (r & 7) * col_bytes; with the
      // stage width known at compile time the walk is a chain of ADDS of two precomputed strides (the 64-bit index
      // arithmetic of these two loops was ~13 % of all instructions, profiles/r2_ncu_pool4_dgrad_lines.txt).
      const size_t row_bytes = (size_t)os * out_w * p.Cout * 2, col_bytes = (size_t)os * p.Cout * 2;
      const int line_r0 = lane >> cpr_shift;
      const size_t line_off0 = ((pix_q0 * p.Cout + t.co0 + (lane & ((1 << cpr_shift) - 1)) * 8) * 2) +
                               (size_t)(line_r0 >> 3) * row_bytes + (size_t)(line_r0 & 7) * col_bytes;
      auto line_step = [&](int i) -> size_t {     // offset of row i + 1 minus offset of row i
        if (CPRS == 3) return (i & 1) ? row_bytes - 4 * col_bytes : 4 * col_bytes;
        if (CPRS == 2) return row_bytes;
        return 2 * row_bytes;
      };
      auto stage_gate = [&](int c0) {
        const int cpr = 1 << cpr_shift;
        const int ch = lane & (cpr - 1);
        if (CPRS >= 0) {
          const char* gptr = reinterpret_cast<const char*>(p.gate_src) + line_off0 + (size_t)c0 * 2;
#pragma unroll
          for (int i = 0; i < (1 << (CPRS >= 0 ? CPRS : 0)); ++i) {
            const int row = line_r0 + i * (32 >> cpr_shift);
            const uint4 gv = *reinterpret_cast<const uint4*>(gptr);
            const uint32_t swz = (uint32_t)(row >> (3 - cpr_shift)) & (uint32_t)(cpr - 1);
            sts128(stg + (uint32_t)row * rb + (((uint32_t)ch ^ swz) << 4), gv);
            gptr += line_step(i);
          }
        } else {
#pragma unroll 4
          for (int row = lane >> cpr_shift; row < 32; row += 32 >> cpr_shift) {
            const size_t gp = pix_q0 + (size_t)(row >> 3) * os * out_w + (size_t)(row & 7) * os;
            const uint4 gv = *reinterpret_cast<const uint4*>(p.gate_src + gp * p.Cout + t.co0 + c0 + ch * 8);
            const uint32_t swz = (uint32_t)(row >> (3 - cpr_shift)) & (uint32_t)(cpr - 1);
            sts128(stg + (uint32_t)row * rb + (((uint32_t)ch ^ swz) << 4), gv);
          }
        }
        __syncwarp();
      };
      // the first round's gate tile is fetched BEFORE waiting for the accumulator: its DRAM latency (14 % of the
      // epilogue warps' time in profiles/r2_ncu_pool4_dgrad_stalls.txt) then overlaps the tile's MMAs
      if (staged && p.gate_src != nullptr && !(p.debug & 8)) stage_gate(0);
      // one warp of the set polls the accumulator barrier, the other seven block in a named barrier: eight warps
      // retrying try_wait cost 15 % of all issued instructions (profiles/r2_ncu_pool4_dgrad_lines.txt)
      if (((warp - 2) & 7) == 0) mbar_wait(&tmem_full[acc], acc_phase);
      asm volatile("bar.sync %0, %1;" ::"r"(3 + eset), "n"(32 * kEpiWarps) : "memory");
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u + (uint32_t)half * 128u;
      const int ncols = (p.debug & 8) ? 0 : p.block_n;
      // dgrad / tangent passes have neither bias nor noise: skip the 16 shared-memory loads and 64 adds per 64 channels
      const bool has_add = btab != nullptr || p.bias != nullptr || p.noise != nullptr;
      for (int c0 = 0; c0 < ncols; c0 += round_cols) {
        if (c0 > 0 && staged && p.gate_src != nullptr) stage_gate(c0);
        for (int c = c0; c < c0 + round_cols; c += 16) {
          uint32_t v[16];
          tmem_ld_x16(taddr + c, v);
          float bn[16];
          if (has_add) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 b4 = btab != nullptr ? __ldg(reinterpret_cast<const float4*>(btab + c + 4 * j4))
                                                : lds128f(bias_s + t.co0 + c + 4 * j4);
              bn[4 * j4 + 0] = b4.x; bn[4 * j4 + 1] = b4.y; bn[4 * j4 + 2] = b4.z; bn[4 * j4 + 3] = b4.w;
            }
          }
          if (p.noise != nullptr) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 n4 = lds128f(nw_s + t.co0 + c + 4 * j4);
              bn[4 * j4 + 0] = fmaf(n4.x, nz, bn[4 * j4 + 0]);
              bn[4 * j4 + 1] = fmaf(n4.y, nz, bn[4 * j4 + 1]);
              bn[4 * j4 + 2] = fmaf(n4.z, nz, bn[4 * j4 + 2]);
              bn[4 * j4 + 3] = fmaf(n4.w, nz, bn[4 * j4 + 3]);
            }
          }
          const uint32_t chunk0 = (uint32_t)((c - c0) >> 3);            // first of this thread's two 16-byte chunks
          const uint32_t sa0 = my_row + (((chunk0) ^ my_swz) << 4);
          const uint32_t sa1 = my_row + (((chunk0 + 1u) ^ my_swz) << 4);
          uint4 ga = make_uint4(0, 0, 0, 0), gb = ga;
          if (p.gate_src != nullptr) {
            if (staged) {
              ga = lds128(sa0);
              gb = lds128(sa1);
            } else if (pool_writer) {
              const uint4* g4 = reinterpret_cast<const uint4*>(p.gate_src + opix_pool * p.Cout + t.co0 + c);
              ga = g4[0];
              gb = g4[1];
            }
          }
          tmem_ld_wait();
          float f[16];
          if (has_add) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + bn[j];
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
          }
          if (p.pool) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float sum = f[j] + __shfl_xor_sync(0xffffffffu, f[j], 1);
              sum += __shfl_xor_sync(0xffffffffu, sum, 8);
              f[j] = 0.25f * sum;
            }
          }
          if (p.act) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], f[j] * p.slope);      // LeakyReLU, 0 < slope < 1
          }
          uint4 o0, o1;
          if (p.gate_src != nullptr) {
            // LeakyReLU backward gate (1 where the saved activation is > 0, slope elsewhere) on PACKED pairs: both
            // candidates f and slope * f are rounded to bf16x2 and one bf16x2 compare mask (set.gt.bf16x2) picks per
            // half — 2.5 instructions per element instead of 4.5 (unpack, FSETP, FSEL, FMUL), same values bit for bit
            const uint32_t gw[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
            uint32_t ow[8];
            const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&gw[j]), zero2);
              const uint32_t keep = pack_bf16x2(f[2 * j], f[2 * j + 1]);
              const uint32_t scaled = pack_bf16x2(f[2 * j] * p.slope, f[2 * j + 1] * p.slope);
              ow[j] = (keep & m) | (scaled & ~m);
            }
            o0 = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            o1 = make_uint4(ow[4], ow[5], ow[6], ow[7]);
          } else {
            o0.x = pack_bf16x2(f[0], f[1]);
            o0.y = pack_bf16x2(f[2], f[3]);
            o0.z = pack_bf16x2(f[4], f[5]);
            o0.w = pack_bf16x2(f[6], f[7]);
            o1.x = pack_bf16x2(f[8], f[9]);
            o1.y = pack_bf16x2(f[10], f[11]);
            o1.z = pack_bf16x2(f[12], f[13]);
            o1.w = pack_bf16x2(f[14], f[15]);
          }
          if (staged) {
            sts128(sa0, o0);
            sts128(sa1, o1);
          } else if (pool_writer && !(p.debug & 1)) {
            uint4* o4 = reinterpret_cast<uint4*>(p.out + opix_pool * p.Cout + t.co0 + c);
            o4[0] = o0;
            o4[1] = o1;
          }
        }
        if (staged) {
          __syncwarp();
          if (!(p.debug & 1)) {
            const int cpr = 1 << cpr_shift;
            const int ch = lane & (cpr - 1);
            char* optr = reinterpret_cast<char*>(p.out) + line_off0 + (size_t)c0 * 2;
#pragma unroll 4
            for (int i = 0; i < cpr; ++i) {
              const int row = line_r0 + i * (32 >> cpr_shift);
              const uint32_t swz = (uint32_t)(row >> (3 - cpr_shift)) & (uint32_t)(cpr - 1);
              const uint4 ov = lds128(stg + (uint32_t)row * rb + (((uint32_t)ch ^ swz) << 4));
              if (CPRS >= 0) {
                *reinterpret_cast<uint4*>(optr) = ov;
                optr += line_step(i);
              } else {
                const size_t op = pix_q0 + (size_t)(row >> 3) * os * out_w + (size_t)(row & 7) * os;
                *reinterpret_cast<uint4*>(p.out + op * p.Cout + t.co0 + c0 + ch * 8) = ov;
              }
              if (st_on) {
                const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f2 = unpack_bf16x2(ow[j]);
                  sa[2 * j] += f2.x;
                  sa[2 * j + 1] += f2.y;
                  sq[2 * j] = fmaf(f2.x, f2.x, sq[2 * j]);
                  sq[2 * j + 1] = fmaf(f2.y, f2.y, sq[2 * j + 1]);
                }
              }
            }
          }
          if (st_on && !st_regs) stats_round(st_slice, sa, sq, cpr_shift, lane, t.co0 + c0);
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCta2 && crank != 0) mbar_arrive_cluster(mapa_cta(smem_u32(&tmem_empty[acc]), 0u));   // the leader's MMA warps wait
        else mbar_arrive(&tmem_empty[acc]);
      }
      if (kSets == 2) {
        acc_phase ^= 1u;
      } else {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    if (st_on && st_n >= 0) {
      if (st_regs) stats_round(st_slice, sa, sq, cpr_shift, lane, 0);
      stats_flush(p, st_set, st_et, st_n, 1 + eset);
    }
}

template <bool kStats, int kFeed, bool kCta2>     // kFeed: 0 TMA halo (plain and pool4), 1 upsample producers
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                 const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  // carve: [B region (1024-aligned tiles)] [A stages] [epilogue transpose stages] [aux]
  const int b_tiles = p.b_resident ? p.k_chunks * p.taps : p.b_stages;
  uint8_t* b_base = smem;
  uint8_t* a_base = b_base + (size_t)b_tiles * p.b_tile_bytes;
  uint8_t* patch_base = a_base + (size_t)p.a_stages * p.a_stage_bytes;
  patch_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(patch_base) + 127) & ~uintptr_t(127));
  uint8_t* epi_base = patch_base + (p.upsample ? (size_t)kPatchStages * p.patch_bytes : 0);
  epi_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(epi_base) + 127) & ~uintptr_t(127));
  constexpr bool kUp = kFeed != 0;        // warps 10..17 feed the halo instead of being the second epilogue set
  const int kEpiWarpsAll = kUp ? kEpiWarps : kEpiWarps * p.epi_sets;
  uint8_t* aux = epi_base + (size_t)kEpiWarpsAll * 32 * p.epi_row_bytes;
  uint64_t* b_full = reinterpret_cast<uint64_t*>(aux);
  uint64_t* b_empty = b_full + kMaxBStages;
  uint64_t* a_full = b_empty + kMaxBStages;
  uint64_t* a_empty = a_full + kMaxAStages;
  uint64_t* tmem_full = a_empty + kMaxAStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);
  float* nw_s = bias_s + 512;
  float* stat_s = nw_s + 512;             // [kEpiWarps][Cout][2], only carved when stats_mode != 0

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform role index
  const int lane = threadIdx.x & 31;
  const uint32_t crank = kCta2 ? cluster_ctarank() : 0u;                   // 0 = leader of the CTA pair

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < kMaxBStages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], (uint32_t)p.issuers);
    }
    for (int s = 0; s < kMaxAStages; ++s) {
      mbar_init(&a_full[s], kUp ? kProdThreads : 1);
      mbar_init(&a_empty[s], (uint32_t)p.issuers);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], (uint32_t)p.issuers);
      mbar_init(&tmem_empty[s], kCta2 ? 2 * kEpiWarps : kEpiWarps);   // pair mode: the leader's copy counts both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kCta2) {
      tmem_alloc2(tmem_slot, kTmemCols);        // one warp of EACH CTA of the pair takes part
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
    // only now may the next kernel of the stream become resident: released at kernel entry, a dependent CTA that landed
    // on this SM could take the TMEM columns first and then wait for this grid, which would be waiting for the columns
    pdl_launch_dependents();
  }
  pdl_wait();        // barrier init, TMEM allocation and descriptor prefetch above overlap the previous kernel's tail
  if (warp >= 2 && warp < 2 + kEpiWarpsAll) {
    for (int c = threadIdx.x - 64; c < p.Cout; c += 32 * kEpiWarpsAll) {
      bias_s[c] = p.bias ? p.bias[c] : 0.f;
      nw_s[c] = p.noise_w ? p.noise_w[c] : 0.f;
    }
    if (kStats)
      for (int i = threadIdx.x - 64; i < kEpiWarpsAll * p.Cout * 2; i += 32 * kEpiWarpsAll) stat_s[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (kCta2) cluster_sync_all();     // the peer's barriers are initialised and its TMEM allocated before anything remote
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  int tile_lo, tile_hi, tile_step;
  tile_range(p, tile_lo, tile_hi, tile_step);

  if (warp == 0) {
    // ------------------------------ weight TMA producer ------------------------------
    if (lane == 0) {
      if (p.b_resident) {
        // residency is only selected when there is a single n-block: the whole pack is loaded once per kernel, or,
        // with per-sample weights, once per sample of this CTA's (contiguous) tile range — after the MMA warp has
        // committed the last tile that used the previous sample's pack (b_empty[0]).
        int cur = -1;
        uint32_t loads = 0;
        for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
          const int wn = p.per_sample_w ? decode_tile(p, tile).n : 0;
          if (wn == cur) continue;
          if (loads > 0) mbar_wait(&b_empty[0], (loads - 1u) & 1u);
          // pair mode: this CTA keeps rows [crank * block_n / 2, ...) of every tile; all bytes count on the leader's barrier
          if (!kCta2 || crank == 0)
            mbar_expect_tx(&b_full[0], (uint32_t)(p.k_chunks * p.taps) * p.b_tx_bytes * (kCta2 ? 2u : 1u));
          const uint32_t full_addr = kCta2 ? mapa_cta(smem_u32(&b_full[0]), 0u) : 0u;
          for (int kcx = 0; kcx < p.k_chunks; ++kcx)
            for (int tap = 0; tap < p.taps; ++tap) {
              uint8_t* dst = b_base + (size_t)(kcx * p.taps + tap) * p.b_tile_bytes;
              if (kCta2) tma_load_4d_2sm(&tmap_w, full_addr, dst, kcx * p.kc, (int)crank * (p.block_n >> 1), tap, wn);
              else tma_load_4d(&tmap_w, &b_full[0], dst, kcx * p.kc, 0, tap, wn);
            }
          cur = wn;
          ++loads;
        }
      } else {
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
          const TileCoord t = decode_tile(p, tile);
          for (int kcx = 0; kcx < p.k_chunks; ++kcx) {
            const int ntap = p.tconv4 ? 4 : p.taps;
            for (int ti = 0; ti < ntap; ++ti) {
              // the order the MMA warp consumes them
              const int tap = p.tconv4 ? t.ph * 4 + ti : (p.pool4_tma ? pool4_tap(ti >> 2, ti & 3) : ti);
              mbar_wait(&b_empty[stage], phase ^ 1u);
              if (kCta2) {
                if (crank == 0) mbar_expect_tx(&b_full[stage], 2u * p.b_tx_bytes);
                tma_load_4d_2sm(&tmap_w, mapa_cta(smem_u32(&b_full[stage]), 0u), b_base + (size_t)stage * p.b_tile_bytes,
                                kcx * p.kc, t.co0 + (int)crank * (p.block_n >> 1), tap, p.per_sample_w ? t.n : 0);
              } else {
                mbar_expect_tx(&b_full[stage], p.b_tx_bytes);
                tma_load_4d(&tmap_w, &b_full[stage], b_base + (size_t)stage * p.b_tile_bytes, kcx * p.kc, t.co0, tap,
                            p.per_sample_w ? t.n : 0);
              }
              if (++stage == p.b_stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1 || (warp == kIssuer2 && p.issuers == 2)) {
    // ------------------------------ MMA issuer(s) ------------------------------
    if (kCta2 && crank != 0) {
      // the leader CTA issues the pair's instructions; this CTA's MMA warps have nothing to do
    } else if (p.issuers == 2) {
      if (warp == 1) issue_dispatch<kStats, kFeed, 0, kCta2>(p, a_base, b_base, b_full, b_empty, a_full, a_empty, tmem_full, tmem_empty, tmem_base);
      else issue_dispatch<kStats, kFeed, 1, kCta2>(p, a_base, b_base, b_full, b_empty, a_full, a_empty, tmem_full, tmem_empty, tmem_base);
    } else {
      issue_dispatch<kStats, kFeed, -1, kCta2>(p, a_base, b_base, b_full, b_empty, a_full, a_empty, tmem_full, tmem_empty, tmem_base);
    }
  } else if (warp < 2 + kEpiWarpsAll) {
    // ------------------------------ epilogue (epilogue_warp above) ------------------------------
    if (p.epi_templated && p.epi_row_bytes == 128u)
      epilogue_warp<kStats, kUp, kCta2, 3>(p, warp, lane, crank, tmem_base, tmem_full, tmem_empty, epi_base, bias_s, nw_s, stat_s, tile_lo, tile_hi, tile_step);
    else if (p.epi_templated && p.epi_row_bytes == 64u)
      epilogue_warp<kStats, kUp, kCta2, 2>(p, warp, lane, crank, tmem_base, tmem_full, tmem_empty, epi_base, bias_s, nw_s, stat_s, tile_lo, tile_hi, tile_step);
    else if (p.epi_templated && p.epi_row_bytes == 32u)
      epilogue_warp<kStats, kUp, kCta2, 1>(p, warp, lane, crank, tmem_base, tmem_full, tmem_empty, epi_base, bias_s, nw_s, stat_s, tile_lo, tile_hi, tile_step);
    else
      epilogue_warp<kStats, kUp, kCta2, -1>(p, warp, lane, crank, tmem_base, tmem_full, tmem_empty, epi_base, bias_s, nw_s, stat_s, tile_lo, tile_hi, tile_step);
  } else {
    // ------------------------------ halo producers (cp.async, zero-fill padding) ------------------------------
    if (kUp) {
      if (warp < 2 + kEpiWarps + kProdWarps) {
        const int pt = threadIdx.x - 32 * (2 + kEpiWarps);       // 0..255
        if (p.cpp_shift == 3) halo_producer_upsample<3>(p, a_base, patch_base, a_full, a_empty, pt);
        else if (p.cpp_shift == 2) halo_producer_upsample<2>(p, a_base, patch_base, a_full, a_empty, pt);
        else halo_producer_upsample<1>(p, a_base, patch_base, a_full, a_empty, pt);
      }
    } else if (warp == 2 + kEpiWarps * kEpiSets && lane == 0) {
      halo_tma_loop<kCta2>(p, &tmap_x, a_base, a_full, a_empty, crank);        // warp 18
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kCta2) cluster_sync_all();     // the peer's shared memory / barriers stay valid until both CTAs are done with them
  if (warp == 1) {
    tc_fence_after();
    if (kCta2) tmem_dealloc2(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

int ilog2(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return s;
}

int pick_block_n(int Cout) {
  for (int c = kMaxN; c >= 16; c -= 16)
    if (Cout % c == 0) return c;
  return 0;
}

}  // namespace

bool conv_halo_supported(int N, int H, int W, int Cin, int Cout, int ksize) {
  if (!(ksize == 3 && H >= 16 && W >= 16 && (H & (H - 1)) == 0 && (W & (W - 1)) == 0 && Cin % 16 == 0 &&
        Cout % 16 == 0 && Cout <= 512 && N > 0))
    return false;
  const int bn = pick_block_n(Cout);
  if (bn == 0) return false;
  const int nb = Cout / bn;
  if ((nb & (nb - 1)) != 0) return false;                    // tile decode uses shifts
  if (!(bn == 16 || bn == 32 || bn % 64 == 0)) return false; // epilogue transpose rounds are 16 / 32 / 64 columns
  return true;
}

// Host launcher; same arguments as launch_conv_fprop (ksize must be 3) plus the fused-pool switch.
int launch_conv_halo(const void* x, const void* wpack, void* out, int N, int H, int W, int Cin, int Cout,
                     const float* bias, const float* noise, const float* noise_w, const void* gate_src, int act,
                     int pool, float slope, float* stats, int stats_mode, const float* bias_tab, int per_sample_w,
                     int upsample, cudaStream_t stream) {
  // pool: 0 none, 1 conv3x3 then 2x2 average in the epilogue, 2 "pool4": the same function as ONE 4x4 stride-2 conv
  // (wpack is then the 16-tap pack of bg_pack_weight_pool4).  H, W are always the conv INPUT's.
  // 3 "tconv4": the input-gradient of pool4 (x = pooled gradient (N,H,W,Cin), out = (N,2H,2W,Cout), wpack = the
  // 16-tile pack of bg_pack_weight_tconv4).
  const bool pool4 = pool == 2;
  const bool tconv4 = pool == 3;
  if (tconv4) {
    BG_REQUIRE(!upsample && !per_sample_w && noise == nullptr && bias == nullptr && (stats == nullptr || stats_mode == 2),
               "conv_halo tconv4: no upsample / per-sample weights / noise / bias; stats_mode 2 only");
    pool = 0;
  }
  if (pool4) {
    BG_REQUIRE(H >= 32 && W >= 32 && Cin % 32 == 0 && !upsample && stats == nullptr && !per_sample_w && noise == nullptr,
               "conv_halo pool4: needs H,W >= 32, Cin %% 32 == 0 and no upsample / stats / noise (H %d W %d Cin %d)", H, W, Cin);
    pool = 0;
  }
  const int Hc = pool4 ? H / 2 : H, Wc = pool4 ? W / 2 : W;      // the map the tiles and the epilogue work on
  BG_REQUIRE(conv_halo_supported(N, Hc, Wc, Cin, Cout, 3), "conv_halo: unsupported shape N %d H %d W %d Cin %d Cout %d", N,
             H, W, Cin, Cout);
  BG_REQUIRE(!(upsample && pool), "conv_halo: upsample and pool cannot be combined");
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = Hc; p.W = Wc; p.Cin = Cin; p.Cout = Cout;
  p.pool4 = pool4 ? 1 : 0;
  p.tconv4 = tconv4 ? 1 : 0;
  p.ph_shift = tconv4 ? 2 : 0;
  p.taps = (pool4 || tconv4) ? 16 : 9;
  p.Hin = H; p.Win = W;
  p.tw_shift = ilog2(Wc / kTile);
  p.th_shift = ilog2(Hc / kTile);
  const int bn_ch = pick_block_n(Cout);
  p.block_n = bn_ch;
  const int n_blocks = Cout / bn_ch;
  p.nb_shift = ilog2(n_blocks);
  p.pool4_tma = pool4 ? 1 : 0;
  p.kc = (Cin % 64 == 0) ? 64 : (Cin % 32 == 0 ? 32 : 16);
  p.k_chunks = Cin / p.kc;
  p.cpp_shift = p.kc == 64 ? 3 : (p.kc == 32 ? 2 : 1);
  const uint32_t row_bytes = (uint32_t)p.kc * 2u;
  p.b_layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  p.b_sbo = 8u * row_bytes;
  {
    static int cta2_on = -1;                                  // BG_CTA2=0: every layer with one CTA per tile (A/B switch)
    if (cta2_on < 0) { const char* e = getenv("BG_CTA2"); cta2_on = (e && e[0] == '0') ? 0 : 1; }
    static int cta2_min_n = -2;                               // BG_CTA2_MIN_N=n: pair every tile >= n wide (experiments)
    if (cta2_min_n == -2) { const char* e = getenv("BG_CTA2_MIN_N"); cta2_min_n = e ? atoi(e) : -1; }
    // pair mode (TMA halo feed, at least two tile columns) where halving the weight operand pays: N = 128 per
    // instruction, and N = 64 with >= 128 input channels (measured on B200, 32 x 128^2: 128->64 88 -> 75 us; the
    // 32/64-channel-input layers at 256^2 lose 15-25 % in pairs, profiles/r2_pair_mode_narrow_layers.txt)
    const bool wide = cta2_min_n >= 0 ? (bn_ch >= cta2_min_n && bn_ch >= 32) : (bn_ch == 128 || (bn_ch == 64 && Cin >= 128));
    p.cta2 = (cta2_on && wide && !upsample && (Wc / kTile) >= 2) ? 1 : 0;
  }
  p.b_tx_bytes = (uint32_t)(p.block_n >> p.cta2) * row_bytes;     // pair mode: each CTA holds half of the weight tile
  p.b_tile_bytes = (p.b_tx_bytes + 1023u) & ~1023u;
  p.a_layout = p.b_layout;                                   // same row width (kc bf16) on both operands
  p.a_sbo = (uint32_t)(pool4 ? kTile + 1 : kHalo) * row_bytes;   // next 8-pixel group = one (phase-)tile row down
  p.a_tx_bytes = (uint32_t)kHaloPix * row_bytes;
  p.phase_bytes = ((uint32_t)((kTile + 1) * (kTile + 1)) * row_bytes + 1023u) & ~1023u;
  p.a_stage_bytes = pool4 ? 4u * p.phase_bytes : ((p.a_tx_bytes + 1023u) & ~1023u);
  p.epi_row_bytes = (uint32_t)(bn_ch >= 64 ? 64 : bn_ch) * 2u;
  p.upsample = upsample ? 1 : 0;
  {
    static int up_packed = -1;
    if (up_packed < 0) { const char* e = getenv("BG_UP_PACKED"); up_packed = (e && e[0] == '0') ? 0 : 1; }
    p.up_packed = up_packed;
  }
  if (p.pool4_tma) {
    // the pipeline unit is one phase tile (4 per chunk); two epilogue sets: keep the transpose stages at 32 KB
    p.a_tx_bytes = (uint32_t)((kTile + 1) * (kTile + 1)) * row_bytes;
    p.a_stage_bytes = p.phase_bytes;
    if (p.epi_row_bytes > 64u) p.epi_row_bytes = 64u;
  }
  const bool feed_warps = p.upsample != 0;                   // warps 10..17 build the halo: one epilogue set
  p.patch_bytes = (uint32_t)(kPatch * kPatch) * (uint32_t)p.kc * 2u;
  const uint32_t patch_total = p.upsample ? (uint32_t)kPatchStages * p.patch_bytes + 128u : 0u;
  const uint32_t resident_bytes = (uint32_t)(p.k_chunks * p.taps) * p.b_tile_bytes;
  // Shared-memory plan.  Two epilogue sets (plain mode) double the transpose stages and the stats slices; fall back
  // to one set when that would cost the weight residency or does not fit at all.
  struct Plan {
    bool ok;
    int sets, a_stages, b_stages, resident;
    uint32_t epi_bytes, aux_bytes;
  };
  auto make_plan = [&](int sets) {
    Plan pl;
    memset(&pl, 0, sizeof(pl));
    pl.sets = sets;
    const uint32_t epi_warps = (uint32_t)(kEpiWarps * sets);
    pl.epi_bytes = epi_warps * 32u * p.epi_row_bytes + 128u;
    const uint32_t stat_bytes = (stats != nullptr) ? epi_warps * (uint32_t)Cout * 2u * 4u : 0u;
    pl.aux_bytes = 8 * (2 * kMaxBStages + 2 * kMaxAStages + 4) + 16 + 2 * 512 * 4 + 64 + stat_bytes;
    const uint32_t fixed = 1024u + pl.aux_bytes + pl.epi_bytes + patch_total;
    if (fixed + 2u * p.a_stage_bytes + 2u * p.b_tile_bytes > 227u * 1024u) return pl;
    const uint32_t total = 227u * 1024u - fixed;
    // prefer resident weights (3 halo stages if they fit, else 2); otherwise stream the weights past 3 halo stages
    pl.a_stages = 3;
    if (n_blocks == 1 && !p.pool4_tma) {
      if (resident_bytes + 3u * p.a_stage_bytes <= total) {
        pl.resident = 1;
        // small stages (kc = 16 / 32): more stages, until ~96 KB of halo loads can be in flight (HBM latency x per-SM
        // bandwidth), as far as shared memory allows
        if (!feed_warps) {
          while (pl.a_stages < kMaxAStages && (uint32_t)(pl.a_stages - 1) * p.a_stage_bytes < 96u * 1024u &&
                 resident_bytes + (uint32_t)(pl.a_stages + 1) * p.a_stage_bytes <= total)
            ++pl.a_stages;
        }
      } else if (resident_bytes + 2u * p.a_stage_bytes <= total) {
        pl.resident = 1;
        pl.a_stages = 2;
      }
    }
    if (pl.resident) {
      pl.b_stages = 1;
      pl.ok = true;
    } else {
      // streamed weights: 3 halo stages when at least 3 weight stages still fit, else 2 halo stages
      if (p.pool4_tma) pl.a_stages = 4;                      // phase-sized stages: one whole chunk in flight
      int st = total > (uint32_t)pl.a_stages * p.a_stage_bytes
                   ? (int)((total - (uint32_t)pl.a_stages * p.a_stage_bytes) / p.b_tile_bytes) : 0;
      while (st < 3 && pl.a_stages > 2) {
        --pl.a_stages;
        st = total > (uint32_t)pl.a_stages * p.a_stage_bytes
                 ? (int)((total - (uint32_t)pl.a_stages * p.a_stage_bytes) / p.b_tile_bytes) : 0;
      }
      if (st > kMaxBStages) st = kMaxBStages;
      pl.b_stages = st;
      pl.ok = st >= 2;
    }
    return pl;
  };
  Plan plan = make_plan(1);
  if (!feed_warps) {
    const Plan two = make_plan(kEpiSets);
    if (two.ok && (two.resident || !plan.resident || !plan.ok)) plan = two;
  }
  const bool planned = plan.ok;
  const uint32_t epi_bytes = plan.epi_bytes, aux_bytes = plan.aux_bytes;
  p.epi_sets = plan.sets;
  {
    static int issuers = 0;                                     // BG_MMA_ISSUERS=1: the single-thread issue of round-1's first kernels
    if (issuers == 0) { const char* e = getenv("BG_MMA_ISSUERS"); issuers = (e && e[0] == '1') ? 1 : 2; }
    p.issuers = issuers;
  }
  {
    static int epi_t = -1;
    if (epi_t < 0) { const char* e = getenv("BG_EPI_TEMPLATED"); epi_t = (e && e[0] == '0') ? 0 : 1; }
    p.epi_templated = epi_t;
  }
  p.a_stages = plan.a_stages;
  p.b_stages = plan.b_stages;
  p.b_resident = plan.resident;
  BG_REQUIRE(planned, "conv_halo: weight tile does not fit shared memory");
  p.num_tiles = (Wc / kTile) * (Hc / kTile) * N * n_blocks * (tconv4 ? 4 : 1);
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.bias = bias; p.noise = noise; p.noise_w = noise_w;
  p.gate_src = reinterpret_cast<const __nv_bfloat16*>(gate_src);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.act = act; p.pool = pool; p.slope = slope;
  p.n_blocks = n_blocks;
  p.stats = stats;
  p.stats_mode = stats != nullptr ? stats_mode : 0;
  p.per_sample_w = per_sample_w ? 1 : 0;
  p.bias_tab = bias_tab;
  p.contig = (p.stats_mode != 0 || p.per_sample_w) ? 1 : 0;
  BG_REQUIRE(p.stats_mode == 0 || ((stats_mode == 1 || stats_mode == 2) && !pool),
             "conv_halo: stats_mode must be 1 or 2 and cannot be combined with the fused pool");
  if (p.stats_mode)
    if (launch_zero(stats, (stats_mode == 1 ? (size_t)N * Cout * 2 : (size_t)Cout) * sizeof(float), stream) != 0) return 1;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("BG_HALO_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  if ((p.debug & 16) && !p.per_sample_w) p.contig ^= 1;       // A/B the tile order

  CUtensorMap tmw;
  {
    // [sample][tap][Cout][Cin]; the sample dimension has extent 1 for the ordinary shared pack
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)p.taps, (uint64_t)(p.per_sample_w ? N : 1)};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2, (uint64_t)p.taps * Cout * Cin * 2};
    uint32_t box[4] = {(uint32_t)p.kc, (uint32_t)(p.block_n >> p.cta2), 1u, 1u};
    if (make_tmap_bf16(&tmw, wpack, 4, dims, str, box, (int)row_bytes) != 0) return 1;
  }
  CUtensorMap tmx;
  {
    // plain mode: NHWC input as (Cin, W, H, N); the halo box starts at (w0-1, h0-1) and relies on OOB zero fill.
    // (upsample mode builds the halo with the producer warps; the map is still encoded so the kernel signature is one)
    const int Hx = upsample ? H / 2 : H, Wx = upsample ? W / 2 : W;
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Wx, (uint64_t)Hx, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)Wx * Cin * 2, (uint64_t)Hx * Wx * Cin * 2};
    uint32_t box[4] = {(uint32_t)p.kc, (uint32_t)(feed_warps ? 8 : kHalo), (uint32_t)(feed_warps ? 8 : kHalo), 1u};
    if (p.pool4_tma) {
      // every other pixel of a 34-wide span -> 17 x 17 pixels per phase tile
      uint32_t box2[4] = {(uint32_t)p.kc, 2u * (kTile + 1), 2u * (kTile + 1), 1u};
      uint32_t est[4] = {1u, 2u, 2u, 1u};
      if (make_tmap_bf16_strided(&tmx, x, 4, dims, str, box2, est, (int)row_bytes) != 0) return 1;
    } else if (make_tmap_bf16(&tmx, x, 4, dims, str, box, (int)row_bytes) != 0) return 1;
  }
  const size_t b_tiles = p.b_resident ? (size_t)p.k_chunks * p.taps : (size_t)p.b_stages;
  const size_t smem_bytes = b_tiles * p.b_tile_bytes + (size_t)p.a_stages * p.a_stage_bytes + patch_total + epi_bytes +
                            aux_bytes + 1024;
  BG_REQUIRE(smem_bytes <= 227 * 1024, "conv_halo: shared memory budget exceeded (%zu)", smem_bytes);
  static bool attr_set = false;
  if (!attr_set) {
    BG_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<false, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BG_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<true, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BG_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<false, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BG_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<true, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BG_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<false, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BG_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<true, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (p.cta2) {
    // one CTA pair (2-CTA cluster) per tile pair; persistent over at most one pair per two SMs
    const int pairs_total = p.num_tiles / 2;
    const int pairs = pairs_total < num_sms() / 2 ? pairs_total : num_sms() / 2;
    if (p.stats_mode) BG_CHECK_CUDA(launch_pdl_cluster(conv_halo_kernel<true, 0, true>, 2 * pairs, kThreads, smem_bytes, stream, 2, tmw, tmx, p));
    else BG_CHECK_CUDA(launch_pdl_cluster(conv_halo_kernel<false, 0, true>, 2 * pairs, kThreads, smem_bytes, stream, 2, tmw, tmx, p));
    BG_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  if (p.stats_mode && p.upsample) BG_CHECK_CUDA(launch_pdl(conv_halo_kernel<true, 1, false>, grid, kThreads, smem_bytes, stream, tmw, tmx, p));
  else if (p.stats_mode) BG_CHECK_CUDA(launch_pdl(conv_halo_kernel<true, 0, false>, grid, kThreads, smem_bytes, stream, tmw, tmx, p));
  else if (p.upsample) BG_CHECK_CUDA(launch_pdl(conv_halo_kernel<false, 1, false>, grid, kThreads, smem_bytes, stream, tmw, tmx, p));
  else BG_CHECK_CUDA(launch_pdl(conv_halo_kernel<false, 0, false>, grid, kThreads, smem_bytes, stream, tmw, tmx, p));
  BG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace bg
