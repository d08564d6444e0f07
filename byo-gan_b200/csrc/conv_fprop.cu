// Implicit-GEMM convolution (3x3 pad 1, or 1x1) on tcgen05 tensor cores, NHWC bf16 -> NHWC bf16.
//
// Replaces the reference's F.conv2d call in EqualizedConv2d.forward (gan.py:29-38) for every 3x3
// layer of the generator (gan.py:83) and the critic (gan.py:240,254,259).  The same kernel runs the
// data-gradient pass (autograd's convolution_backward w.r.t. the input) when it is handed the
// flipped/transposed weight pack, and the R1 tangent pass (the double-backward forward conv).
//
// GEMM view:   D[pixel, co] = sum_{tap, ci} X[pixel + tap, ci] * Wp[tap][co][ci]
//   M tile = 128 output pixels = a (bw x bh x bn) box of (w, h, n); the A operand for a tap is ONE
//            TMA tiled-box load of the input shifted by the tap offset, zero padding comes from the
//            TMA out-of-bounds fill, so there is no im2col buffer and no halo bookkeeping;
//   N tile = block_n output channels (<= 256, multiple of 16);
//   K loop = (Cin / kc) channel chunks x taps, kc in {64, 32, 16} -> 128B / 64B / 32B swizzle.
// Accumulators live in TMEM (2 stages x 256 fp32 columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.  Persistent grid: one CTA per SM, static round-robin over tiles.
//
// Warp roles: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane) + TMEM allocator,
// warps 2..5 = epilogue (TMEM -> registers -> fused bias / noise / LeakyReLU / gate -> global).
#include "common.cuh"

#include <stdlib.h>

namespace bg {

namespace {

constexpr int kEpiWarps = 4;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kMaxStages = 8;
constexpr int kMaxCout = 640;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccStride = 256;

struct FpropParams {
  int N, H, W, Cin, Cout;
  int ksize, pad, taps;
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n;
  int block_n, n_blocks;
  int kc, k_chunks;
  int stages;
  uint32_t a_bytes, b_bytes, stage_bytes;
  uint32_t layout_type, sbo_bytes;
  int num_tiles;
  const float* bias;
  const float* noise;
  const float* noise_w;
  const __nv_bfloat16* gate_src;
  __nv_bfloat16* out;
  int act;
  float slope;
  // A-operand multicast (the small maps this kernel serves are bound by the L2 -> shared-memory fill, and 2/3 of that
  // fill is the pixel tile that every n-block of a layer re-reads): the `csize` CTAs of a cluster work on the SAME pixel
  // tile and consecutive n-blocks; each loads 1/csize of the tile (slice_rows pixel rows) and multicasts it to all.
  int csize, slice_rows, slice_n0_div, slice_h_rows;
  uint32_t slice_bytes;
};

struct TileCoord {
  int w0, h0, n0, co0;
};

__device__ __forceinline__ TileCoord decode_tile(const FpropParams& p, int tile) {
  TileCoord t;
  int nb = tile % p.n_blocks;
  int pt = tile / p.n_blocks;
  int tw = pt % p.tiles_w;
  int th = (pt / p.tiles_w) % p.tiles_h;
  int tn = pt / (p.tiles_w * p.tiles_h);
  t.w0 = tw * p.bw;
  t.h0 = th * p.bh;
  t.n0 = tn * p.bn;
  t.co0 = nb * p.block_n;
  return t;
}

template <bool kMc>
__global__ void __launch_bounds__(kThreads, 1)
conv_fprop_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  const FpropParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint8_t* aux = smem + (size_t)p.stages * p.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);
  float* nw_s = bias_s + kMaxCout;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = kMc ? cluster_ctarank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << p.csize) - 1u);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kMc ? (uint32_t)p.csize : 1u);   // multicast: every CTA of the cluster must have consumed it
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps);
    }
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
  if (warp >= 2) {
    for (int c = threadIdx.x - 64; c < p.Cout; c += 32 * kEpiWarps) {
      bias_s[c] = p.bias ? p.bias[c] : 0.f;
      nw_s[c] = p.noise_w ? p.noise_w[c] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kMc) cluster_sync_all();       // every CTA's barriers exist before a peer multicasts into it
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int k_blocks = p.k_chunks * p.taps;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        for (int kcx = 0; kcx < p.k_chunks; ++kcx) {
          for (int tap = 0; tap < p.taps; ++tap) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
            uint8_t* sb = sa + p.a_bytes;
            mbar_expect_tx(&full_bar[stage], p.a_bytes + p.b_bytes);
            const int ky = tap / p.ksize, kx = tap % p.ksize;
            if (kMc) {
              // this CTA's slice of the pixel tile: rows [crank * slice_rows, ...) in (w, h, n) order, to every CTA
              const int first = (int)crank * p.slice_rows;
              const int ni0 = first / (p.bw * p.bh), hi0 = (first / p.bw) % p.bh;
              tma_load_4d_mc(&tmap_x, &full_bar[stage], sa + (size_t)crank * p.slice_bytes, kcx * p.kc, t.w0 + kx - p.pad,
                             t.h0 + hi0 + ky - p.pad, t.n0 + ni0, cmask);
            } else {
              tma_load_4d(&tmap_x, &full_bar[stage], sa, kcx * p.kc, t.w0 + kx - p.pad, t.h0 + ky - p.pad, t.n0);
            }
            tma_load_3d(&tmap_w, &full_bar[stage], sb, kcx * p.kc, t.co0, tap);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.block_n, 0, 0);
      const int ksteps = p.kc / 16;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccStride;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
          const uint32_t sb = sa + p.a_bytes;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t adesc = umma_desc(sa + k * 32, 16, p.sbo_bytes, p.layout_type);
            const uint64_t bdesc = umma_desc(sb + k * 32, 16, p.sbo_bytes, p.layout_type);
            tc_mma_bf16(d_tmem, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          if (kMc) tc_commit_mc(&empty_bar[stage], cmask);
          else tc_commit(&empty_bar[stage]);
          if (kb == k_blocks - 1) tc_commit(&tmem_full[acc]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int wi = row % p.bw;
    const int hi = (row / p.bw) % p.bh;
    const int ni = row / (p.bw * p.bh);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int n = t.n0 + ni, h = t.h0 + hi, w = t.w0 + wi;
      const bool valid = (n < p.N) && (h < p.H) && (w < p.W);
      const size_t pix = ((size_t)n * p.H + h) * p.W + w;
      const float nz = (p.noise != nullptr && valid) ? p.noise[pix] : 0.f;
      __nv_bfloat16* orow = p.out + pix * p.Cout + t.co0;
      const __nv_bfloat16* grow = p.gate_src ? p.gate_src + pix * p.Cout + t.co0 : nullptr;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * kAccStride;
      for (int c = 0; c < p.block_n; c += 16) {
        uint32_t v[16];
        tmem_ld_x16(taddr + c, v);
        tmem_ld_wait();
        if (valid) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int co = t.co0 + c + j;
            float x = __uint_as_float(v[j]) + bias_s[co] + nw_s[co] * nz;
            if (p.act) x = x > 0.f ? x : x * p.slope;
            f[j] = x;
          }
          if (grow != nullptr) {
            const uint4* g4 = reinterpret_cast<const uint4*>(grow + c);
            uint4 ga = g4[0], gb = g4[1];
            const uint32_t gw[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 gv = unpack_bf16x2(gw[j]);
              f[2 * j] *= gv.x > 0.f ? 1.f : p.slope;
              f[2 * j + 1] *= gv.y > 0.f ? 1.f : p.slope;
            }
          }
          uint4 o0, o1;
          o0.x = pack_bf16x2(f[0], f[1]);
          o0.y = pack_bf16x2(f[2], f[3]);
          o0.z = pack_bf16x2(f[4], f[5]);
          o0.w = pack_bf16x2(f[6], f[7]);
          o1.x = pack_bf16x2(f[8], f[9]);
          o1.y = pack_bf16x2(f[10], f[11]);
          o1.z = pack_bf16x2(f[12], f[13]);
          o1.w = pack_bf16x2(f[14], f[15]);
          uint4* o4 = reinterpret_cast<uint4*>(orow + c);
          o4[0] = o0;
          o4[1] = o1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kMc) cluster_sync_all();       // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

// Host launcher.  x: (N,H,W,Cin) bf16; wpack: [taps][Cout][Cin] bf16; out: (N,H,W,Cout) bf16.
int launch_conv_fprop(const void* x, const void* wpack, void* out, int N, int H, int W, int Cin, int Cout, int ksize,
                      const float* bias, const float* noise, const float* noise_w, const void* gate_src, int act,
                      float slope, cudaStream_t stream) {
  BG_REQUIRE(ksize == 3 || ksize == 1, "conv_fprop: ksize must be 1 or 3 (got %d)", ksize);
  BG_REQUIRE(Cin % 16 == 0 && Cout % 16 == 0, "conv_fprop: channels must be multiples of 16 (Cin %d Cout %d)", Cin,
             Cout);
  BG_REQUIRE(Cout <= kMaxCout, "conv_fprop: Cout %d exceeds %d", Cout, kMaxCout);
  BG_REQUIRE(N > 0 && H > 0 && W > 0, "conv_fprop: empty tensor");
  BG_REQUIRE((W & (W - 1)) == 0 && (H & (H - 1)) == 0, "conv_fprop: H and W must be powers of two (%d x %d)", H, W);

  FpropParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.ksize = ksize; p.pad = ksize / 2; p.taps = ksize * ksize;
  p.bw = W < 16 ? W : 16;
  p.bh = H < (128 / p.bw) ? H : (128 / p.bw);
  p.bn = 128 / (p.bw * p.bh);
  p.tiles_w = W / p.bw;
  p.tiles_h = H / p.bh;
  p.tiles_n = (N + p.bn - 1) / p.bn;
  // N tile: the largest multiple of 16 that divides Cout, is <= 256 and still yields ~one CTA per SM.  This kernel
  // only runs the small maps (<= 8x8: 4..64 pixel tiles), where a CTA's time is the L2->SMEM fill of its K loop
  // (~80 GB/s per SM): narrower N tiles spread that over more SMs (8x8, 512->512: 32 CTAs x 3.4 MB -> 128 x 1.7 MB).
  int bn_ch = 0;
  const int pixel_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  for (int c = 256; c >= 16; c -= 16) {
    if (Cout % c != 0) continue;
    bn_ch = c;
    if (pixel_tiles * (Cout / c) >= (num_sms() * 3) / 4) break;
  }
  BG_REQUIRE(bn_ch > 0, "conv_fprop: no valid N tile for Cout %d", Cout);
  p.block_n = bn_ch;
  p.n_blocks = Cout / bn_ch;
  p.kc = (Cin % 64 == 0) ? 64 : (Cin % 32 == 0 ? 32 : 16);
  p.k_chunks = Cin / p.kc;
  const int row_bytes = p.kc * 2;
  p.layout_type = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  p.sbo_bytes = 8u * row_bytes;
  p.a_bytes = 128u * row_bytes;
  p.b_bytes = (uint32_t)p.block_n * row_bytes;
  p.stage_bytes = (p.a_bytes + p.b_bytes + 1023u) & ~1023u;
  const uint32_t aux_bytes = 8 * (2 * kMaxStages + 4) + 16 + 2 * kMaxCout * 4;
  const uint32_t budget = 227u * 1024u - 1024u - aux_bytes;
  int stages = (int)(budget / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  BG_REQUIRE(stages >= 2, "conv_fprop: tile does not fit shared memory");
  p.stages = stages;
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_blocks;
  p.bias = bias; p.noise = noise; p.noise_w = noise_w;
  p.gate_src = reinterpret_cast<const __nv_bfloat16*>(gate_src);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.act = act; p.slope = slope;

  // cluster size for the A multicast: the largest of 8 / 4 / 2 that divides the number of n-blocks (128-byte rows and
  // whole 8-row swizzle atoms per slice only).  OFF by default (BG_FPROP_MC=1 enables it): measured on B200 it HALVES
  // the speed of the 4x4 / 8x8 layers (512 -> 512, batch 32: 37 -> 73 us, profiles/r2_conv_fprop_multicast.txt) — eight
  // CTAs advancing in lock step through 2 KB multicast slices lose more to the cluster-wide stage hand-shake than the
  // 2.4x lower L2 -> shared-memory traffic gains.  Kept as a measured negative result.
  p.csize = 1;
  {
    static int mc_on = -1;
    if (mc_on < 0) { const char* e = getenv("BG_FPROP_MC"); mc_on = (e && e[0] == '1') ? 1 : 0; }
    if (mc_on && row_bytes == 128 && p.n_blocks >= 2 && p.bw * p.bh * p.bn == 128)
      for (int c = 8; c >= 2; c >>= 1)
        if (p.n_blocks % c == 0) { p.csize = c; break; }
  }
  p.slice_rows = 128 / p.csize;
  p.slice_bytes = (uint32_t)p.slice_rows * (uint32_t)row_bytes;
  CUtensorMap tmx, tmw;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {(uint32_t)p.kc, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    if (p.csize > 1) {
      // one slice: slice_rows consecutive rows of the (w, h, n)-ordered tile = whole image rows, or whole images
      const int per_img = p.bw * p.bh;
      if (p.slice_rows >= per_img) { box[3] = (uint32_t)(p.slice_rows / per_img); }
      else { box[2] = (uint32_t)(p.slice_rows / p.bw); box[3] = 1u; }
      BG_REQUIRE(p.slice_rows % p.bw == 0 && (p.slice_rows >= per_img ? p.slice_rows % per_img == 0 : per_img % p.slice_rows == 0),
                 "conv_fprop: multicast slice does not tile the pixel box");
    }
    if (make_tmap_bf16(&tmx, x, 4, dims, str, box, row_bytes) != 0) return 1;
  }
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)p.taps};
    uint64_t str[2] = {(uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
    uint32_t box[3] = {(uint32_t)p.kc, (uint32_t)p.block_n, 1u};
    if (make_tmap_bf16(&tmw, wpack, 3, dims, str, box, row_bytes) != 0) return 1;
  }

  const size_t smem_bytes = (size_t)p.stages * p.stage_bytes + aux_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    BG_CHECK_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BG_CHECK_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  if (p.csize > 1) {
    // whole clusters only: the CTAs of a cluster walk the tile list in lock step (same pixel tile, consecutive n-blocks)
    grid = (grid / p.csize) * p.csize;
    BG_CHECK_CUDA(launch_pdl_cluster(conv_fprop_kernel<true>, grid, kThreads, smem_bytes, stream, p.csize, tmx, tmw, p));
    return 0;
  }
  BG_CHECK_CUDA(launch_pdl(conv_fprop_kernel<false>, grid, kThreads, smem_bytes, stream, tmx, tmw, p));
  return 0;
}

}  // namespace bg
