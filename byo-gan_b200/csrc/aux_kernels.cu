// HBM-bound helper kernels around the tensor-core convolutions.  All feature maps are NHWC bf16 with
// C a multiple of 8, processed as 16-byte vectors (8 channels) per thread, coalesced along C.
// Images at the module boundary are NCHW fp32 exactly as the reference passes them (gan.py:183,331).
#include "common.cuh"

namespace bg {

namespace {

constexpr int kBlock = 256;

struct F8 {
  float v[8];
};
__device__ __forceinline__ uint4 ld_raw8(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ F8 unpack8(const uint4 u) {
  F8 r;
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
  r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
  return r;
}
__device__ __forceinline__ F8 ld8(const __nv_bfloat16* p) { return unpack8(ld_raw8(p)); }
__device__ __forceinline__ void st8(__nv_bfloat16* p, const F8& r) {
  uint4 u;
  u.x = pack_bf16x2(r.v[0], r.v[1]);
  u.y = pack_bf16x2(r.v[2], r.v[3]);
  u.z = pack_bf16x2(r.v[4], r.v[5]);
  u.w = pack_bf16x2(r.v[6], r.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
// Index decomposition for the grid-stride elementwise kernels: every divisor of the model's maps (C/8, W, H) is a power
// of two, so quotient / remainder are a shift and a mask; 32-bit division is the fallback (element counts < 2^32).
struct Div32 {
  uint32_t d;
  int shift;      // >= 0: d == 1 << shift
};
inline Div32 make_div(uint32_t d) {
  Div32 k;
  k.d = d;
  k.shift = -1;
  if (d != 0 && (d & (d - 1)) == 0) {
    k.shift = 0;
    while ((1u << k.shift) < d) ++k.shift;
  }
  return k;
}
__device__ __forceinline__ uint32_t divmod(uint32_t x, const Div32& k, uint32_t& rem) {
  if (k.shift >= 0) {
    rem = x & (k.d - 1u);
    return x >> k.shift;
  }
  const uint32_t q = x / k.d;
  rem = x - q * k.d;
  return q;
}

// Blocks of `kernel` that are resident on the whole device at once: the reductions below give every block an equal share
// of the pixels, so a grid of exactly this many blocks is a single wave with no tail (a grid a few blocks larger costs a
// whole extra wave: 608 blocks on 592 slots ran in twice the time of 592).
template <typename K>
inline int resident_blocks(K kernel, int threads, size_t smem) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) occ = 1;
  return occ * num_sms();
}

inline int grid_for(size_t work, int cap_mult = 8) {
  size_t blocks = (work + kBlock - 1) / kBlock;
  size_t cap = (size_t)num_sms() * cap_mult;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}


// Block-level tail of a fused per-channel reduction in a grid-stride kernel whose thread <-> channel-group mapping
// is fixed (blockDim and gridDim * blockDim multiples of cv): thread t holds `nk` partial sums for the 8 channels of
// group t % cv.  red: [blockDim][nk][8] floats of shared memory.  out[k * C + c] += block total.
template <int NK>
__device__ __forceinline__ void block_channel_reduce(float (&acc)[NK][8], float* red, float* __restrict__ out, int C) {
  const int cv = C / 8;
#pragma unroll
  for (int k = 0; k < NK; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[((size_t)threadIdx.x * NK + k) * 8 + j] = acc[k][j];
  __syncthreads();
  const int rows = blockDim.x / cv;
  for (int idx = threadIdx.x; idx < NK * C; idx += blockDim.x) {
    const int k = idx / C, c = idx % C;
    const int grp = c >> 3, j = c & 7;
    float sum = 0.f;
    for (int r = 0; r < rows; ++r) sum += red[((size_t)(r * cv + grp) * NK + k) * 8 + j];
    atomicAdd(out + (size_t)k * C + c, sum);
  }
}

// ---------------------------------------------------------------------------------------------
// weight pack / unpack (equalized-lr coefficient folded in, gan.py:14,27,32)
// ---------------------------------------------------------------------------------------------
// w: fp32 (Cout, Cin, ks, ks).  wf: bf16 [tap][Cout][Cin_pad];  wd: bf16 [tap'][Cin_pad][Cout] with
// tap' = flipped tap, so that the dgrad pass is the same forward kernel run on the gradient.
__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                   __nv_bfloat16* __restrict__ wd, int Cout, int Cin, int Cin_pad, int ks, float coef) {
  pdl_prologue();
  const int taps = ks * ks;
  const size_t total = (size_t)taps * Cout * Cin_pad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin_pad);
    const int co = (int)((i / Cin_pad) % Cout);
    const int tap = (int)(i / ((size_t)Cin_pad * Cout));
    float v = 0.f;
    if (ci < Cin) v = w[((size_t)co * Cin + ci) * taps + tap] * coef;
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    if (wf) wf[i] = b;
    if (wd) wd[((size_t)(taps - 1 - tap) * Cin_pad + ci) * Cout + co] = b;
  }
}

// conv3x3 (pad 1) followed by AvgPool2d(2) (CriticBlock.conv_2, gan.py:258-260) is a 4x4 stride-2 convolution (pad 1)
// whose kernel is the quarter sum of the four shifted copies of the 3x3 one:
//   W4[a][b] = 1/4 * sum_{dy,dx in {0,1}} W3[a-dy][b-dx]        (terms with an index outside 0..2 drop out)
// w: fp32 (Cout, Cin, 3, 3) -> w16: bf16 [a*4+b][Cout][Cin], equalized coefficient folded in; summed in fp32, rounded once.
__global__ void pack_weight_pool4_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ w16, int Cout, int Cin,
                                         float coef) {
  pdl_prologue();
  const size_t total = (size_t)16 * Cout * Cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int co = (int)((i / Cin) % Cout);
    const int tap = (int)(i / ((size_t)Cin * Cout));
    const int a = tap >> 2, b = tap & 3;
    const float* w3 = w + ((size_t)co * Cin + ci) * 9;
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int ky = a - dy, kx = b - dx;
        if (ky >= 0 && ky <= 2 && kx >= 0 && kx <= 2) acc += w3[ky * 3 + kx];
      }
    w16[i] = __float2bfloat16_rn(0.25f * coef * acc);
  }
}

// Input-gradient (transposed conv) of that 4x4 stride-2 conv.  Output pixel (2i'+py, 2j'+px) receives the taps
// a = ((py + 1) & 1) + 2 ta, b = ((px + 1) & 1) + 2 tb from pooled pixel (i' + (py + 1 - a) / 2, j' + (px + 1 - b) / 2).
// wt: bf16 [(2 py + px) * 4 + 2 ta + tb][Cin][Cout] = W4[a][b][co][ci] transposed (GEMM N = ci, K = co).
__global__ void pack_weight_tconv4_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wt, int Cout, int Cin,
                                          float coef) {
  pdl_prologue();
  const size_t total = (size_t)16 * Cin * Cout;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const int ci = (int)((i / Cout) % Cin);
    const int idx = (int)(i / ((size_t)Cin * Cout));
    const int ph = idx >> 2, t = idx & 3;
    const int a = (((ph >> 1) + 1) & 1) + 2 * (t >> 1), b = (((ph & 1) + 1) & 1) + 2 * (t & 1);
    const float* w3 = w + ((size_t)co * Cin + ci) * 9;
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int ky = a - dy, kx = b - dx;
        if (ky >= 0 && ky <= 2 && kx >= 0 && kx <= 2) acc += w3[ky * 3 + kx];
      }
    wt[i] = __float2bfloat16_rn(0.25f * coef * acc);
  }
}

// All stale packs of a network in ONE launch (after an optimizer step every conv weight of the active stage is stale:
// 13..26 layers).  Group g owns blocks [blk0[g], blk0[g+1]).
constexpr int kMaxPackGroups = 32;
struct PackGroups {
  const float* w[kMaxPackGroups];
  __nv_bfloat16* wf[kMaxPackGroups];
  __nv_bfloat16* wd[kMaxPackGroups];
  int Cout[kMaxPackGroups], Cin[kMaxPackGroups], Cin_pad[kMaxPackGroups], ks[kMaxPackGroups];
  float coef[kMaxPackGroups];
  int blk0[kMaxPackGroups + 1];
  int groups;
};
__global__ void pack_weight_grouped_kernel(const PackGroups G) {
  pdl_prologue();
  int g = 0;
  while (g + 1 < G.groups && (int)blockIdx.x >= G.blk0[g + 1]) ++g;
  const int Cout = G.Cout[g], Cin = G.Cin[g], Cin_pad = G.Cin_pad[g], taps = G.ks[g] * G.ks[g];
  const float coef = G.coef[g];
  const float* __restrict__ w = G.w[g];
  __nv_bfloat16* __restrict__ wf = G.wf[g];
  __nv_bfloat16* __restrict__ wd = G.wd[g];
  const size_t total = (size_t)taps * Cout * Cin_pad;
  const size_t nthreads = (size_t)(G.blk0[g + 1] - G.blk0[g]) * blockDim.x;
  for (size_t i = ((size_t)blockIdx.x - G.blk0[g]) * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int ci = (int)(i % Cin_pad);
    const int co = (int)((i / Cin_pad) % Cout);
    const int tap = (int)(i / ((size_t)Cin_pad * Cout));
    float v = 0.f;
    if (ci < Cin) v = w[((size_t)co * Cin + ci) * taps + tap] * coef;
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    wf[i] = b;
    wd[((size_t)(taps - 1 - tap) * Cin_pad + ci) * Cout + co] = b;
  }
}

// Tiled form for 3x3 layers (every group ks == 3): a block owns a (64 co x 64 ci) tile of one layer; it reads the 9-tap runs
// w[co][ci0..ci0+63][0..8] (576 contiguous floats per co) coalesced, keeps the bf16 tile in shared memory and writes 128-byte
// runs on both sides: wf[tap][co][ci0..] and wd[8 - tap][ci][co0..].  The element-wise kernel above wrote wd with one
// 2-byte store per 128-byte line (154 us for a network's packs that move in ~40 us); since the pack cache was fixed to
// notice fused optimizer steps this runs twice per iteration.
constexpr int kPackTile = 64;
constexpr int kPackStride = kPackTile + 8;                        // 144-byte rows: 16-byte aligned, conflict-free 16-byte reads
constexpr int kPackTapStride = kPackTile * kPackStride + 8;       // 16-byte aligned, tap planes 4 banks apart
__global__ void __launch_bounds__(256)
pack_weight_grouped_tiled_kernel(const PackGroups G) {
  pdl_prologue();
  extern __shared__ __align__(16) __nv_bfloat16 ptile[];    // [9][64 co][72 ci]
  int g = 0;
  while (g + 1 < G.groups && (int)blockIdx.x >= G.blk0[g + 1]) ++g;
  const int Cout = G.Cout[g], Cin = G.Cin[g], Cin_pad = G.Cin_pad[g];      // Cout % 16 == 0, Cin_pad % 8 == 0 (launcher)
  const float coef = G.coef[g];
  const float* __restrict__ w = G.w[g];
  __nv_bfloat16* __restrict__ wf = G.wf[g];
  __nv_bfloat16* __restrict__ wd = G.wd[g];
  const int tiles_ci = (Cin_pad + kPackTile - 1) / kPackTile;
  const int tile = (int)blockIdx.x - G.blk0[g];
  const int co0 = (tile / tiles_ci) * kPackTile, ci0 = (tile % tiles_ci) * kPackTile;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  // read: warp w takes the columns co0 + w, w + 8, ...; a column is one run of 64 * 9 floats (zero beyond Cin), 18 coalesced
  // loads per lane, all in flight; every index division is by a constant
  for (int col = wp; col < kPackTile; col += 8) {
    if (co0 + col >= Cout) break;
    const float* run = w + ((size_t)(co0 + col) * Cin + ci0) * 9;
    float v[18];
#pragma unroll
    for (int it = 0; it < 18; ++it) {
      const int j = it * 32 + lane;
      v[it] = (ci0 + j / 9 < Cin) ? run[j] : 0.f;
    }
#pragma unroll
    for (int it = 0; it < 18; ++it) {
      const int j = it * 32 + lane;
      const int cil = j / 9, tap = j - cil * 9;
      ptile[tap * kPackTapStride + col * kPackStride + cil] = __float2bfloat16_rn(v[it] * coef);
    }
  }
  __syncthreads();
  // wf[tap][co][ci]: 16-byte stores, 8 lanes per 128-byte row
  for (int idx = threadIdx.x; idx < 9 * kPackTile * 8; idx += 256) {
    const int vq = idx & 7, col = (idx >> 3) & (kPackTile - 1), tap = idx >> 9;
    if (co0 + col < Cout && ci0 + vq * 8 < Cin_pad)
      *reinterpret_cast<uint4*>(wf + ((size_t)tap * Cout + co0 + col) * Cin_pad + ci0 + vq * 8) =
          *reinterpret_cast<const uint4*>(ptile + tap * kPackTapStride + col * kPackStride + vq * 8);
  }
  // wd[8 - tap][ci][co]: a lane gathers 16 co of one ci (lanes along ci: conflict-free 2-byte reads) and stores 32 bytes
  for (int idx = threadIdx.x; idx < 9 * 4 * kPackTile; idx += 256) {
    const int cil = idx & (kPackTile - 1), cg = (idx >> 6) & 3, tap = idx >> 8;
    if (ci0 + cil >= Cin_pad || co0 + cg * 16 >= Cout) continue;
    const __nv_bfloat16* src = ptile + tap * kPackTapStride + (cg * 16) * kPackStride + cil;
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t lo = *reinterpret_cast<const unsigned short*>(src + (2 * i) * kPackStride);
      const uint32_t hi = *reinterpret_cast<const unsigned short*>(src + (2 * i + 1) * kPackStride);
      pk[i] = lo | (hi << 16);
    }
    uint4* dst = reinterpret_cast<uint4*>(wd + ((size_t)(8 - tap) * Cin_pad + ci0 + cil) * Cout + co0 + cg * 16);
    dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

// dwp: fp32 [tap][Cout][Cin_pad] -> dw: fp32 (Cout, Cin, ks, ks), scaled by coef; optionally accumulates.
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Cout, int Cin,
                                    int Cin_pad, int ks, float coef, int accumulate) {
  pdl_prologue();
  const int taps = ks * ks;
  const size_t total = (size_t)Cout * Cin * taps;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const int ci = (int)((i / taps) % Cin);
    const int co = (int)(i / ((size_t)taps * Cin));
    const float v = dwp[((size_t)tap * Cout + co) * Cin_pad + ci] * coef;
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

// Tiled form of the two unpack kernels for 3x3 layers: a block owns (one co, 128 ci); it reads the 9 (or 16) tap planes
// coalesced along ci into shared memory and writes the (ci, ky, kx)-ordered run of 128 * 9 floats coalesced.  The
// element-wise kernels above read 9 tap planes with 16 useful bytes per 128-byte line (17 us for a 512 x 512 layer whose
// 19 MB move in 3 us).  MODE 0: plain transpose; MODE 1: the pool4 fold dW3[ky][kx] = 1/4 sum dW4[ky+dy][kx+dx].
template <int MODE>
__global__ void __launch_bounds__(128)
unpack_wgrad_tiled_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Cout, int Cin, int Cin_pad, float coef,
                          int accumulate) {
  pdl_prologue();
  constexpr int kTapsIn = MODE == 0 ? 9 : 16;
  __shared__ float tile[kTapsIn][129];
  const int co = blockIdx.y, ci0 = blockIdx.x * 128, t = threadIdx.x;
#pragma unroll
  for (int tap = 0; tap < kTapsIn; ++tap)
    tile[tap][t] = (ci0 + t < Cin) ? dwp[((size_t)tap * Cout + co) * Cin_pad + ci0 + t] : 0.f;
  __syncthreads();
  const int nci = min(128, Cin - ci0);
  float* out = dw + ((size_t)co * Cin + ci0) * 9;
  for (int j = t; j < nci * 9; j += 128) {
    const int cil = j / 9, tap = j - cil * 9;
    float v;
    if (MODE == 0) {
      v = tile[tap][cil];
    } else {
      const int ky = tap / 3, kx = tap - ky * 3;
      v = 0.25f * (tile[ky * 4 + kx][cil] + tile[ky * 4 + kx + 1][cil] + tile[(ky + 1) * 4 + kx][cil] +
                   tile[(ky + 1) * 4 + kx + 1][cil]);
    }
    v *= coef;
    out[j] = accumulate ? out[j] + v : v;
  }
}

// dw4: fp32 [16][Cout][Cin] gradient of the 4x4 stride-2 kernel W4[a][b] = 1/4 sum_{dy,dx} W3[a-dy][b-dx]  ->
// dw: fp32 (Cout, Cin, 3, 3):  dW3[ky][kx] = coef / 4 * sum_{dy,dx in {0,1}} dW4[ky+dy][kx+dx]   (adjoint of the pack)
__global__ void unpack_wgrad_pool4_kernel(const float* __restrict__ dw4, float* __restrict__ dw, int Cout, int Cin,
                                          float coef, int accumulate) {
  pdl_prologue();
  const size_t total = (size_t)Cout * Cin * 9;
  const size_t plane = (size_t)Cout * Cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 9);
    const size_t cc = i / 9;                      // co * Cin + ci
    const int ky = tap / 3, kx = tap % 3;
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) acc += dw4[(size_t)((ky + dy) * 4 + kx + dx) * plane + cc];
    const float v = 0.25f * coef * acc;
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

// ---------------------------------------------------------------------------------------------
// LeakyReLU gate:  out = g * (y > 0 ? 1 : slope)      (backward of nn.LeakyReLU(0.2), gan.py:86,241...)
// ---------------------------------------------------------------------------------------------
__global__ void act_gate_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ y,
                                __nv_bfloat16* __restrict__ out, size_t nvec, float slope) {
  pdl_prologue();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    F8 a = ld8(g + i * 8);
    const F8 b = ld8(y + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) a.v[j] *= b.v[j] > 0.f ? 1.f : slope;
    st8(out + i * 8, a);
  }
}

// out = ca * a + cb * b  (torch.lerp on feature maps, gan.py:347, and its gradient scalings)
__global__ void axpby_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                             __nv_bfloat16* __restrict__ out, size_t nvec, float ca, float cb) {
  pdl_prologue();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    F8 x = ld8(a + i * 8);
    if (b != nullptr) {
      const F8 y = ld8(b + i * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) x.v[j] = ca * x.v[j] + cb * y.v[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x.v[j] = ca * x.v[j];
    }
    st8(out + i * 8, x);
  }
}

// ---------------------------------------------------------------------------------------------
// AvgPool2d(2) + LeakyReLU (critic block tail, gan.py:258-262) and its adjoint
// ---------------------------------------------------------------------------------------------
// mode 0: y = lrelu(avg4(u));  mode 1: y = avg4(u) * gate(gate_src)   (tangent pass of the backward)
__global__ void pool_act_fwd_kernel(const __nv_bfloat16* __restrict__ u, const __nv_bfloat16* __restrict__ gate_src,
                                    __nv_bfloat16* __restrict__ y, int N, int Ho, int Wo, int C, float slope,
                                    int mode, Div32 dcv, Div32 dw, Div32 dh) {
  pdl_prologue();
  const int cv = C / 8;
  const size_t total = (size_t)N * Ho * Wo * cv;
  const int W = Wo * 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t cg, uw, uh;
    const uint32_t po = divmod((uint32_t)i, dcv, cg);
    const uint32_t n = divmod(divmod(po, dw, uw), dh, uh);
    const int c = (int)cg * 8, wo = (int)uw, ho = (int)uh;
    const size_t base = (((size_t)n * Ho * 2 + ho * 2) * W + wo * 2) * C + c;
    const F8 a = ld8(u + base), b = ld8(u + base + C), d = ld8(u + base + (size_t)W * C),
             e = ld8(u + base + (size_t)W * C + C);
    F8 r;
    if (mode == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float s = 0.25f * (a.v[j] + b.v[j] + d.v[j] + e.v[j]);
        r.v[j] = s > 0.f ? s : s * slope;
      }
    } else {
      const F8 gt = ld8(gate_src + (size_t)po * C + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float s = 0.25f * (a.v[j] + b.v[j] + d.v[j] + e.v[j]);
        r.v[j] = s * (gt.v[j] > 0.f ? 1.f : slope);
      }
    }
    st8(y + (size_t)po * C + c, r);
  }
}

// gu[2h+dy, 2w+dx] = 0.25 * gy[h, w] * gate(y[h, w]);  csum (optional, zeroed by the launcher): csum[c] += sum over
// the full-resolution map of gu[.,c] = sum_{n,h,w} gy * gate — the bias gradient of the conv that feeds the pool.
__global__ void pool_act_bwd_kernel(const __nv_bfloat16* __restrict__ gy, const __nv_bfloat16* __restrict__ y,
                                    __nv_bfloat16* __restrict__ gu, int N, int Ho, int Wo, int C, float slope,
                                    float* __restrict__ csum, Div32 dcv, Div32 dw, Div32 dh) {
  pdl_prologue();
  extern __shared__ float red[];
  const int cv = C / 8;
  const size_t total = (size_t)N * Ho * Wo * cv;
  const int W = Wo * 2;
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t cg, uw, uh;
    const uint32_t po = divmod((uint32_t)i, dcv, cg);
    const uint32_t n = divmod(divmod(po, dw, uw), dh, uh);
    const int c = (int)cg * 8, wo = (int)uw, ho = (int)uh;
    F8 g = ld8(gy + (size_t)po * C + c);
    const F8 yy = ld8(y + (size_t)po * C + c);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      g.v[j] *= 0.25f * (yy.v[j] > 0.f ? 1.f : slope);
      acc[0][j] += g.v[j];
    }
    const size_t base = (((size_t)n * Ho * 2 + ho * 2) * W + wo * 2) * C + c;
    st8(gu + base, g);
    st8(gu + base + C, g);
    st8(gu + base + (size_t)W * C, g);
    st8(gu + base + (size_t)W * C + C, g);
  }
  if (csum != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] *= 4.f;
    block_channel_reduce<1>(acc, red, csum, C);
  }
}

// ---------------------------------------------------------------------------------------------
// bilinear x2 upsample, align_corners=False (nn.Upsample, gan.py:112,123): taps .75/.25, edge clamp
// ---------------------------------------------------------------------------------------------
__global__ void upsample2x_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N,
                                      int H, int W, int C, Div32 dcv, Div32 dw, Div32 dh) {
  pdl_prologue();
  // one thread per (source pixel, 8 channels): the clamped 3 x 3 source window gives the 2 x 2 output quad, 9 loads
  // for 4 stores (a thread per output pixel would load 16)
  const int cv = C / 8;
  const int Wo = 2 * W;
  const size_t total = (size_t)N * H * W * cv;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t cg, uw, uh;
    const uint32_t p = divmod((uint32_t)i, dcv, cg);
    const uint32_t n = divmod(divmod(p, dw, uw), dh, uh);
    const int c = (int)cg * 8, w = (int)uw, h = (int)uh;
    const int hs[3] = {max(h - 1, 0), h, min(h + 1, H - 1)};
    const int ws[3] = {max(w - 1, 0), w, min(w + 1, W - 1)};
    const __nv_bfloat16* xb = x + (size_t)n * H * W * C + c;
    uint4 raw[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < 3; ++q) raw[r][q] = ld_raw8(xb + ((size_t)hs[r] * W + ws[q]) * C);
    const F8 a = unpack8(raw[1][1]);
    __nv_bfloat16* yb = y + (((size_t)n * 2 * H + 2 * h) * Wo + 2 * w) * C + c;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      // output row 2h + dy: even rows lean on h-1, odd rows on h+1 (taps .75 / .25, edge clamp)
      const F8 d = unpack8(raw[2 * dy][1]);
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const F8 b = unpack8(raw[1][2 * dx]), e = unpack8(raw[2 * dy][2 * dx]);
        F8 r;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          r.v[j] = 0.5625f * a.v[j] + 0.1875f * (b.v[j] + d.v[j]) + 0.0625f * e.v[j];
        st8(yb + ((size_t)dy * Wo + dx) * C, r);
      }
    }
  }
}

// weight with which hi-res index r (0..2L-1) reads lo-res index l, after the edge clamp
__device__ __forceinline__ float up_weight(int r, int l, int L) {
  const int h = r >> 1;
  const int h2 = (r & 1) ? min(h + 1, L - 1) : max(h - 1, 0);
  float wgt = 0.f;
  if (h == l) wgt += 0.75f;
  if (h2 == l) wgt += 0.25f;
  return wgt;
}

// adjoint of the above: gx[h,w] = sum_{r,s} up_weight(r,h) up_weight(s,w) gy[r,s].  Separable: a thread owns 8 channels of
// kUpStrip vertically adjacent low-res pixels, walks the 2 * kUpStrip + 2 hi-res rows that reach them once, folds each row's
// four columns with the horizontal weights and adds the result to the (at most two) outputs the row feeds:
// (2S + 2) * 4 loads for S outputs instead of 16 per output.
constexpr int kUpStrip = 4;
__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ gy, __nv_bfloat16* __restrict__ gx, int N,
                                      int H, int W, int C, Div32 dcv, Div32 dw, Div32 dh) {
  pdl_prologue();
  const int cv = C / 8;
  const int Ho = 2 * H, Wo = 2 * W;
  const size_t total = (size_t)N * (H / kUpStrip) * W * cv;           // dh divides by H / kUpStrip
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t cg, uw, uh;
    const uint32_t p = divmod((uint32_t)i, dcv, cg);
    const uint32_t n = divmod(divmod(p, dw, uw), dh, uh);
    const int c = (int)cg * 8, w = (int)uw, h0 = (int)uh * kUpStrip;
    const __nv_bfloat16* gb = gy + (size_t)n * Ho * Wo * C + c;
    float wsv[4];
    int cx[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int sx = 2 * w - 1 + q;
      wsv[q] = (sx >= 0 && sx < Wo) ? up_weight(sx, w, W) : 0.f;     // out-of-range columns: clamped address, weight 0
      cx[q] = min(max(sx, 0), Wo - 1);
    }
    F8 acc[kUpStrip];
#pragma unroll
    for (int k = 0; k < kUpStrip; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[k].v[j] = 0.f;
#pragma unroll
    for (int qq = 0; qq < 2 * kUpStrip + 2; ++qq) {
      const int r = 2 * h0 - 1 + qq;
      const bool live = r >= 0 && r < Ho;
      const int rc = min(max(r, 0), Ho - 1);
      uint4 raw[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) raw[t] = ld_raw8(gb + ((size_t)rc * Wo + cx[t]) * C);
      F8 row;
#pragma unroll
      for (int j = 0; j < 8; ++j) row.v[j] = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const F8 g = unpack8(raw[t]);
#pragma unroll
        for (int j = 0; j < 8; ++j) row.v[j] += wsv[t] * g.v[j];
      }
      // hi-res row r = 2 * h0 - 1 + qq reaches the outputs k with 2k <= qq <= 2k + 3
#pragma unroll
      for (int k = 0; k < kUpStrip; ++k) {
        if (qq >= 2 * k && qq <= 2 * k + 3) {
          const float wr = live ? up_weight(r, h0 + k, H) : 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[k].v[j] += wr * row.v[j];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kUpStrip; ++k) st8(gx + (((size_t)n * H + h0 + k) * W + w) * C + c, acc[k]);
  }
}

// ---------------------------------------------------------------------------------------------
// per-channel weighted sums over pixels:
//   out[0][c] = sum_p g[p,c];   out[1+j][c] = sum_p g[p,c] * plane_j[p]   (j < nplanes <= 3)
// plane_j[p] = planes[(p / HW) * img_stride + j * plane_stride + (p % HW)]  (fp32)
// Serves: bias grads, noise-weight grads (gan.py:52), fromRGB / toRGB weight grads.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3)
channel_wsum_kernel(const __nv_bfloat16* __restrict__ g, const float* __restrict__ planes,
                                    float* __restrict__ out, size_t P, int C, int HW, size_t img_stride,
                                    size_t plane_stride, int nplanes, int pix_per_block, int hw_shift) {
  pdl_prologue();
  extern __shared__ float red[];  // [rows][4][C]
  const int cv = C / 8;
  const int rows = blockDim.x / cv;  // pixel lanes per block
  const int tc = threadIdx.x % cv;
  const int tr = threadIdx.x / cv;
  float acc[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  const size_t p0 = (size_t)blockIdx.x * pix_per_block;
  size_t p1 = p0 + pix_per_block;
  if (p1 > P) p1 = P;
  if (tr < rows) {
    // HW is a power of two for every map of the model: the plane index costs a shift and a mask instead of two 64-bit
    // divisions per pixel (hw_shift < 0: general fallback)
    constexpr int kU = 4;              // pixels per trip, loads issued ahead of the arithmetic
    size_t p = p0 + tr;
    for (; p + (size_t)(kU - 1) * rows < p1; p += (size_t)kU * rows) {
      uint4 raw[kU];
      float s[kU][3];
#pragma unroll
      for (int u = 0; u < kU; ++u) raw[u] = ld_raw8(g + (p + (size_t)u * rows) * C + tc * 8);
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        s[u][0] = s[u][1] = s[u][2] = 0.f;
        if (nplanes > 0) {
          const size_t pu = p + (size_t)u * rows;
          const size_t b = hw_shift >= 0 ? (pu >> hw_shift) * img_stride + (pu & (size_t)(HW - 1))
                                         : (pu / HW) * img_stride + (pu % HW);
#pragma unroll
          for (int j = 0; j < 3; ++j)
            if (j < nplanes) s[u][j] = planes[b + j * plane_stride];
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const F8 v = unpack8(raw[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] += v.v[j];
          acc[1][j] += v.v[j] * s[u][0];
          acc[2][j] += v.v[j] * s[u][1];
          acc[3][j] += v.v[j] * s[u][2];
        }
      }
    }
    for (; p < p1; p += rows) {
      const F8 v = ld8(g + p * C + tc * 8);
      float s[3] = {0.f, 0.f, 0.f};
      if (nplanes > 0) {
        const size_t b = hw_shift >= 0 ? (p >> hw_shift) * img_stride + (p & (size_t)(HW - 1))
                                       : (p / HW) * img_stride + (p % HW);
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (j < nplanes) s[j] = planes[b + j * plane_stride];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += v.v[j];
        acc[1][j] += v.v[j] * s[0];
        acc[2][j] += v.v[j] * s[1];
        acc[3][j] += v.v[j] * s[2];
      }
    }
  }
  const int nk = 1 + nplanes;
  if (tr < rows) {
#pragma unroll
    for (int k = 0; k < 4; ++k)          // fully unrolled: a runtime index would push acc[][] into local memory
      if (k < nk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[((size_t)tr * 4 + k) * C + tc * 8 + j] = acc[k][j];
      }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < nk * C; idx += blockDim.x) {
    const int k = idx / C, c = idx % C;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += red[((size_t)r * 4 + k) * C + c];
    atomicAdd(out + (size_t)k * C + c, s);
  }
}

// ---------------------------------------------------------------------------------------------
// 3-plane NCHW fp32 image  <->  NHWC bf16 features through a (C x 3) matrix
// ---------------------------------------------------------------------------------------------
// out[p, c] = act( coef * sum_j img[n, j, hw] * Wm[c * ws_c + j * ws_j] + bias[c] )
//   fromRGB forward (gan.py:351-355): Wm = weight (C,3,1,1) -> ws_c = 3, ws_j = 1, bias, act
//   toRGB input-grad:                 Wm = weight (3,C,1,1) -> ws_c = 1, ws_j = C, no bias, no act
__global__ void planes3_to_nhwc_kernel(const float* __restrict__ img, const float* __restrict__ Wm,
                                       const float* __restrict__ bias, const __nv_bfloat16* __restrict__ gate_src,
                                       __nv_bfloat16* __restrict__ out, size_t P, int HW, int C, int ws_c, int ws_j,
                                       float coef, int act, float slope, Div32 dcv, Div32 dhw) {
  pdl_prologue();
  extern __shared__ float sw[];  // [C][3] + [C]
  for (int i = threadIdx.x; i < C * 3; i += blockDim.x) {
    const int c = i / 3, j = i % 3;
    sw[i] = Wm[(size_t)c * ws_c + (size_t)j * ws_j] * coef;
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) sw[C * 3 + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int cv = C / 8;
  const size_t total = P * cv;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t cg, hw;
    const uint32_t p = divmod((uint32_t)i, dcv, cg);
    const uint32_t n = divmod(p, dhw, hw);
    const int c = (int)cg * 8;
    const size_t b = (size_t)n * (size_t)(3 * HW) + hw;
    const float i0 = img[b], i1 = img[b + HW], i2 = img[b + 2 * (size_t)HW];
    F8 r;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* w = sw + (c + j) * 3;
      float v = i0 * w[0] + i1 * w[1] + i2 * w[2] + sw[C * 3 + c + j];
      if (act) v = v > 0.f ? v : v * slope;
      r.v[j] = v;
    }
    if (gate_src != nullptr) {
      const F8 gt = ld8(gate_src + (size_t)p * C + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) r.v[j] *= gt.v[j] > 0.f ? 1.f : slope;
    }
    st8(out + (size_t)p * C + c, r);
  }
}

// Same map, four consecutive pixels of one sample and 8 channels per thread: three 16-byte plane loads, the 24
// weights + 8 biases of the channel group held in registers across the four pixels (HW % 4 == 0, img 16-byte aligned).
__global__ void planes3_to_nhwc_quad_kernel(const float* __restrict__ img, const float* __restrict__ Wm,
                                            const float* __restrict__ bias, const __nv_bfloat16* __restrict__ gate_src,
                                            __nv_bfloat16* __restrict__ out, size_t P, int HW, int C, int ws_c, int ws_j,
                                            float coef, int act, float slope, Div32 dcv, Div32 dq) {
  pdl_prologue();
  extern __shared__ float sw[];  // [C][3] + [C]
  for (int i = threadIdx.x; i < C * 3; i += blockDim.x) {
    const int c = i / 3, j = i % 3;
    sw[i] = Wm[(size_t)c * ws_c + (size_t)j * ws_j] * coef;
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) sw[C * 3 + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int cv = C / 8;
  const size_t total = (P / 4) * cv;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t cg, qi;
    const uint32_t q = divmod((uint32_t)i, dcv, cg);          // pixel quad
    const uint32_t n = divmod(q, dq, qi);                      // dq = HW / 4 quads per sample
    const int c = (int)cg * 8;
    const size_t b = (size_t)n * (size_t)(3 * HW) + (size_t)qi * 4;
    const float4 p0 = *reinterpret_cast<const float4*>(img + b);
    const float4 p1 = *reinterpret_cast<const float4*>(img + b + HW);
    const float4 p2 = *reinterpret_cast<const float4*>(img + b + 2 * (size_t)HW);
    const size_t pix = (size_t)q * 4;
    uint4 graw[4];
    if (gate_src != nullptr) {
#pragma unroll
      for (int k = 0; k < 4; ++k) graw[k] = ld_raw8(gate_src + (pix + k) * C + c);
    }
    float w0[8], w1[8], w2[8], bb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      w0[j] = sw[(c + j) * 3];
      w1[j] = sw[(c + j) * 3 + 1];
      w2[j] = sw[(c + j) * 3 + 2];
      bb[j] = sw[C * 3 + c + j];
    }
    const float i0[4] = {p0.x, p0.y, p0.z, p0.w}, i1[4] = {p1.x, p1.y, p1.z, p1.w}, i2[4] = {p2.x, p2.y, p2.z, p2.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      F8 r;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = i0[k] * w0[j] + i1[k] * w1[j] + i2[k] * w2[j] + bb[j];
        if (act) v = v > 0.f ? v : v * slope;
        r.v[j] = v;
      }
      if (gate_src != nullptr) {
        const F8 gt = unpack8(graw[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) r.v[j] *= gt.v[j] > 0.f ? 1.f : slope;
      }
      st8(out + (pix + k) * C + c, r);
    }
  }
}

// out[n, j, hw] = coef * sum_c x[p, c] * Wm[c * ws_c + j * ws_j] + bias[j]
//   toRGB forward (gan.py:172-179,218,222): Wm = weight (3,C,1,1) -> ws_c = 1, ws_j = C, bias
//   fromRGB input-grad:                     Wm = weight (C,3,1,1) -> ws_c = 3, ws_j = 1, no bias
// One warp per pixel group: lanes split the channel vectors, shuffle-reduce the 3 dot products.  kRegW (C <= 256: a lane
// sees one channel vector only): its 24 weights live in registers; kNhwcUnroll pixel groups per trip keep that many
// 16-byte loads in flight per lane.
constexpr int kNhwcUnroll = 4;
template <bool kRegW>
__global__ void nhwc_to_planes3_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ Wm,
                                       const float* __restrict__ bias, float* __restrict__ out, size_t P, int HW,
                                       int C, int ws_c, int ws_j, float coef) {
  pdl_prologue();
  extern __shared__ float sw[];  // [3][C]
  for (int i = threadIdx.x; i < C * 3; i += blockDim.x) {
    const int j = i / C, c = i % C;
    sw[i] = Wm[(size_t)c * ws_c + (size_t)j * ws_j] * coef;
  }
  __syncthreads();
  const int cv = C / 8;
  const int lanes_per_pix = cv < 32 ? cv : 32;   // 2,4,8,16,32
  const int pix_per_warp = 32 / lanes_per_pix;
  const int lane = threadIdx.x & 31;
  const int sub = lane % lanes_per_pix;
  const int pw = lane / lanes_per_pix;
  const size_t warp_global = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t groups = (P + pix_per_warp - 1) / pix_per_warp;
  const float b0 = bias ? bias[0] : 0.f, b1 = bias ? bias[1] : 0.f, b2 = bias ? bias[2] : 0.f;
  float w0[8], w1[8], w2[8];
  if (kRegW) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      w0[j] = sw[sub * 8 + j];
      w1[j] = sw[C + sub * 8 + j];
      w2[j] = sw[2 * C + sub * 8 + j];
    }
  }
  for (size_t g0 = warp_global; g0 < groups; g0 += nwarps * kNhwcUnroll) {
    float s0[kNhwcUnroll], s1[kNhwcUnroll], s2[kNhwcUnroll];
    if (kRegW) {
      uint4 raw[kNhwcUnroll];
#pragma unroll
      for (int u = 0; u < kNhwcUnroll; ++u) {
        const size_t p = (g0 + (size_t)u * nwarps) * pix_per_warp + pw;
        raw[u] = p < P ? ld_raw8(x + p * C + sub * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < kNhwcUnroll; ++u) {
        const F8 a = unpack8(raw[u]);
        s0[u] = s1[u] = s2[u] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s0[u] += a.v[j] * w0[j];
          s1[u] += a.v[j] * w1[j];
          s2[u] += a.v[j] * w2[j];
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < kNhwcUnroll; ++u) {
        const size_t p = (g0 + (size_t)u * nwarps) * pix_per_warp + pw;
        s0[u] = s1[u] = s2[u] = 0.f;
        if (p < P) {
          for (int v = sub; v < cv; v += lanes_per_pix) {
            const F8 a = ld8(x + p * C + v * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = v * 8 + j;
              s0[u] += a.v[j] * sw[c];
              s1[u] += a.v[j] * sw[C + c];
              s2[u] += a.v[j] * sw[2 * C + c];
            }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kNhwcUnroll; ++u) {
      for (int o = lanes_per_pix >> 1; o > 0; o >>= 1) {
        s0[u] += __shfl_xor_sync(0xffffffffu, s0[u], o);
        s1[u] += __shfl_xor_sync(0xffffffffu, s1[u], o);
        s2[u] += __shfl_xor_sync(0xffffffffu, s2[u], o);
      }
      const size_t p = (g0 + (size_t)u * nwarps) * pix_per_warp + pw;
      if (p < P && sub == 0) {
        const size_t b = (p / HW) * (size_t)(3 * HW) + (p % HW);
        out[b] = s0[u] + b0;
        out[b + HW] = s1[u] + b1;
        out[b + 2 * (size_t)HW] = s2[u] + b2;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// instance norm + AdaIN (gan.py:55-71): statistics, apply, and backward
// ---------------------------------------------------------------------------------------------
// sums[n][c][0] += sum_hw a (* b if b given);  sums[n][c][1] += sum_hw a*a   (mode 0: statistics)
// mode 1 (backward): sums[n][c][0] += sum g, sums[n][c][1] += sum g * ahat, ahat = (a - mean) * rstd
__global__ void __launch_bounds__(256, 3)
in_reduce_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ g,
                                 const float* __restrict__ stats, float* __restrict__ sums, int HW, int C,
                                 int pix_per_block, int blocks_per_img, float eps, int mode) {
  pdl_prologue();
  extern __shared__ float red[];  // [rows][2][C]
  const int n = blockIdx.x / blocks_per_img;
  const int chunk = blockIdx.x % blocks_per_img;
  const int cv = C / 8;
  const int rows = blockDim.x / cv;
  const int tc = threadIdx.x % cv;
  const int tr = threadIdx.x / cv;
  float s1[8], s2[8], mean[8], rstd[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (mode == 1 && tr < rows) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* st = stats + ((size_t)n * C + tc * 8 + j) * 2;
      const float m = st[0] / HW;
      const float var = fmaxf(st[1] / HW - m * m, 0.f);
      mean[j] = m;
      rstd[j] = rsqrtf(var + eps);
    }
  }
  const int p0 = chunk * pix_per_block;
  const int p1 = min(p0 + pix_per_block, HW);
  if (tr < rows) {
    // four pixels per trip, all loads issued ahead of the arithmetic (the kernel is a pure HBM stream)
    constexpr int kU = 4;
    const __nv_bfloat16* ab = a + (size_t)n * HW * C + tc * 8;
    const __nv_bfloat16* gb = mode == 0 ? ab : g + (size_t)n * HW * C + tc * 8;
    int p = p0 + tr;
    for (; p + (kU - 1) * rows < p1; p += kU * rows) {
      uint4 ra[kU], rg[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) ra[u] = ld_raw8(ab + (size_t)(p + u * rows) * C);
      if (mode != 0) {
#pragma unroll
        for (int u = 0; u < kU; ++u) rg[u] = ld_raw8(gb + (size_t)(p + u * rows) * C);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const F8 av = unpack8(ra[u]);
        if (mode == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            s1[j] += av.v[j];
            s2[j] += av.v[j] * av.v[j];
          }
        } else {
          const F8 gv = unpack8(rg[u]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            s1[j] += gv.v[j];
            s2[j] += gv.v[j] * (av.v[j] - mean[j]) * rstd[j];
          }
        }
      }
    }
    for (; p < p1; p += rows) {
      const F8 av = ld8(ab + (size_t)p * C);
      if (mode == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += av.v[j];
          s2[j] += av.v[j] * av.v[j];
        }
      } else {
        const F8 gv = ld8(gb + (size_t)p * C);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += gv.v[j];
          s2[j] += gv.v[j] * (av.v[j] - mean[j]) * rstd[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[((size_t)tr * 2 + 0) * C + tc * 8 + j] = s1[j];
      red[((size_t)tr * 2 + 1) * C + tc * 8 + j] = s2[j];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * C; idx += blockDim.x) {
    const int k = idx / C, c = idx % C;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += red[((size_t)r * 2 + k) * C + c];
    atomicAdd(sums + ((size_t)n * C + c) * 2 + k, s);
  }
}

// x = gamma * (a - mean) * rstd + beta;   style = [gamma (C) | beta (C)] per sample (gan.py:66-69)
// grid = (blocks per sample, N): a thread keeps ONE (sample, 8-channel group) for its whole loop, so the per-(n,c)
// scale / shift are computed once into registers and the inner loop is load - 8 FMA - store (HBM-bound).
__global__ void adain_apply_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ stats,
                                   const float* __restrict__ style, __nv_bfloat16* __restrict__ x, int N, int HW,
                                   int C, float eps) {
  pdl_prologue();
  const int cv = C / 8;
  const int n = blockIdx.y;
  const int tc = threadIdx.x % cv;                   // channel group, fixed (blockDim.x and gridDim.x*blockDim.x % cv == 0)
  const int c = tc * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float* st = stats + ((size_t)n * C + c + j) * 2;
    const float m = st[0] / HW;
    const float var = fmaxf(st[1] / HW - m * m, 0.f);
    sc[j] = style[(size_t)n * 2 * C + c + j] * rsqrtf(var + eps);
    sh[j] = style[(size_t)n * 2 * C + C + c + j] - sc[j] * m;
  }
  const __nv_bfloat16* ab = a + (size_t)n * HW * C + c;
  __nv_bfloat16* xb = x + (size_t)n * HW * C + c;
  const int rows = blockDim.x / cv;
  const int step = gridDim.x * rows;
  int p = blockIdx.x * rows + threadIdx.x / cv;
  for (; p + 3 * step < HW; p += 4 * step) {            // four independent 16-byte loads in flight per thread
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) raw[u] = ld_raw8(ab + (size_t)(p + u * step) * C);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      F8 v = unpack8(raw[u]);
#pragma unroll
      for (int j = 0; j < 8; ++j) v.v[j] = fmaf(v.v[j], sc[j], sh[j]);
      st8(xb + (size_t)(p + u * step) * C, v);
    }
  }
  for (; p < HW; p += step) {
    F8 v = ld8(ab + (size_t)p * C);
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] = fmaf(v.v[j], sc[j], sh[j]);
    st8(xb + (size_t)p * C, v);
  }
}

// gpre = gate(a) * gamma * rstd * (g - S1/HW - ahat * S2/HW)   (instance-norm backward + LeakyReLU gate)
// wsum (optional, zeroed by the launcher): wsum[0][c] += sum gpre (conv bias gradient, gan.py:30),
// wsum[1][c] += sum gpre * noise[n,hw] (InjectSecondaryNoise weight gradient, gan.py:52).
// grid = (blocks per sample, N), fixed (sample, channel group) per thread: with k1 = gamma*rstd, k2 = S1/HW,
// k3 = rstd*S2/HW the inner loop is  v = k1 * (g - k2 - (a - m) * k3)  on registers.
__global__ void adain_bwd_apply_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ a,
                                       const float* __restrict__ stats, const float* __restrict__ style,
                                       const float* __restrict__ bsums, __nv_bfloat16* __restrict__ out, int N,
                                       int HW, int C, float eps, float slope, int gate,
                                       const float* __restrict__ noise, float* __restrict__ wsum) {
  pdl_prologue();
  extern __shared__ float red[];
  const int cv = C / 8;
  const int n = blockIdx.y;
  const int c = (threadIdx.x % cv) * 8;
  const float inv = 1.f / HW;
  float k1[8], k2[8], k3[8], mu[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const size_t sc = ((size_t)n * C + c + j) * 2;
    const float m = stats[sc] * inv;
    const float var = fmaxf(stats[sc + 1] * inv - m * m, 0.f);
    const float rs = rsqrtf(var + eps);
    mu[j] = m;
    k1[j] = style[(size_t)n * 2 * C + c + j] * rs;
    k2[j] = bsums[sc] * inv;
    k3[j] = rs * bsums[sc + 1] * inv;
  }
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  const size_t base = (size_t)n * HW * C + c;
  const int rows = blockDim.x / cv;
  const int step = gridDim.x * rows;
  const bool want_nz = wsum != nullptr && noise != nullptr;
  int p = blockIdx.x * rows + threadIdx.x / cv;
  for (; p + step < HW; p += 2 * step) {                 // two pixels per trip: four 16-byte loads in flight
    uint4 ra[2], rg[2];
    float nz[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      ra[u] = ld_raw8(a + base + (size_t)(p + u * step) * C);
      rg[u] = ld_raw8(g + base + (size_t)(p + u * step) * C);
      nz[u] = want_nz ? noise[(size_t)n * HW + p + u * step] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const F8 av = unpack8(ra[u]), gv = unpack8(rg[u]);
      F8 r;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = k1[j] * (gv.v[j] - k2[j] - (av.v[j] - mu[j]) * k3[j]);
        if (gate) v *= av.v[j] > 0.f ? 1.f : slope;
        r.v[j] = v;
        acc[0][j] += v;
        acc[1][j] = fmaf(v, nz[u], acc[1][j]);
      }
      st8(out + base + (size_t)(p + u * step) * C, r);
    }
  }
  for (; p < HW; p += step) {
    const F8 av = ld8(a + base + (size_t)p * C);
    const F8 gv = ld8(g + base + (size_t)p * C);
    const float nz = want_nz ? noise[(size_t)n * HW + p] : 0.f;
    F8 r;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = k1[j] * (gv.v[j] - k2[j] - (av.v[j] - mu[j]) * k3[j]);
      if (gate) v *= av.v[j] > 0.f ? 1.f : slope;
      r.v[j] = v;
      acc[0][j] += v;
      acc[1][j] = fmaf(v, nz, acc[1][j]);
    }
    st8(out + base + (size_t)p * C, r);
  }
  if (wsum != nullptr) block_channel_reduce<2>(acc, red, wsum, C);
}

// ---------------------------------------------------------------------------------------------
// AdaIN folded into the NEXT conv's operands (generator forward without materialising the normalised map):
//   xo[n,p,ci] = s[n,ci] * a[n,p,ci] + t[n,ci],   s = gamma * rstd,  t = beta - s * mean      (gan.py:65-71)
//   conv3x3(xo)[n,p,co] = sum_{tap,ci} (coef W[co,ci,tap] s[n,ci]) a[n,p+tap,ci]  +  sum_{tap in bounds} sum_ci coef W t
// (the same holds through the bilinear upsample, whose taps sum to 1).  One block per (co, n):
//   wmod[n][tap][co][ci] = bf16(coef * W[co][ci][tap] * s[n][ci])
//   btab[n][cls][co]     = bias[co] + sum over the taps that are inside the image for border class cls = 3*rc + cc
//                          (rc/cc: 0 first row/column, 1 interior, 2 last) of  sum_ci coef * W[co][ci][tap] * t[n][ci]
// ---------------------------------------------------------------------------------------------
__global__ void style_modulate_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                      const float* __restrict__ stats, const float* __restrict__ style,
                                      __nv_bfloat16* __restrict__ wmod, float* __restrict__ btab, int N, int Cin, int Cout,
                                      int HW, float coef, float eps) {
  pdl_prologue();
  __shared__ float red[9][32];
  const int co = blockIdx.x, n = blockIdx.y;
  const float inv = 1.f / HW;
  float acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.f;
  for (int ci = threadIdx.x; ci < Cin; ci += blockDim.x) {
    const float* st = stats + ((size_t)n * Cin + ci) * 2;
    const float m = st[0] * inv;
    const float var = fmaxf(st[1] * inv - m * m, 0.f);
    const float sc = style[(size_t)n * 2 * Cin + ci] * rsqrtf(var + eps);
    const float sh = style[(size_t)n * 2 * Cin + Cin + ci] - sc * m;
    const float* w = W + ((size_t)co * Cin + ci) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float wk = w[k] * coef;
      wmod[(((size_t)n * 9 + k) * Cout + co) * Cin + ci] = __float2bfloat16_rn(wk * sc);
      acc[k] = fmaf(wk, sh, acc[k]);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float v = warp_sum(acc[k]);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    const int cls = threadIdx.x, rc = cls / 3, cc = cls % 3;
    const int nw = blockDim.x >> 5;
    float tot = bias != nullptr ? bias[co] : 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      if ((rc == 0 && ky == 0) || (rc == 2 && ky == 2)) continue;      // that tap row reads the zero padding
      for (int kx = 0; kx < 3; ++kx) {
        if ((cc == 0 && kx == 0) || (cc == 2 && kx == 2)) continue;
        for (int wv = 0; wv < nw; ++wv) tot += red[ky * 3 + kx][wv];
      }
    }
    btab[((size_t)n * 9 + cls) * Cout + co] = tot;
  }
}

// toRGB (1x1 conv to 3 planes, gan.py:172-179) applied to AdaIN(a) without materialising it: per sample the AdaIN
// scale is folded into the (3 x C) matrix and the shift into the bias.  grid = (blocks per sample, N).
__global__ void to_rgb_adain_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ stats,
                                    const float* __restrict__ style, const float* __restrict__ Wm,
                                    const float* __restrict__ bias, float* __restrict__ out, int HW, int C, float coef,
                                    float eps) {
  pdl_prologue();
  extern __shared__ float sw[];  // [3][C] + [3]
  const int n = blockIdx.y;
  const float inv = 1.f / HW;
  float part[3] = {0.f, 0.f, 0.f};
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* st = stats + ((size_t)n * C + c) * 2;
    const float m = st[0] * inv;
    const float var = fmaxf(st[1] * inv - m * m, 0.f);
    const float sc = style[(size_t)n * 2 * C + c] * rsqrtf(var + eps);
    const float sh = style[(size_t)n * 2 * C + C + c] - sc * m;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float w = Wm[(size_t)j * C + c] * coef;
      sw[j * C + c] = w * sc;
      part[j] = fmaf(w, sh, part[j]);
    }
  }
  __shared__ float redb[3][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float v = warp_sum(part[j]);
    if (lane == 0) redb[j][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float tot = bias != nullptr ? bias[threadIdx.x] : 0.f;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) tot += redb[threadIdx.x][wv];
    sw[3 * C + threadIdx.x] = tot;
  }
  __syncthreads();
  const float b0 = sw[3 * C], b1 = sw[3 * C + 1], b2 = sw[3 * C + 2];
  const int cv = C / 8;
  const int lanes_per_pix = cv < 32 ? cv : 32;
  const int pix_per_warp = 32 / lanes_per_pix;
  const int sub = lane % lanes_per_pix;
  const int pw = lane / lanes_per_pix;
  const int warps_per_block = blockDim.x >> 5;
  const int groups = (HW + pix_per_warp - 1) / pix_per_warp;
  const __nv_bfloat16* ab = a + (size_t)n * HW * C;
  float* ob = out + (size_t)n * 3 * HW;
  for (int gidx = blockIdx.x * warps_per_block + warp; gidx < groups; gidx += gridDim.x * warps_per_block) {
    const int p = gidx * pix_per_warp + pw;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (p < HW) {
      for (int v = sub; v < cv; v += lanes_per_pix) {
        const F8 x = ld8(ab + (size_t)p * C + v * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = v * 8 + j;
          s0 = fmaf(x.v[j], sw[c], s0);
          s1 = fmaf(x.v[j], sw[C + c], s1);
          s2 = fmaf(x.v[j], sw[2 * C + c], s2);
        }
      }
    }
    for (int o = lanes_per_pix >> 1; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (p < HW && sub == 0) {
      ob[p] = s0 + b0;
      ob[p + HW] = s1 + b1;
      ob[p + 2 * (size_t)HW] = s2 + b2;
    }
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
int launch_pack_weight(const float* w, void* wf, void* wd, int Cout, int Cin, int Cin_pad, int ks, float coef,
                       cudaStream_t s) {
  BG_REQUIRE(ks == 1 || ks == 3, "pack_weight: ks must be 1 or 3");
  BG_REQUIRE(Cin_pad >= Cin, "pack_weight: Cin_pad < Cin");
  const size_t total = (size_t)ks * ks * Cout * Cin_pad;
  BG_CHECK_CUDA(launch_pdl(pack_weight_kernel, grid_for(total), kBlock, 0, s, w, (__nv_bfloat16*)wf, (__nv_bfloat16*)wd,
                           Cout, Cin, Cin_pad, ks, coef));
  return 0;
}

int launch_unpack_wgrad(const float* dwp, float* dw, int Cout, int Cin, int Cin_pad, int ks, float coef,
                        int accumulate, cudaStream_t s) {
  const size_t total = (size_t)ks * ks * Cout * Cin;
  if (ks == 3 && Cout <= 65535) {
    BG_CHECK_CUDA(launch_pdl(unpack_wgrad_tiled_kernel<0>, dim3((Cin + 127) / 128, Cout), 128, 0, s, dwp, dw, Cout, Cin,
                             Cin_pad, coef, accumulate));
    return 0;
  }
  BG_CHECK_CUDA(launch_pdl(unpack_wgrad_kernel, grid_for(total), kBlock, 0, s, dwp, dw, Cout, Cin, Cin_pad, ks, coef,
                           accumulate));
  return 0;
}

int launch_act_gate(const void* g, const void* y, void* out, size_t n, float slope, cudaStream_t s) {
  BG_REQUIRE(n % 8 == 0, "act_gate: element count must be a multiple of 8");
  BG_CHECK_CUDA(launch_pdl(act_gate_kernel, grid_for(n / 8), kBlock, 0, s, (const __nv_bfloat16*)g,
                           (const __nv_bfloat16*)y, (__nv_bfloat16*)out, n / 8, slope));
  return 0;
}

int launch_axpby(const void* a, const void* b, void* out, size_t n, float ca, float cb, cudaStream_t s) {
  BG_REQUIRE(n % 8 == 0, "axpby: element count must be a multiple of 8");
  BG_CHECK_CUDA(launch_pdl(axpby_kernel, grid_for(n / 8), kBlock, 0, s, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                           (__nv_bfloat16*)out, n / 8, ca, cb));
  return 0;
}

int launch_pool_act_fwd(const void* u, const void* gate_src, void* y, int N, int Ho, int Wo, int C, float slope,
                        int mode, cudaStream_t s) {
  BG_REQUIRE(C % 8 == 0, "pool_act_fwd: C must be a multiple of 8");
  BG_REQUIRE(mode == 0 || gate_src != nullptr, "pool_act_fwd: mode 1 needs gate_src");
  const size_t total = (size_t)N * Ho * Wo * (C / 8);
  BG_REQUIRE(total < (1ull << 32), "pool_act_fwd: map too large");
  BG_CHECK_CUDA(launch_pdl(pool_act_fwd_kernel, grid_for(total), kBlock, 0, s, (const __nv_bfloat16*)u,
                           (const __nv_bfloat16*)gate_src, (__nv_bfloat16*)y, N, Ho, Wo, C, slope, mode, make_div(C / 8),
                           make_div(Wo), make_div(Ho)));
  return 0;
}

int launch_pool_act_bwd(const void* gy, const void* y, void* gu, int N, int Ho, int Wo, int C, float slope,
                        float* csum, cudaStream_t s) {
  BG_REQUIRE(C % 8 == 0, "pool_act_bwd: C must be a multiple of 8");
  const size_t total = (size_t)N * Ho * Wo * (C / 8);
  size_t smem = 0;
  if (csum != nullptr) {
    BG_REQUIRE(kBlock % (C / 8) == 0, "pool_act_bwd: fused bias-gradient sum needs C/8 to divide %d (C %d)", kBlock, C);
    if (launch_zero(csum, (size_t)C * sizeof(float), s) != 0) return 1;
    smem = (size_t)kBlock * 8 * sizeof(float);
  }
  BG_REQUIRE(total < (1ull << 32), "pool_act_bwd: map too large");
  BG_CHECK_CUDA(launch_pdl(pool_act_bwd_kernel, grid_for(total), kBlock, smem, s, (const __nv_bfloat16*)gy,
                           (const __nv_bfloat16*)y, (__nv_bfloat16*)gu, N, Ho, Wo, C, slope, csum, make_div(C / 8),
                           make_div(Wo), make_div(Ho)));
  return 0;
}

int launch_upsample2x_fwd(const void* x, void* y, int N, int H, int W, int C, cudaStream_t s) {
  BG_REQUIRE(C % 8 == 0, "upsample2x_fwd: C must be a multiple of 8");
  const size_t total = (size_t)N * H * W * (C / 8);
  BG_REQUIRE(total < (1ull << 32), "upsample2x_fwd: map too large");
  BG_CHECK_CUDA(launch_pdl(upsample2x_fwd_kernel, grid_for(total), kBlock, 0, s, (const __nv_bfloat16*)x,
                           (__nv_bfloat16*)y, N, H, W, C, make_div(C / 8), make_div(W), make_div(H)));
  return 0;
}

int launch_upsample2x_bwd(const void* gy, void* gx, int N, int H, int W, int C, cudaStream_t s) {
  BG_REQUIRE(C % 8 == 0 && H % kUpStrip == 0, "upsample2x_bwd: C must be a multiple of 8 and H of %d", kUpStrip);
  BG_REQUIRE((size_t)N * H * W * (C / 8) < (1ull << 32), "upsample2x_bwd: map too large");
  const size_t total = (size_t)N * (H / kUpStrip) * W * (C / 8);
  BG_CHECK_CUDA(launch_pdl(upsample2x_bwd_kernel, grid_for(total), kBlock, 0, s, (const __nv_bfloat16*)gy,
                           (__nv_bfloat16*)gx, N, H, W, C, make_div(C / 8), make_div(W), make_div(H / kUpStrip)));
  return 0;
}

int launch_channel_wsum(const void* g, const float* planes, float* out, size_t P, int C, int HW, size_t img_stride,
                        size_t plane_stride, int nplanes, cudaStream_t s) {
  BG_REQUIRE(C % 8 == 0 && C <= 1024, "channel_wsum: unsupported C %d", C);
  BG_REQUIRE(nplanes >= 0 && nplanes <= 3, "channel_wsum: nplanes must be 0..3");
  const int cv = C / 8;
  const int threads = cv >= 256 ? cv : 256;
  const int rows = threads / cv;
  if (launch_zero(out, (size_t)(1 + nplanes) * C * sizeof(float), s) != 0) return 1;
  // enough blocks to fill the chip, but at least `rows * 8` pixels each
  const size_t smem = (size_t)rows * 4 * C * sizeof(float);
  BG_REQUIRE(smem <= 48 * 1024, "channel_wsum: shared memory %zu too large", smem);
  static int cap_cache[2] = {0, 0};
  if (cap_cache[1] == 0 || cap_cache[0] != (int)smem) {
    cap_cache[1] = resident_blocks(channel_wsum_kernel, threads, smem);
    cap_cache[0] = (int)smem;
  }
  const size_t cap = (size_t)cap_cache[1];                  // one resident wave, equal shares
  size_t ppb = (P + cap - 1) / cap;
  if (ppb < (size_t)rows * 8) ppb = (size_t)rows * 8;
  const size_t blocks = (P + ppb - 1) / ppb;
  int hw_shift = -1;
  if (HW > 0 && (HW & (HW - 1)) == 0) {
    hw_shift = 0;
    while ((1 << hw_shift) < HW) ++hw_shift;
  }
  BG_CHECK_CUDA(launch_pdl(channel_wsum_kernel, (int)blocks, threads, smem, s, (const __nv_bfloat16*)g, planes, out, P,
                           C, HW, img_stride, plane_stride, nplanes, (int)ppb, hw_shift));
  return 0;
}

int launch_planes3_to_nhwc(const float* img, const float* Wm, const float* bias, const void* gate_src, void* out,
                           size_t P, int HW, int C, int ws_c, int ws_j, float coef, int act, float slope,
                           cudaStream_t s) {
  BG_REQUIRE(C % 8 == 0 && C <= 1024, "planes3_to_nhwc: unsupported C %d", C);
  if (HW % 4 == 0 && P % 4 == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0 && P / 4 * (C / 8) < (1ull << 32)) {
    BG_CHECK_CUDA(launch_pdl(planes3_to_nhwc_quad_kernel, grid_for(P / 4 * (C / 8)), kBlock,
                             (size_t)C * 4 * sizeof(float), s, img, Wm, bias, (const __nv_bfloat16*)gate_src,
                             (__nv_bfloat16*)out, P, HW, C, ws_c, ws_j, coef, act, slope, make_div(C / 8),
                             make_div(HW / 4)));
    return 0;
  }
  const size_t total = P * (C / 8);
  BG_CHECK_CUDA(launch_pdl(planes3_to_nhwc_kernel, grid_for(total), kBlock, (size_t)C * 4 * sizeof(float), s, img, Wm,
                           bias, (const __nv_bfloat16*)gate_src, (__nv_bfloat16*)out, P, HW, C, ws_c, ws_j, coef, act,
                           slope, make_div(C / 8), make_div(HW)));
  return 0;
}

int launch_nhwc_to_planes3(const void* x, const float* Wm, const float* bias, float* out, size_t P, int HW, int C,
                           int ws_c, int ws_j, float coef, cudaStream_t s) {
  BG_REQUIRE(C % 16 == 0 && C <= 1024, "nhwc_to_planes3: unsupported C %d", C);
  const int cv = C / 8;
  const int lanes_per_pix = cv < 32 ? cv : 32;
  const size_t groups = (P + (32 / lanes_per_pix) - 1) / (32 / lanes_per_pix);
  const int grid = grid_for((groups + kNhwcUnroll - 1) / kNhwcUnroll * 32);
  if (cv <= 32)
    BG_CHECK_CUDA(launch_pdl(nhwc_to_planes3_kernel<true>, grid, kBlock, (size_t)C * 3 * sizeof(float), s,
                             (const __nv_bfloat16*)x, Wm, bias, out, P, HW, C, ws_c, ws_j, coef));
  else
    BG_CHECK_CUDA(launch_pdl(nhwc_to_planes3_kernel<false>, grid, kBlock, (size_t)C * 3 * sizeof(float), s,
                             (const __nv_bfloat16*)x, Wm, bias, out, P, HW, C, ws_c, ws_j, coef));
  return 0;
}

static int in_reduce_launch(const void* a, const void* g, const float* stats, float* sums, int N, int HW, int C,
                            float eps, int mode, cudaStream_t s) {
  BG_REQUIRE(C % 8 == 0 && C <= 1024, "in_reduce: unsupported C %d", C);
  const int cv = C / 8;
  const int threads = cv >= 256 ? cv : 256;
  const int rows = threads / cv;
  if (launch_zero(sums, (size_t)N * C * 2 * sizeof(float), s) != 0) return 1;
  const size_t smem = (size_t)rows * 2 * C * sizeof(float);
  BG_REQUIRE(smem <= 48 * 1024, "in_reduce: shared memory %zu too large", smem);
  static int cap_cache[2] = {0, 0};                         // [0] = smem bytes the capacity was computed for
  if (cap_cache[1] == 0 || cap_cache[0] != (int)smem) {
    cap_cache[1] = resident_blocks(in_reduce_kernel, threads, smem);
    cap_cache[0] = (int)smem;
  }
  int blocks_per_img = cap_cache[1] / N;                    // floor: the grid never exceeds one resident wave
  int max_bpi = (HW + rows * 4 - 1) / (rows * 4);
  if (blocks_per_img > max_bpi) blocks_per_img = max_bpi;
  if (blocks_per_img < 1) blocks_per_img = 1;
  if (deterministic()) blocks_per_img = 1;                  // one contributor per (sample, channel): ordered sums only
  const int ppb = (HW + blocks_per_img - 1) / blocks_per_img;
  blocks_per_img = (HW + ppb - 1) / ppb;
  BG_CHECK_CUDA(launch_pdl(in_reduce_kernel, N * blocks_per_img, threads, smem, s, (const __nv_bfloat16*)a,
                           (const __nv_bfloat16*)g, stats, sums, HW, C, ppb, blocks_per_img, eps, mode));
  return 0;
}

int launch_in_stats(const void* a, float* sums, int N, int HW, int C, cudaStream_t s) {
  return in_reduce_launch(a, nullptr, nullptr, sums, N, HW, C, 0.f, 0, s);
}

int launch_adain_bwd_reduce(const void* g, const void* a, const float* stats, float* bsums, int N, int HW, int C,
                            float eps, cudaStream_t s) {
  return in_reduce_launch(a, g, stats, bsums, N, HW, C, eps, 1, s);
}

// blocks per sample for the sample-aligned elementwise kernels: ~8 blocks per SM overall, at least 8 pixel rows each
static int blocks_per_sample(int N, int HW, int C) {
  const int rows = kBlock / (C / 8);
  int bps = (num_sms() * 8 + N - 1) / N;
  const int max_bps = (HW + rows * 8 - 1) / (rows * 8);
  if (bps > max_bps) bps = max_bps;
  return bps < 1 ? 1 : bps;
}

int launch_adain_apply(const void* a, const float* stats, const float* style, void* x, int N, int HW, int C, float eps,
                       cudaStream_t s) {
  BG_REQUIRE(C % 8 == 0 && kBlock % (C / 8) == 0, "adain_apply: C/8 must divide %d (C %d)", kBlock, C);
  BG_CHECK_CUDA(launch_pdl(adain_apply_kernel, dim3(blocks_per_sample(N, HW, C), N), kBlock, 0, s,
                           (const __nv_bfloat16*)a, stats, style, (__nv_bfloat16*)x, N, HW, C, eps));
  return 0;
}

int launch_adain_bwd_apply(const void* g, const void* a, const float* stats, const float* style, const float* bsums,
                           void* out, int N, int HW, int C, float eps, float slope, int gate, const float* noise,
                           float* wsum, cudaStream_t s) {
  BG_REQUIRE(C % 8 == 0 && kBlock % (C / 8) == 0, "adain_bwd_apply: C/8 must divide %d (C %d)", kBlock, C);
  size_t smem = 0;
  if (wsum != nullptr) {
    if (launch_zero(wsum, (size_t)2 * C * sizeof(float), s) != 0) return 1;
    smem = (size_t)kBlock * 16 * sizeof(float);
  }
  BG_CHECK_CUDA(launch_pdl(adain_bwd_apply_kernel, dim3(blocks_per_sample(N, HW, C), N), kBlock, smem, s,
                           (const __nv_bfloat16*)g, (const __nv_bfloat16*)a, stats, style, bsums, (__nv_bfloat16*)out, N,
                           HW, C, eps, slope, gate, noise, wsum));
  return 0;
}

int launch_style_modulate(const float* W, const float* bias, const float* stats, const float* style, void* wmod,
                          float* btab, int N, int Cin, int Cout, int HW, float coef, float eps, cudaStream_t s) {
  BG_REQUIRE(N > 0 && Cin > 0 && Cout > 0 && HW > 0, "style_modulate: bad shape N %d Cin %d Cout %d HW %d", N, Cin, Cout, HW);
  const int threads = Cin >= 256 ? 256 : (Cin >= 64 ? 64 : 32);
  BG_CHECK_CUDA(launch_pdl(style_modulate_kernel, dim3(Cout, N), threads, 0, s, W, bias, stats, style,
                           (__nv_bfloat16*)wmod, btab, N, Cin, Cout, HW, coef, eps));
  return 0;
}

int launch_to_rgb_adain(const void* a, const float* stats, const float* style, const float* Wm, const float* bias,
                        float* out, int N, int HW, int C, float coef, float eps, cudaStream_t s) {
  BG_REQUIRE(C % 16 == 0 && C <= 1024, "to_rgb_adain: unsupported C %d", C);
  int bps = (num_sms() * 8 + N - 1) / N;                    // blocks per sample
  const int cv = C / 8;
  const int ppw = 32 / (cv < 32 ? cv : 32);
  const int max_bps = (HW / ppw + 63) / 64;                 // at least ~8 pixel groups per warp
  if (bps > max_bps) bps = max_bps;
  if (bps < 1) bps = 1;
  BG_CHECK_CUDA(launch_pdl(to_rgb_adain_kernel, dim3(bps, N), kBlock, (size_t)(3 * C + 4) * sizeof(float), s,
                           (const __nv_bfloat16*)a, stats, style, Wm, bias, out, HW, C, coef, eps));
  return 0;
}

int launch_pack_weight_grouped(const float* const* w, void* const* wf, void* const* wd, const int* Cout, const int* Cin,
                               const int* Cin_pad, const int* ks, const float* coef, int groups, cudaStream_t s) {
  BG_REQUIRE(groups > 0 && groups <= kMaxPackGroups, "pack_weight_grouped: 1..%d groups (got %d)", kMaxPackGroups, groups);
  PackGroups G;
  memset(&G, 0, sizeof(G));
  G.groups = groups;
  int blocks = 0;
  for (int g = 0; g < groups; ++g) {
    BG_REQUIRE((ks[g] == 1 || ks[g] == 3) && Cin_pad[g] >= Cin[g], "pack_weight_grouped: bad group %d", g);
    G.w[g] = w[g];
    G.wf[g] = (__nv_bfloat16*)wf[g];
    G.wd[g] = (__nv_bfloat16*)wd[g];
    G.Cout[g] = Cout[g]; G.Cin[g] = Cin[g]; G.Cin_pad[g] = Cin_pad[g]; G.ks[g] = ks[g];
    G.coef[g] = coef[g];
    G.blk0[g] = blocks;
    const size_t total = (size_t)ks[g] * ks[g] * Cout[g] * Cin_pad[g];
    size_t b = (total + (size_t)kBlock * 8 - 1) / ((size_t)kBlock * 8);      // ~8 elements per thread
    if (b < 1) b = 1;
    if (b > 1184) b = 1184;
    blocks += (int)b;
  }
  G.blk0[groups] = blocks;
  bool all3 = true;                    // the tiled kernel: 3x3 layers with 16-byte-aligned rows on both packs
  for (int g = 0; g < groups; ++g)
    all3 = all3 && ks[g] == 3 && Cout[g] % 16 == 0 && Cin_pad[g] % 8 == 0 &&
           ((reinterpret_cast<uintptr_t>(wf[g]) | reinterpret_cast<uintptr_t>(wd[g])) & 15) == 0;
  if (all3) {
    int tiles = 0;
    for (int g = 0; g < groups; ++g) {
      G.blk0[g] = tiles;
      tiles += ((Cout[g] + kPackTile - 1) / kPackTile) * ((Cin_pad[g] + kPackTile - 1) / kPackTile);
    }
    G.blk0[groups] = tiles;
    const size_t smem = (size_t)9 * kPackTapStride * sizeof(__nv_bfloat16);
    static bool attr_set = false;
    if (!attr_set) {
      BG_CHECK_CUDA(cudaFuncSetAttribute(pack_weight_grouped_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    BG_CHECK_CUDA(launch_pdl(pack_weight_grouped_tiled_kernel, tiles, 256, smem, s, G));
    return 0;
  }
  BG_CHECK_CUDA(launch_pdl(pack_weight_grouped_kernel, blocks, kBlock, 0, s, G));
  return 0;
}

int launch_pack_weight_pool4(const float* w, void* w16, int Cout, int Cin, float coef, cudaStream_t s) {
  BG_REQUIRE(Cout > 0 && Cin > 0, "pack_weight_pool4: bad shape");
  BG_CHECK_CUDA(launch_pdl(pack_weight_pool4_kernel, grid_for((size_t)16 * Cout * Cin), kBlock, 0, s, w,
                           (__nv_bfloat16*)w16, Cout, Cin, coef));
  return 0;
}

int launch_pack_weight_tconv4(const float* w, void* wt, int Cout, int Cin, float coef, cudaStream_t s) {
  BG_REQUIRE(Cout > 0 && Cin > 0, "pack_weight_tconv4: bad shape");
  BG_CHECK_CUDA(launch_pdl(pack_weight_tconv4_kernel, grid_for((size_t)16 * Cout * Cin), kBlock, 0, s, w,
                           (__nv_bfloat16*)wt, Cout, Cin, coef));
  return 0;
}

int launch_unpack_wgrad_pool4(const float* dw4, float* dw, int Cout, int Cin, float coef, int accumulate, cudaStream_t s) {
  if (Cout <= 65535) {
    BG_CHECK_CUDA(launch_pdl(unpack_wgrad_tiled_kernel<1>, dim3((Cin + 127) / 128, Cout), 128, 0, s, dw4, dw, Cout, Cin, Cin,
                             coef, accumulate));
    return 0;
  }
  BG_CHECK_CUDA(launch_pdl(unpack_wgrad_pool4_kernel, grid_for((size_t)Cout * Cin * 9), kBlock, 0, s, dw4, dw, Cout, Cin,
                           coef, accumulate));
  return 0;
}

}  // namespace bg
