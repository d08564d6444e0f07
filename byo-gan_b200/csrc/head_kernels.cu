// Small fp32 kernels around the convolution stack: EqualizedLinear (mapping network, style FCs, critic
// head), the learned constant, image-plane fade ops, minibatch-stddev (forward / tangent / backward /
// second order) and the loss terms.  All of these touch KBs..MBs; they are latency- and launch-bound, so
// the design goal is one launch per logical op with enough CTAs to cover the chip, fp32 throughout.
#include "common.cuh"

namespace bg {

namespace {

__device__ __forceinline__ float lrelu_f(float v, float slope) { return v > 0.f ? v : v * slope; }
__device__ __forceinline__ float gate_f(float y, float slope) { return y > 0.f ? 1.f : slope; }

// ---------------------------------------------------------------------------------------------
// EqualizedLinear.forward (gan.py:16-17):  y[m,n] = act(coef * sum_k x[m,k] W[n,k] + b[n])
// One warp per (n, group of MT rows): W[n,:] is streamed once per warp with float4 loads.
// ---------------------------------------------------------------------------------------------
constexpr int kLinMT = 8;

__global__ void linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                  const float* __restrict__ bias, float* __restrict__ y, int M, int N, int K,
                                  float coef, int act, float slope) {
  pdl_prologue();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int mgroups = (M + kLinMT - 1) / kLinMT;
  if (warp >= N * mgroups) return;
  const int n = warp % N;
  const int m0 = (warp / N) * kLinMT;
  float acc[kLinMT];
#pragma unroll
  for (int i = 0; i < kLinMT; ++i) acc[i] = 0.f;
  const float* wrow = W + (size_t)n * K;
  if ((K & 3) == 0) {
    for (int k = lane * 4; k < K; k += 128) {
      const float4 wv = *reinterpret_cast<const float4*>(wrow + k);
#pragma unroll
      for (int i = 0; i < kLinMT; ++i) {
        if (m0 + i < M) {
          const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)(m0 + i) * K + k);
          acc[i] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
        }
      }
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      const float wv = wrow[k];
#pragma unroll
      for (int i = 0; i < kLinMT; ++i)
        if (m0 + i < M) acc[i] += x[(size_t)(m0 + i) * K + k] * wv;
    }
  }
#pragma unroll
  for (int i = 0; i < kLinMT; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
    const float b = bias ? bias[n] : 0.f;
#pragma unroll
    for (int i = 0; i < kLinMT; ++i) {
      if (m0 + i < M) {
        float v = acc[i] * coef + b;
        if (act) v = lrelu_f(v, slope);
        y[(size_t)(m0 + i) * N + n] = v;
      }
    }
  }
}

// dW[n,k] (+)= coef * sum_m gy[m,n] x[m,k];  db[n] (+)= sum_m gy[m,n]
constexpr int kLbwNT = 8;
__global__ void linear_bwd_weight_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                         float* __restrict__ dW, float* __restrict__ db, int M, int N, int K,
                                         float coef, int accumulate) {
  pdl_prologue();
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int n0 = blockIdx.y * kLbwNT;
  float acc[kLbwNT][4];
#pragma unroll
  for (int i = 0; i < kLbwNT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  if (k < K) {
#pragma unroll 8
    for (int m = 0; m < M; ++m) {                    // rows are independent loads: keep 8 in flight (M is the batch)
      const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)m * K + k);
#pragma unroll
      for (int i = 0; i < kLbwNT; ++i) {
        const float g = (n0 + i < N) ? gy[(size_t)m * N + n0 + i] : 0.f;
        acc[i][0] += g * xv.x;
        acc[i][1] += g * xv.y;
        acc[i][2] += g * xv.z;
        acc[i][3] += g * xv.w;
      }
    }
#pragma unroll
    for (int i = 0; i < kLbwNT; ++i) {
      if (n0 + i < N) {
        float4* dst = reinterpret_cast<float4*>(dW + (size_t)(n0 + i) * K + k);
        float4 v = make_float4(acc[i][0] * coef, acc[i][1] * coef, acc[i][2] * coef, acc[i][3] * coef);
        if (accumulate) {
          const float4 o = *dst;
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        *dst = v;
      }
    }
  }
  if (db != nullptr && blockIdx.x == 0 && threadIdx.x < kLbwNT && n0 + threadIdx.x < N) {
    float s = 0.f;
#pragma unroll 8
    for (int m = 0; m < M; ++m) s += gy[(size_t)m * N + n0 + threadIdx.x];
    db[n0 + threadIdx.x] = accumulate ? db[n0 + threadIdx.x] + s : s;
  }
}

// ---------------------------------------------------------------------------------------------
// Grouped EqualizedLinear: the 2 x steps AdaIN style FCs of the generator (gan.py:60,66) all read the SAME latent
// w (M x K), so forward, weight gradient and input gradient of all of them are one launch each instead of 14 (the
// layers are tiny and each launch is latency-bound).  Groups are described by a by-value table of pointers.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxLinGroups = 16;
struct LinGroups {
  const float* W[kMaxLinGroups];     // (N_g, K) weights
  const float* x[kMaxLinGroups];     // bwd_weight: the layer's input (M, K)
  const float* b[kMaxLinGroups];     // fwd: bias (may be null)
  float* y[kMaxLinGroups];           // fwd: output (M, N_g);  bwd_*: gy (M, N_g) (read)
  float* dW[kMaxLinGroups];          // bwd_weight: (N_g, K)
  float* db[kMaxLinGroups];          // bwd_weight: (N_g) (may be null)
  int N[kMaxLinGroups];
  int blk0[kMaxLinGroups + 1];       // first block of each group (fwd / bwd_weight)
  float coef[kMaxLinGroups];
  int groups;
};

// fwd: one warp per (n, group of kLinMT rows) of one group; same inner loop as linear_fwd_kernel
__global__ void linear_fwd_grouped_kernel(const float* __restrict__ x, const LinGroups G, int M, int K, int act,
                                          float slope) {
  pdl_prologue();
  int g = 0;
  while (g + 1 < G.groups && (int)blockIdx.x >= G.blk0[g + 1]) ++g;
  const int N = G.N[g];
  const int warp = (((int)blockIdx.x - G.blk0[g]) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int mgroups = (M + kLinMT - 1) / kLinMT;
  if (warp >= N * mgroups) return;
  const int n = warp % N;
  const int m0 = (warp / N) * kLinMT;
  float acc[kLinMT];
#pragma unroll
  for (int i = 0; i < kLinMT; ++i) acc[i] = 0.f;
  const float* wrow = G.W[g] + (size_t)n * K;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 wv = *reinterpret_cast<const float4*>(wrow + k);
#pragma unroll
    for (int i = 0; i < kLinMT; ++i) {
      if (m0 + i < M) {
        const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)(m0 + i) * K + k);
        acc[i] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kLinMT; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
    const float b = G.b[g] ? G.b[g][n] : 0.f;
#pragma unroll
    for (int i = 0; i < kLinMT; ++i) {
      if (m0 + i < M) {
        float v = acc[i] * G.coef[g] + b;
        if (act) v = lrelu_f(v, slope);
        G.y[g][(size_t)(m0 + i) * N + n] = v;
      }
    }
  }
}

// bwd_weight: blocks of group g cover (K/4 threads) x (N_g / kLbwNT row groups); dW = coef * gy^T x, db = sum_m gy
__global__ void linear_bwd_weight_grouped_kernel(const LinGroups G, int M, int K) {
  pdl_prologue();
  int g = 0;
  while (g + 1 < G.groups && (int)blockIdx.x >= G.blk0[g + 1]) ++g;
  const int N = G.N[g];
  const float* __restrict__ gy = G.y[g];
  const float* __restrict__ x = G.x[g];
  const int kblocks = (K / 4 + 127) / 128;
  const int local = (int)blockIdx.x - G.blk0[g];
  const int k = ((local % kblocks) * blockDim.x + threadIdx.x) * 4;
  const int n0 = (local / kblocks) * kLbwNT;
  float acc[kLbwNT][4];
#pragma unroll
  for (int i = 0; i < kLbwNT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  if (k < K) {
#pragma unroll 8
    for (int m = 0; m < M; ++m) {
      const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)m * K + k);
#pragma unroll
      for (int i = 0; i < kLbwNT; ++i) {
        const float gg = (n0 + i < N) ? gy[(size_t)m * N + n0 + i] : 0.f;
        acc[i][0] += gg * xv.x;
        acc[i][1] += gg * xv.y;
        acc[i][2] += gg * xv.z;
        acc[i][3] += gg * xv.w;
      }
    }
    const float coef = G.coef[g];
#pragma unroll
    for (int i = 0; i < kLbwNT; ++i) {
      if (n0 + i < N)
        *reinterpret_cast<float4*>(G.dW[g] + (size_t)(n0 + i) * K + k) =
            make_float4(acc[i][0] * coef, acc[i][1] * coef, acc[i][2] * coef, acc[i][3] * coef);
    }
  }
  if (G.db[g] != nullptr && (local % kblocks) == 0 && threadIdx.x < kLbwNT && n0 + threadIdx.x < N) {
    float sacc = 0.f;
#pragma unroll 8
    for (int m = 0; m < M; ++m) sacc += gy[(size_t)m * N + n0 + threadIdx.x];
    G.db[g][n0 + threadIdx.x] = sacc;
  }
}

// out[c][r] = in[r][c]  (fp32), 32x32 smem tiles
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[(size_t)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) out[(size_t)c * R + r] = tile[threadIdx.x][i];
  }
}

// out = g * gate(y) (+ add)   fp32 vectors (LeakyReLU backward on the small fp32 paths)
__global__ void act_gate_f32_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                    float* __restrict__ out, size_t n, float slope) {
  pdl_prologue();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = g[i] * gate_f(y[i], slope);
}

// out = ca*a + cb*b  fp32 (b may be null)
__global__ void axpby_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                 size_t n, float ca, float cb) {
  pdl_prologue();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = ca * a[i] + (b ? cb * b[i] : 0.f);
}

// ---------------------------------------------------------------------------------------------
// learned constant + noise + LeakyReLU (StyleConvBlock with is_initial, gan.py:81,92,96-97)
//   a[n,h,w,c] = lrelu(const[c,h,w] + nw[c] * noise[n,h,w]);   HW = 16
// ---------------------------------------------------------------------------------------------
__global__ void const_noise_act_kernel(const float* __restrict__ cst, const float* __restrict__ noise,
                                       const float* __restrict__ nw, __nv_bfloat16* __restrict__ a, int N, int HW,
                                       int C, float slope) {
  pdl_prologue();
  const size_t total = (size_t)N * HW * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t p = i / C;
    const int hw = (int)(p % HW);
    const float v = cst[(size_t)c * HW + hw] + nw[c] * noise[p];
    a[i] = __float2bfloat16_rn(lrelu_f(v, slope));
  }
}
// dconst[c,hw] = sum_n g[n,hw,c]
__global__ void const_bwd_kernel(const __nv_bfloat16* __restrict__ g, float* __restrict__ dconst, int N, int HW,
                                 int C) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW * C) return;
  const int c = i % C, hw = i / C;
  float s = 0.f;
  for (int n = 0; n < N; ++n) s += __bfloat162float(g[((size_t)n * HW + hw) * C + c]);
  dconst[(size_t)c * HW + hw] = s;
}

// ---------------------------------------------------------------------------------------------
// image-plane ops, NCHW fp32 (B,3,R,R) viewed as P = B*3 planes
// ---------------------------------------------------------------------------------------------
// F.avg_pool2d(images, 2) (gan.py:345)
__global__ void img_avgpool2_kernel(const float* __restrict__ img, float* __restrict__ out, int P, int Ho, int Wo) {
  pdl_prologue();
  const size_t total = (size_t)P * Ho * Wo;
  const int W = 2 * Wo;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int wo = (int)(i % Wo);
    const int ho = (int)((i / Wo) % Ho);
    const size_t pl = i / ((size_t)Wo * Ho);
    const float* b = img + (pl * 2 * Ho + 2 * ho) * W + 2 * wo;
    out[i] = 0.25f * (b[0] + b[1] + b[W] + b[W + 1]);
  }
}
// gimg[2h+dy,2w+dx] (+)= 0.25 * scale * g[h,w]
__global__ void img_avgpool2_bwd_kernel(const float* __restrict__ g, float* __restrict__ gimg, int P, int Ho, int Wo,
                                        float scale, int accumulate) {
  pdl_prologue();
  const int W = 2 * Wo, H = 2 * Ho;
  const size_t total = (size_t)P * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const size_t pl = i / ((size_t)W * H);
    const float v = 0.25f * scale * g[(pl * Ho + (h >> 1)) * Wo + (w >> 1)];
    gimg[i] = accumulate ? gimg[i] + v : v;
  }
}
__device__ __forceinline__ void up_taps(int o, int L, int& i0, int& i1, float& w0, float& w1) {
  // bilinear x2, align_corners=False: src = (o + .5)/2 - .5, clamped
  const int h = o >> 1;
  i0 = h;
  i1 = (o & 1) ? min(h + 1, L - 1) : max(h - 1, 0);
  w0 = 0.75f;
  w1 = 0.25f;
}
// out = (1-alpha) * bilinear_up2(small) + alpha * large       (gan.py:213-220)
__global__ void img_up2_lerp_kernel(const float* __restrict__ small, const float* __restrict__ large,
                                    float* __restrict__ out, int P, int H, int W, float alpha) {
  pdl_prologue();
  const int Ho = 2 * H, Wo = 2 * W;
  const size_t total = (size_t)P * Ho * Wo;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int wo = (int)(i % Wo);
    const int ho = (int)((i / Wo) % Ho);
    const size_t pl = i / ((size_t)Wo * Ho);
    int h0, h1, w0, w1;
    float a0, a1, b0, b1;
    up_taps(ho, H, h0, h1, a0, a1);
    up_taps(wo, W, w0, w1, b0, b1);
    const float* s = small + pl * H * W;
    const float up = a0 * (b0 * s[h0 * W + w0] + b1 * s[h0 * W + w1]) + a1 * (b0 * s[h1 * W + w0] + b1 * s[h1 * W + w1]);
    // torch.lerp(start, end, w) = start + w * (end - start)
    out[i] = up + alpha * (large[i] - up);
  }
}
// gsmall[h,w] = scale * sum over the hi-res pixels that read (h,w) of their tap weight * g
__global__ void img_up2_bwd_kernel(const float* __restrict__ g, float* __restrict__ gsmall, int P, int H, int W,
                                   float scale) {
  pdl_prologue();
  const int Ho = 2 * H, Wo = 2 * W;
  const size_t total = (size_t)P * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const size_t pl = i / ((size_t)W * H);
    const float* gp = g + pl * Ho * Wo;
    float acc = 0.f;
    for (int r = 2 * h - 1; r <= 2 * h + 2; ++r) {
      if (r < 0 || r >= Ho) continue;
      int i0, i1;
      float a0, a1;
      up_taps(r, H, i0, i1, a0, a1);
      const float wr = (i0 == h ? a0 : 0.f) + (i1 == h ? a1 : 0.f);
      if (wr == 0.f) continue;
      for (int s = 2 * w - 1; s <= 2 * w + 2; ++s) {
        if (s < 0 || s >= Wo) continue;
        int j0, j1;
        float b0, b1;
        up_taps(s, W, j0, j1, b0, b1);
        const float ws = (j0 == w ? b0 : 0.f) + (j1 == w ? b1 : 0.f);
        if (ws != 0.f) acc += wr * ws * gp[r * Wo + s];
      }
    }
    gsmall[i] = scale * acc;
  }
}

// sums[j] = sum over n, hw of g[n,j,hw]   (toRGB bias gradient);   planes laid out (B,3,HW)
__global__ void plane_sums_kernel(const float* __restrict__ g, float* __restrict__ sums, int B, int HW) {
  pdl_prologue();
  __shared__ float red[32];
  const int j = blockIdx.y;
  float s = 0.f;
  const size_t total = (size_t)B * HW;
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  if (HW % 4 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    // 16-byte loads (a group of four never straddles an image plane), four in flight per thread
    const size_t total4 = total / 4, hw4 = (size_t)HW / 4;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    auto at = [&](size_t i) { return g4[((i / hw4) * 3 + j) * hw4 + i % hw4]; };
    size_t i = tid;
    for (; i + 3 * nthr < total4; i += 4 * nthr) {
      const float4 a = at(i), b = at(i + nthr), c = at(i + 2 * nthr), d = at(i + 3 * nthr);
      s += (a.x + a.y + a.z + a.w) + (b.x + b.y + b.z + b.w) + (c.x + c.y + c.z + c.w) + (d.x + d.y + d.z + d.w);
    }
    for (; i < total4; i += nthr) {
      const float4 a = at(i);
      s += a.x + a.y + a.z + a.w;
    }
  } else {
    for (size_t i = tid; i < total; i += nthr) {
      const size_t n = i / HW, hw = i % HW;
      s += g[(n * 3 + j) * HW + hw];
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(sums + j, s);
  }
}

// ---------------------------------------------------------------------------------------------
// NHWC bf16 <-> NCHW fp32 for the critic head (nn.Flatten of a (B,512,1,1)/(B,512,4,4) map, gan.py:245-247)
// ---------------------------------------------------------------------------------------------
__global__ void nhwc_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int N, int HW,
                                        int C) {
  pdl_prologue();
  const size_t total = (size_t)N * HW * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int hw = (int)((i / C) % HW);
    const size_t n = i / ((size_t)C * HW);
    out[(n * C + c) * HW + hw] = __bfloat162float(x[i]);
  }
}
// out[n,hw,c] = g[n,c,hw] * gate(gate_src[n,hw,c])
__global__ void nchw_f32_to_nhwc_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ gate_src,
                                        __nv_bfloat16* __restrict__ out, int N, int HW, int C, float slope) {
  pdl_prologue();
  const size_t total = (size_t)N * HW * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int hw = (int)((i / C) % HW);
    const size_t n = i / ((size_t)C * HW);
    float v = g[(n * C + c) * HW + hw];
    if (gate_src != nullptr) v *= gate_f(__bfloat162float(gate_src[i]), slope);
    out[i] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------------------------------------
// y = act(coef * x W^T + b) for the two skinny shapes of the critic's 4x4 -> 1x1 convolution (K = 8192 forward, N = 8192
// input gradient), where one warp per output column re-reads the same 8 rows of x once per column (0.5 GB of L2 traffic
// for a 17 MB weight).  Here a block owns kLinNT columns x kLinMT rows; its KS warps split K, every x vector loaded is
// used for kLinNT columns, and the partial sums meet in shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kLinNT = 4;

template <int KS>
__global__ void __launch_bounds__(32 * KS) linear_fwd_tile_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                                  const float* __restrict__ bias, float* __restrict__ y,
                                                                  int M, int N, int K, float coef, int act, float slope) {
  pdl_prologue();
  __shared__ float part[KS][kLinMT * kLinNT];
  const int lane = threadIdx.x & 31, kw = threadIdx.x >> 5;
  const int ntiles = (N + kLinNT - 1) / kLinNT;
  const int n0 = ((int)blockIdx.x % ntiles) * kLinNT;
  const int m0 = ((int)blockIdx.x / ntiles) * kLinMT;
  float acc[kLinMT][kLinNT];
#pragma unroll
  for (int i = 0; i < kLinMT; ++i)
#pragma unroll
    for (int c = 0; c < kLinNT; ++c) acc[i][c] = 0.f;
  const int kspan = ((K / 4 + KS - 1) / KS) * 4;          // K % 4 == 0 (launcher)
  const int k_lo = kw * kspan, k_hi = min(K, k_lo + kspan);
  for (int k = k_lo + lane * 4; k < k_hi; k += 128) {
    float4 wv[kLinNT], xv[kLinMT];
#pragma unroll
    for (int c = 0; c < kLinNT; ++c)
      wv[c] = *reinterpret_cast<const float4*>(W + (size_t)min(n0 + c, N - 1) * K + k);
#pragma unroll
    for (int i = 0; i < kLinMT; ++i)
      xv[i] = *reinterpret_cast<const float4*>(x + (size_t)min(m0 + i, M - 1) * K + k);
#pragma unroll
    for (int i = 0; i < kLinMT; ++i)
#pragma unroll
      for (int c = 0; c < kLinNT; ++c)
        acc[i][c] += xv[i].x * wv[c].x + xv[i].y * wv[c].y + xv[i].z * wv[c].z + xv[i].w * wv[c].w;
  }
#pragma unroll
  for (int i = 0; i < kLinMT; ++i)
#pragma unroll
    for (int c = 0; c < kLinNT; ++c) {
      const float v = warp_sum(acc[i][c]);
      if (lane == 0) part[kw][i * kLinNT + c] = v;
    }
  __syncthreads();
  if (threadIdx.x < kLinMT * kLinNT) {
    const int i = threadIdx.x / kLinNT, c = threadIdx.x % kLinNT;
    if (m0 + i < M && n0 + c < N) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < KS; ++w) v += part[w][threadIdx.x];
      v = v * coef + (bias ? bias[n0 + c] : 0.f);
      if (act) v = lrelu_f(v, slope);
      y[(size_t)(m0 + i) * N + n0 + c] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// MiniBatchStdDev (gan.py:273-298).  x: (B,HW,C) bf16, J = HW*C positions, G groups, M = B/G slots.
//   mu[j] = mean_n x[n][j];  d = x - mu;  var_m[j] = mean_g d[g*M+m][j]^2;  sig = sqrt(var + eps)
//   s[m] = mean_j sig_m[j];  sample n gets plane value s[n mod M].
// One thread per position j; each block reduces its partial sums per slot and adds them atomically.
// Dynamic smem: M floats.
// ---------------------------------------------------------------------------------------------
// mode 0: s[m]    += sum_j sig_m[j] / J
// mode 1: sdot[m] += sum_j (sum_g d_g * ddot_g) / (G * sig_m[j]) / J          (tangent; v = x-dot)
__global__ void mbstd_reduce_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ v,
                                    float* __restrict__ out, int B, int G, int J, float eps, int mode) {
  pdl_prologue();
  extern __shared__ float part[];
  const int M = B / G;
  for (int i = threadIdx.x; i < M; i += blockDim.x) part[i] = 0.f;
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = j < J;
  const int jj = active ? j : J - 1;                      // idle lanes still take part in the warp sums
  float mu = 0.f, mud = 0.f;
#pragma unroll 8
  for (int n = 0; n < B; ++n) {                            // batch loads are independent: keep 8 in flight
    mu += __bfloat162float(x[(size_t)n * J + jj]);
    if (mode == 1) mud += __bfloat162float(v[(size_t)n * J + jj]);
  }
  mu /= B;
  mud /= B;
  for (int m = 0; m < M; ++m) {
    float sq = 0.f, dd = 0.f;
#pragma unroll 4
    for (int g = 0; g < G; ++g) {
      const size_t off = (size_t)(g * M + m) * J + jj;
      const float d = __bfloat162float(x[off]) - mu;
      sq += d * d;
      if (mode == 1) dd += d * (__bfloat162float(v[off]) - mud);
    }
    const float sig = sqrtf(sq / G + eps);
    float val = mode == 0 ? sig : dd / (G * sig);
    val = warp_sum(active ? val : 0.f);
    if ((threadIdx.x & 31) == 0) atomicAdd(&part[m], val);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < M; i += blockDim.x) atomicAdd(out + i, part[i] / J);
}

// Deterministic variant (bg_set_deterministic): block m owns slot m alone; threads stride over the positions, the block
// total is an ordered two-stage sum (xor-butterfly inside a warp, warps in index order).
__global__ void mbstd_reduce_det_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ v,
                                        float* __restrict__ out, int B, int G, int J, float eps, int mode) {
  pdl_prologue();
  __shared__ float red[32];
  const int M = B / G;
  const int m = blockIdx.x;
  float acc = 0.f;
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    float mu = 0.f, mud = 0.f;
    for (int n = 0; n < B; ++n) {
      mu += __bfloat162float(x[(size_t)n * J + j]);
      if (mode == 1) mud += __bfloat162float(v[(size_t)n * J + j]);
    }
    mu /= B;
    mud /= B;
    float sq = 0.f, dd = 0.f;
    for (int g = 0; g < G; ++g) {
      const size_t off = (size_t)(g * M + m) * J + j;
      const float d = __bfloat162float(x[off]) - mu;
      sq += d * d;
      if (mode == 1) dd += d * (__bfloat162float(v[off]) - mud);
    }
    const float sig = sqrtf(sq / G + eps);
    acc += mode == 0 ? sig : dd / (G * sig);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    out[m] = t / J;
  }
}

// xpad[n][hw][0..C) = x;  xpad[n][hw][C] = plane[n mod M];  xpad[n][hw][C+1..Cpad) = 0
__global__ void mbstd_pad_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ plane,
                                 __nv_bfloat16* __restrict__ xpad, int B, int HW, int C, int Cpad, int M) {
  pdl_prologue();
  const size_t total = (size_t)B * HW * Cpad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    const size_t p = i / Cpad;
    const int n = (int)(p / HW);
    __nv_bfloat16 v;
    if (c < C) v = x[p * C + c];
    else if (c == C) v = __float2bfloat16_rn(plane[n % M]);
    else v = __float2bfloat16_rn(0.f);
    xpad[i] = v;
  }
}

// gs[m] = sum over n == m (mod M), hw of gpad[n][hw][C]      (gradient reaching the stddev plane)
__global__ void mbstd_plane_grad_kernel(const __nv_bfloat16* __restrict__ gpad, float* __restrict__ gs, int B, int HW,
                                        int C, int Cpad, int M) {
  pdl_prologue();
  const int m = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < (B / M) * HW; i += blockDim.x) {
    const int g = i / HW, hw = i % HW;
    s += __bfloat162float(gpad[((size_t)(g * M + m) * HW + hw) * Cpad + C]);
  }
  s = warp_sum(s);
  __shared__ float red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    gs[m] = t;
  }
}

// gx[n][j] = gpad[n][j(:C)] + (1/(J*G)) * ( r[n][j] - mean_n' r[n'][j] ),
//   r[n][j] = gs[m(n)] * d[n][j] / sig_m[j]                                       (first-order VJP)
//           + gs2[m(n)] * ( ddot[n][j]/sig - A_m[j] * d[n][j] / (G * sig^3) )     (second-order term, optional)
//   A_m[j] = sum_g d_g * ddot_g.   gs2/v null -> first-order only.  gpad null -> no pass-through term.
// Register-resident form for the usual batches (B <= 32): every x / v / plane-gradient value of position j is loaded ONCE,
// all loads in flight together (the generic kernel below walks the batch four times with 8 loads in flight: 64 blocks
// of dependent L1 round trips, 29 us for 1 MB of input).  Same arithmetic in the same order as the generic kernel.
template <int kB, int kG, bool kV>
__global__ void __launch_bounds__(128)
mbstd_bwd_reg_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ v,
                     const __nv_bfloat16* __restrict__ gpad, const float* __restrict__ gs, const float* __restrict__ gs2,
                     __nv_bfloat16* __restrict__ gx, int HW, int C, int Cpad, float eps) {
  pdl_prologue();
  constexpr int kM = kB / kG;
  const int J = HW * C;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= J) return;
  const int c = j % C, hw = j / C;
  float xv[kB], vv[kV ? kB : 1], gp[kB];
#pragma unroll
  for (int n = 0; n < kB; ++n) {
    xv[n] = __bfloat162float(x[(size_t)n * J + j]);
    if (kV) vv[n] = __bfloat162float(v[(size_t)n * J + j]);
    gp[n] = gpad != nullptr ? __bfloat162float(gpad[((size_t)n * HW + hw) * Cpad + c]) : 0.f;
  }
  float mu = 0.f, mud = 0.f;
#pragma unroll
  for (int n = 0; n < kB; ++n) {
    mu += xv[n];
    if (kV) mud += vv[n];
  }
  mu /= kB;
  mud /= kB;
  float sig[kM], Am[kM], gsm[kM], gs2m[kM];
  float rmean = 0.f;
#pragma unroll
  for (int m = 0; m < kM; ++m) {
    float sq = 0.f, A = 0.f, sd = 0.f, sdd = 0.f;
#pragma unroll
    for (int g = 0; g < kG; ++g) {
      const float d = xv[g * kM + m] - mu;
      sq += d * d;
      sd += d;
      if (kV) {
        const float dd = vv[g * kM + m] - mud;
        A += d * dd;
        sdd += dd;
      }
    }
    sig[m] = sqrtf(sq / kG + eps);
    Am[m] = A;
    gsm[m] = gs ? gs[m] : 0.f;
    gs2m[m] = kV ? gs2[m] : 0.f;
    float r = gsm[m] * sd / sig[m];
    if (kV) r += gs2m[m] * (sdd / sig[m] - A * sd / (kG * sig[m] * sig[m] * sig[m]));
    rmean += r;
  }
  rmean /= kB;
  const float scale = 1.f / ((float)J * kG);
#pragma unroll
  for (int m = 0; m < kM; ++m) {
#pragma unroll
    for (int g = 0; g < kG; ++g) {
      const int n = g * kM + m;
      const float d = xv[n] - mu;
      float r = gsm[m] * d / sig[m];
      if (kV) {
        const float dd = vv[n] - mud;
        r += gs2m[m] * (dd / sig[m] - Am[m] * d / (kG * sig[m] * sig[m] * sig[m]));
      }
      float o = scale * (r - rmean);
      if (gpad != nullptr) o += gp[n];
      gx[(size_t)n * J + j] = __float2bfloat16_rn(o);
    }
  }
}

__global__ void mbstd_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ v,
                                 const __nv_bfloat16* __restrict__ gpad, const float* __restrict__ gs,
                                 const float* __restrict__ gs2, __nv_bfloat16* __restrict__ gx, int B, int G, int HW,
                                 int C, int Cpad, float eps) {
  pdl_prologue();
  const int J = HW * C;
  const int M = B / G;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= J) return;
  const int c = j % C, hw = j / C;
  float mu = 0.f, mud = 0.f;
#pragma unroll 8
  for (int n = 0; n < B; ++n) {
    mu += __bfloat162float(x[(size_t)n * J + j]);
    if (v != nullptr) mud += __bfloat162float(v[(size_t)n * J + j]);
  }
  mu /= B;
  mud /= B;
  // pass 1: mean over n of r[n][j]
  float rmean = 0.f;
  for (int m = 0; m < M; ++m) {
    float sq = 0.f, A = 0.f, sd = 0.f, sdd = 0.f;
#pragma unroll 4
    for (int g = 0; g < G; ++g) {
      const size_t off = (size_t)(g * M + m) * J + j;
      const float d = __bfloat162float(x[off]) - mu;
      sq += d * d;
      sd += d;
      if (v != nullptr) {
        const float dd = __bfloat162float(v[off]) - mud;
        A += d * dd;
        sdd += dd;
      }
    }
    const float sig = sqrtf(sq / G + eps);
    float r = (gs ? gs[m] : 0.f) * sd / sig;
    if (v != nullptr) r += gs2[m] * (sdd / sig - A * sd / (G * sig * sig * sig));
    rmean += r;
  }
  rmean /= B;
  // pass 2: per-sample values
  const float scale = 1.f / ((float)J * G);
  for (int m = 0; m < M; ++m) {
    float sq = 0.f, A = 0.f;
#pragma unroll 4
    for (int g = 0; g < G; ++g) {
      const size_t off = (size_t)(g * M + m) * J + j;
      const float d = __bfloat162float(x[off]) - mu;
      sq += d * d;
      if (v != nullptr) A += d * (__bfloat162float(v[off]) - mud);
    }
    const float sig = sqrtf(sq / G + eps);
    for (int g = 0; g < G; ++g) {
      const int n = g * M + m;
      const size_t off = (size_t)n * J + j;
      const float d = __bfloat162float(x[off]) - mu;
      float r = (gs ? gs[m] : 0.f) * d / sig;
      if (v != nullptr) {
        const float dd = __bfloat162float(v[off]) - mud;
        r += gs2[m] * (dd / sig - A * d / (G * sig * sig * sig));
      }
      float o = scale * (r - rmean);
      if (gpad != nullptr) o += __bfloat162float(gpad[((size_t)n * HW + hw) * Cpad + c]);
      gx[off] = __float2bfloat16_rn(o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// loss terms (gan.py:228, 396, 406):  loss = mean softplus(sign * pred);  seed = d loss / d pred * seed_scale
// single block; n <= a few hundred
// ---------------------------------------------------------------------------------------------
__global__ void logistic_loss_kernel(const float* __restrict__ pred, int n, float sign, float* __restrict__ loss,
                                     float* __restrict__ seed, float seed_scale) {
  pdl_prologue();
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float t = sign * pred[i];
    // softplus with torch's threshold (beta=1, threshold=20)
    s += t > 20.f ? t : log1pf(expf(t));
    if (seed != nullptr) seed[i] = seed_scale * sign / (1.f + expf(-t)) / n;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    loss[0] = t / n;
  }
}

// out[0] += scale * sum x^2
__global__ void sumsq_kernel(const float* __restrict__ x, size_t n, float scale, float* __restrict__ out) {
  pdl_prologue();
  __shared__ float red[32];
  float s = 0.f;
  // a pure stream (25 MB for the R1 penalty at 256x256): 16-byte loads, four in flight per thread
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const size_t n4 = n / 4;
    size_t i = tid;
    for (; i + 3 * nthr < n4; i += 4 * nthr) {
      const float4 a = x4[i], b = x4[i + nthr], c = x4[i + 2 * nthr], d = x4[i + 3 * nthr];
      s += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w +
           c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w + d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
    }
    for (; i < n4; i += nthr) {
      const float4 a = x4[i];
      s += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    for (size_t k = n4 * 4 + tid; k < n; k += nthr) s += x[k] * x[k];
  } else {
    for (size_t i = tid; i < n; i += nthr) s += x[i] * x[i];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, s * scale);
  }
}

inline int grid1d(size_t work, int block = 256, int cap_mult = 8) {
  size_t blocks = (work + block - 1) / block;
  const size_t cap = (size_t)num_sms() * cap_mult;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
int launch_linear_fwd(const float* x, const float* W, const float* bias, float* y, int M, int N, int K, float coef,
                      int act, float slope, cudaStream_t s) {
  BG_REQUIRE(M > 0 && N > 0 && K > 0, "linear_fwd: bad shape M %d N %d K %d", M, N, K);
  const int mgroups = (M + kLinMT - 1) / kLinMT;
  if (K % 4 == 0 && (K >= 512 || N >= 4096) && N >= 64 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
    const unsigned tiles = (unsigned)(((N + kLinNT - 1) / kLinNT) * mgroups);
    if (K >= 4096)
      BG_CHECK_CUDA(launch_pdl(linear_fwd_tile_kernel<8>, tiles, 256, 0, s, x, W, bias, y, M, N, K, coef, act, slope));
    else if (K >= 512 && N < 4096)   // e.g. the mapping network's 512 x 512 layers: one 128-float slice of K per warp
      BG_CHECK_CUDA(launch_pdl(linear_fwd_tile_kernel<4>, tiles, 128, 0, s, x, W, bias, y, M, N, K, coef, act, slope));
    else
      BG_CHECK_CUDA(launch_pdl(linear_fwd_tile_kernel<1>, tiles, 32, 0, s, x, W, bias, y, M, N, K, coef, act, slope));
    return 0;
  }
  const long warps = (long)N * mgroups;
  const int block = 256;
  const long blocks = (warps * 32 + block - 1) / block;
  BG_CHECK_CUDA(launch_pdl(linear_fwd_kernel, (unsigned)blocks, block, 0, s, x, W, bias, y, M, N, K, coef, act, slope));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Input gradient of one or several EqualizedLinear layers that share their input (gan.py:16-17; the AdaIN style FCs,
// gan.py:60,66), straight from the (N_g, K) row-major weights:  gx[m][k] = sum_g coef_g * sum_n gy_g[m][n] * W_g[n][k].
// No transposed weight copies: rows of W are read as they lie (coalesced along k, two k per lane).  A block owns
// (64 k) x (8 rows of gy); the concatenated n range of all groups is split over the `zsplit` CTAs of a cluster, inside a
// CTA over its 8 warps (gy staged, pre-scaled, in shared memory).  Warp partials meet in shared memory, CTA partials in
// the leader's registers through DSMEM in rank order: no atomics, the result does not depend on timing.
// ---------------------------------------------------------------------------------------------
constexpr int kLbiMT = 8;        // rows of gy per block
constexpr int kLbiNC = 512;      // n per staged chunk of gy
constexpr int kLbiMaxZ = 8;      // portable cluster size
__device__ __forceinline__ float2 ld_dsmem_f2(uint32_t cluster_addr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(cluster_addr));
  return v;
}
__global__ void __launch_bounds__(256)
linear_bwd_input_kernel(const LinGroups G, float* __restrict__ gx, int M, int K, int total_n, int zsplit) {
  pdl_prologue();
  __shared__ __align__(16) float gs[kLbiNC][kLbiMT];          // gy chunk, transposed: one n = two 16-byte broadcasts
  __shared__ float2 part[8][kLbiMT][32];
  __shared__ float2 cta_sum[kLbiMT][32];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int z = (int)blockIdx.x % zsplit;                     // rank in the cluster (cluster dims (zsplit, 1, 1))
  const int k = (((int)blockIdx.x / zsplit) * 32 + lane) * 2; // two consecutive k per lane: 256-byte rows per warp
  const int m0 = (int)blockIdx.y * kLbiMT;
  const bool live = k < K;                                    // K % 2 == 0 (launcher)
  float2 acc[kLbiMT];
#pragma unroll
  for (int i = 0; i < kLbiMT; ++i) acc[i] = make_float2(0.f, 0.f);
  const int share = (total_n + zsplit - 1) / zsplit;
  const int z_lo = z * share, z_hi = min(total_n, z_lo + share);
  int off = 0;
  for (int g = 0; g < G.groups; ++g) {
    const int N = G.N[g];
    const int a = max(z_lo, off) - off, b = min(z_hi, off + N) - off;   // this CTA's rows [a, b) of group g
    off += N;
    if (a >= b) continue;
    const float coef = G.coef[g];
    const float* __restrict__ gy = G.y[g];
    const float* __restrict__ W = G.W[g];
    for (int nc0 = a; nc0 < b; nc0 += kLbiNC) {
      const int cn = min(kLbiNC, b - nc0);
      __syncthreads();
      {
        // warp w stages row m0 + w; 16 independent loads in flight per lane
        const bool row_live = m0 + wp < M;
        const float* grow = gy + (size_t)min(m0 + wp, M - 1) * N + nc0;
#pragma unroll 16
        for (int n = lane; n < cn; n += 32) gs[n][wp] = row_live ? grow[n] * coef : 0.f;
      }
      __syncthreads();
      const int per = (cn + 7) >> 3;
      const int lo = wp * per, hi = min(cn, lo + per);
      const float* wp0 = W + (size_t)nc0 * K + (live ? k : 0);
#pragma unroll 16
      for (int n = lo; n < hi; ++n) {                         // 16 independent weight rows in flight per lane
        const float2 w2 = *reinterpret_cast<const float2*>(wp0 + (size_t)n * K);
        const float4 ga = *reinterpret_cast<const float4*>(&gs[n][0]);
        const float4 gb = *reinterpret_cast<const float4*>(&gs[n][4]);
        const float gv[kLbiMT] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
        for (int i = 0; i < kLbiMT; ++i) {
          acc[i].x = fmaf(gv[i], w2.x, acc[i].x);
          acc[i].y = fmaf(gv[i], w2.y, acc[i].y);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kLbiMT; ++i) part[wp][i][lane] = acc[i];
  __syncthreads();
  float2 v = make_float2(0.f, 0.f);                           // 8 warps x 8 rows: warp w finishes row w, fixed order
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const float2 pv = part[w][wp][lane];
    v.x += pv.x;
    v.y += pv.y;
  }
  if (zsplit > 1) {
    cta_sum[wp][lane] = v;
    cluster_sync_all();                                       // every CTA's partial is in its shared memory
    if (z == 0) {
      const uint32_t local = smem_u32(&cta_sum[wp][lane]);
      for (int r = 1; r < zsplit; ++r) {
        const float2 pv = ld_dsmem_f2(mapa_cta(local, (uint32_t)r));
        v.x += pv.x;
        v.y += pv.y;
      }
    }
    cluster_sync_all();                                       // the leader is done reading its peers
    if (z != 0) return;
  }
  if (live && m0 + wp < M) *reinterpret_cast<float2*>(gx + (size_t)(m0 + wp) * K + k) = v;
}

static int launch_linear_bwd_input_groups(const LinGroups& G, float* gx, int M, int K, cudaStream_t s) {
  BG_REQUIRE(M > 0 && K > 0 && K % 2 == 0, "linear_bwd_input: bad shape M %d K %d", M, K);
  BG_REQUIRE((reinterpret_cast<uintptr_t>(gx) & 7) == 0, "linear_bwd_input: gx must be 8-byte aligned");
  int total_n = 0;
  for (int g = 0; g < G.groups; ++g) {
    BG_REQUIRE(G.N[g] > 0 && (reinterpret_cast<uintptr_t>(G.W[g]) & 7) == 0, "linear_bwd_input: bad group %d", g);
    total_n += G.N[g];
  }
  const int ktiles = (K / 2 + 31) / 32, mtiles = (M + kLbiMT - 1) / kLbiMT;
  // split n over a cluster while the grid is short of two waves and every CTA keeps >= 32 rows of W
  int zsplit = 1;
  while (zsplit < kLbiMaxZ && ktiles * mtiles * zsplit < 2 * num_sms() && total_n / (2 * zsplit) >= 32) zsplit *= 2;
  dim3 grid(ktiles * zsplit, mtiles);
  BG_CHECK_CUDA(launch_pdl_cluster(linear_bwd_input_kernel, grid, dim3(256), 0, s, zsplit, G, gx, M, K, total_n, zsplit));
  return 0;
}

int launch_linear_bwd_input(const float* gy, const float* W, float* gx, int M, int N, int K, float coef, cudaStream_t s) {
  LinGroups G;
  memset(&G, 0, sizeof(G));
  G.groups = 1;
  G.W[0] = W;
  G.y[0] = const_cast<float*>(gy);
  G.N[0] = N;
  G.coef[0] = coef;
  return launch_linear_bwd_input_groups(G, gx, M, K, s);
}

int launch_linear_bwd_weight(const float* gy, const float* x, float* dW, float* db, int M, int N, int K, float coef,
                             int accumulate, cudaStream_t s) {
  BG_REQUIRE(M > 0 && N > 0 && K > 0 && K % 4 == 0, "linear_bwd_weight: bad shape M %d N %d K %d", M, N, K);
  dim3 grid((K / 4 + 127) / 128, (N + kLbwNT - 1) / kLbwNT);
  BG_CHECK_CUDA(launch_pdl(linear_bwd_weight_kernel, grid, 128, 0, s, gy, x, dW, db, M, N, K, coef, accumulate));
  return 0;
}

int launch_transpose_f32(const float* in, float* out, int R, int C, cudaStream_t s) {
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  BG_CHECK_CUDA(launch_pdl(transpose_kernel, grid, block, 0, s, in, out, R, C));
  return 0;
}

int launch_act_gate_f32(const float* g, const float* y, float* out, size_t n, float slope, cudaStream_t s) {
  BG_CHECK_CUDA(launch_pdl(act_gate_f32_kernel, grid1d(n), 256, 0, s, g, y, out, n, slope));
  return 0;
}

int launch_axpby_f32(const float* a, const float* b, float* out, size_t n, float ca, float cb, cudaStream_t s) {
  BG_CHECK_CUDA(launch_pdl(axpby_f32_kernel, grid1d(n), 256, 0, s, a, b, out, n, ca, cb));
  return 0;
}

int launch_const_noise_act(const float* cst, const float* noise, const float* nw, void* a, int N, int HW, int C,
                           float slope, cudaStream_t s) {
  BG_CHECK_CUDA(launch_pdl(const_noise_act_kernel, grid1d((size_t)N * HW * C), 256, 0, s, cst, noise, nw,
                           (__nv_bfloat16*)a, N, HW, C, slope));
  return 0;
}

int launch_const_bwd(const void* g, float* dconst, int N, int HW, int C, cudaStream_t s) {
  BG_CHECK_CUDA(launch_pdl(const_bwd_kernel, (HW * C + 255) / 256, 256, 0, s, (const __nv_bfloat16*)g, dconst, N, HW,
                           C));
  return 0;
}

int launch_img_avgpool2(const float* img, float* out, int P, int Ho, int Wo, cudaStream_t s) {
  BG_CHECK_CUDA(launch_pdl(img_avgpool2_kernel, grid1d((size_t)P * Ho * Wo), 256, 0, s, img, out, P, Ho, Wo));
  return 0;
}

int launch_img_avgpool2_bwd(const float* g, float* gimg, int P, int Ho, int Wo, float scale, int accumulate,
                            cudaStream_t s) {
  BG_CHECK_CUDA(launch_pdl(img_avgpool2_bwd_kernel, grid1d((size_t)P * Ho * Wo * 4), 256, 0, s, g, gimg, P, Ho, Wo,
                           scale, accumulate));
  return 0;
}

int launch_img_up2_lerp(const float* small, const float* large, float* out, int P, int H, int W, float alpha,
                        cudaStream_t s) {
  BG_CHECK_CUDA(launch_pdl(img_up2_lerp_kernel, grid1d((size_t)P * H * W * 4), 256, 0, s, small, large, out, P, H, W,
                           alpha));
  return 0;
}

int launch_img_up2_bwd(const float* g, float* gsmall, int P, int H, int W, float scale, cudaStream_t s) {
  BG_CHECK_CUDA(launch_pdl(img_up2_bwd_kernel, grid1d((size_t)P * H * W), 256, 0, s, g, gsmall, P, H, W, scale));
  return 0;
}

int launch_plane_sums(const float* g, float* sums, int B, int HW, cudaStream_t s) {
  if (launch_zero(sums, 3 * sizeof(float), s) != 0) return 1;
  int bx = grid1d((size_t)B * HW, 256, 2);
  BG_CHECK_CUDA(launch_pdl(plane_sums_kernel, dim3(bx, 3), 256, 0, s, g, sums, B, HW));
  return 0;
}

int launch_nhwc_to_nchw_f32(const void* x, float* out, int N, int HW, int C, cudaStream_t s) {
  BG_CHECK_CUDA(launch_pdl(nhwc_to_nchw_f32_kernel, grid1d((size_t)N * HW * C), 256, 0, s, (const __nv_bfloat16*)x, out,
                           N, HW, C));
  return 0;
}

int launch_nchw_f32_to_nhwc(const float* g, const void* gate_src, void* out, int N, int HW, int C, float slope,
                            cudaStream_t s) {
  BG_CHECK_CUDA(launch_pdl(nchw_f32_to_nhwc_kernel, grid1d((size_t)N * HW * C), 256, 0, s, g,
                           (const __nv_bfloat16*)gate_src, (__nv_bfloat16*)out, N, HW, C, slope));
  return 0;
}

int launch_mbstd_fwd(const void* x, const void* v, float* plane, void* xpad, int B, int G, int HW, int C, int Cpad,
                     float eps, cudaStream_t s) {
  BG_REQUIRE(G > 0 && B % G == 0, "mbstd: batch %d is not a multiple of the group size %d", B, G);
  BG_REQUIRE(Cpad > C, "mbstd: Cpad %d must exceed C %d", Cpad, C);
  const int M = B / G, J = HW * C;
  const void* src = v ? v : x;
  if (deterministic()) {
    BG_CHECK_CUDA(launch_pdl(mbstd_reduce_det_kernel, M, 256, 0, s, (const __nv_bfloat16*)x, (const __nv_bfloat16*)v, plane,
                             B, G, J, eps, v ? 1 : 0));
  } else {
    if (launch_zero(plane, M * sizeof(float), s) != 0) return 1;
    BG_CHECK_CUDA(launch_pdl(mbstd_reduce_kernel, (J + 127) / 128, 128, M * sizeof(float), s, (const __nv_bfloat16*)x,
                             (const __nv_bfloat16*)v, plane, B, G, J, eps, v ? 1 : 0));
  }
  if (xpad != nullptr) {
    BG_CHECK_CUDA(launch_pdl(mbstd_pad_kernel, grid1d((size_t)B * HW * Cpad), 256, 0, s, (const __nv_bfloat16*)src,
                             plane, (__nv_bfloat16*)xpad, B, HW, C, Cpad, M));
  }
  return 0;
}

int launch_mbstd_bwd(const void* x, const void* v, const void* gpad, const void* gpad2, float* gs_ws, void* gx, int B,
                     int G, int HW, int C, int Cpad, float eps, cudaStream_t s) {
  BG_REQUIRE(G > 0 && B % G == 0, "mbstd_bwd: batch %d is not a multiple of the group size %d", B, G);
  BG_REQUIRE((v == nullptr) == (gpad2 == nullptr), "mbstd_bwd: tangent v and its plane gradient come together");
  const int M = B / G, J = HW * C;
  float* gs = nullptr;
  float* gs2 = nullptr;
  if (gpad != nullptr) {
    gs = gs_ws;
    BG_CHECK_CUDA(launch_pdl(mbstd_plane_grad_kernel, M, 128, 0, s, (const __nv_bfloat16*)gpad, gs, B, HW, C, Cpad, M));
  }
  if (gpad2 != nullptr) {
    gs2 = gs_ws + M;
    BG_CHECK_CUDA(launch_pdl(mbstd_plane_grad_kernel, M, 128, 0, s, (const __nv_bfloat16*)gpad2, gs2, B, HW, C, Cpad,
                             M));
  }
  const __nv_bfloat16 *xb = (const __nv_bfloat16*)x, *vb = (const __nv_bfloat16*)v, *gb = (const __nv_bfloat16*)gpad;
  __nv_bfloat16* ob = (__nv_bfloat16*)gx;
  const unsigned blocks = (unsigned)((J + 127) / 128);
#define BG_MBSTD_REG(kB, kG)                                                                                              \
  if (B == kB && G == kG) {                                                                                               \
    if (v != nullptr)                                                                                                     \
      BG_CHECK_CUDA(launch_pdl(mbstd_bwd_reg_kernel<kB, kG, true>, blocks, 128, 0, s, xb, vb, gb, gs, gs2, ob, HW, C, Cpad, eps)); \
    else                                                                                                                  \
      BG_CHECK_CUDA(launch_pdl(mbstd_bwd_reg_kernel<kB, kG, false>, blocks, 128, 0, s, xb, vb, gb, gs, gs2, ob, HW, C, Cpad, eps)); \
    return 0;                                                                                                             \
  }
  BG_MBSTD_REG(32, 4)
  BG_MBSTD_REG(16, 4)
  BG_MBSTD_REG(8, 4)
#undef BG_MBSTD_REG
  BG_CHECK_CUDA(launch_pdl(mbstd_bwd_kernel, blocks, 128, 0, s, xb, vb, gb, gs, gs2, ob, B, G, HW, C, Cpad, eps));
  return 0;
}

int launch_logistic_loss(const float* pred, int n, float sign, float* loss, float* seed, float seed_scale,
                         cudaStream_t s) {
  BG_REQUIRE(n > 0, "logistic_loss: empty prediction vector");
  BG_CHECK_CUDA(launch_pdl(logistic_loss_kernel, 1, 256, 0, s, pred, n, sign, loss, seed, seed_scale));
  return 0;
}

// WGAN-GP rows (gan.py:385): block n owns sample n.  r = ||g_n||_2 (ordered two-stage sum: the result feeds the tangent pass);
// pen[0] += pen_scale * (r - 1)^2;  v[n] = v_scale * 2 (r - 1) / r * g_n  (d penalty / d g_n).
__global__ void gp_rows_kernel(const float* __restrict__ g, size_t D, float pen_scale, float v_scale,
                               float* __restrict__ pen, float* __restrict__ v) {
  pdl_prologue();
  __shared__ float red[32];
  __shared__ float coef_s;
  const float* gn = g + (size_t)blockIdx.x * D;
  float* vn = v + (size_t)blockIdx.x * D;
  float s = 0.f;
  for (size_t i = threadIdx.x; i < D; i += blockDim.x) s += gn[i] * gn[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    const float r = sqrtf(t);
    atomicAdd(pen, pen_scale * (r - 1.f) * (r - 1.f));
    coef_s = v_scale * 2.f * (r - 1.f) / fmaxf(r, 1e-30f);
  }
  __syncthreads();
  const float coef = coef_s;
  for (size_t i = threadIdx.x; i < D; i += blockDim.x) vn[i] = coef * gn[i];
}

int launch_gp_rows(const float* g, int B, size_t D, float pen_scale, float v_scale, float* pen, float* v,
                   cudaStream_t s) {
  BG_REQUIRE(B > 0 && D > 0, "gp_rows: empty input");
  if (launch_zero(pen, sizeof(float), s) != 0) return 1;
  BG_CHECK_CUDA(launch_pdl(gp_rows_kernel, B, 1024, 0, s, g, D, pen_scale, v_scale, pen, v));
  return 0;
}

// Data feed (train.py:43-50 on the device): uint8 (B,H,W,3) -> fp32 (B,3,H,W), x / 127.5 - 1 (ToTensor + Normalize(.5,.5)),
// sample n mirrored along W when flip[n] != 0 (RandomHorizontalFlip).  One thread per output pixel: the three colour
// planes are written coalesced, the 3-byte source pixels of a warp are 96 contiguous bytes.
__global__ void image_feed_u8_kernel(const uint8_t* __restrict__ src, const uint8_t* __restrict__ flip,
                                     float* __restrict__ out, int B, int H, int W) {
  pdl_prologue();
  const size_t HW = (size_t)H * W, total = (size_t)B * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / HW, hw = i % HW;
    const int h = (int)(hw / W), w = (int)(hw % W);
    const int ws = (flip != nullptr && flip[n] != 0) ? W - 1 - w : w;
    const uint8_t* px = src + ((n * H + h) * (size_t)W + ws) * 3;
    float* o = out + n * 3 * HW + hw;
    o[0] = (float)px[0] * (1.f / 127.5f) - 1.f;
    o[HW] = (float)px[1] * (1.f / 127.5f) - 1.f;
    o[2 * HW] = (float)px[2] * (1.f / 127.5f) - 1.f;
  }
}

int launch_image_feed_u8(const void* src, const void* flip, float* out, int B, int H, int W, cudaStream_t s) {
  BG_REQUIRE(B > 0 && H > 0 && W > 0, "image_feed_u8: empty batch");
  BG_CHECK_CUDA(launch_pdl(image_feed_u8_kernel, grid1d((size_t)B * H * W, 256, 8), 256, 0, s, (const uint8_t*)src,
                           (const uint8_t*)flip, out, B, H, W));
  return 0;
}

int launch_sumsq(const float* x, size_t n, float scale, float* out, cudaStream_t s) {
  if (launch_zero(out, sizeof(float), s) != 0) return 1;
  BG_CHECK_CUDA(launch_pdl(sumsq_kernel, grid1d(n, 256, 2), 256, 0, s, x, n, scale, out));
  return 0;
}

// mode 0: forward, 1: weight gradient, 2: input gradient.  Pointer arrays are HOST arrays of device pointers.
int launch_linear_grouped(int mode, const float* x, const float* const* xs, const float* const* W,
                          const float* const* bias, float* const* y, float* const* dW, float* const* db, const int* N,
                          const float* coef, int groups, int M, int K, int act, float slope, float* gx, cudaStream_t s) {
  BG_REQUIRE(groups > 0 && groups <= kMaxLinGroups, "linear_grouped: 1..%d groups (got %d)", kMaxLinGroups, groups);
  BG_REQUIRE(M > 0 && K > 0 && K % 4 == 0, "linear_grouped: bad shape M %d K %d", M, K);
  BG_REQUIRE(mode != 1 || xs != nullptr, "linear_grouped: the weight gradient needs one input per layer");
  LinGroups G;
  memset(&G, 0, sizeof(G));
  G.groups = groups;
  const int mgroups = (M + kLinMT - 1) / kLinMT;
  int blocks = 0;
  for (int g = 0; g < groups; ++g) {
    BG_REQUIRE(N[g] > 0 && N[g] % 4 == 0, "linear_grouped: N[%d] = %d must be a positive multiple of 4", g, N[g]);
    G.W[g] = W[g];
    G.x[g] = xs ? xs[g] : nullptr;
    G.b[g] = bias ? bias[g] : nullptr;
    G.y[g] = y[g];
    G.dW[g] = dW ? dW[g] : nullptr;
    G.db[g] = db ? db[g] : nullptr;
    G.N[g] = N[g];
    G.coef[g] = coef[g];
    G.blk0[g] = blocks;
    if (mode == 0) blocks += (int)(((long)N[g] * mgroups * 32 + 255) / 256);
    else if (mode == 1) blocks += ((K / 4 + 127) / 128) * ((N[g] + kLbwNT - 1) / kLbwNT);
  }
  G.blk0[groups] = blocks;
  if (mode == 0) {
    BG_CHECK_CUDA(launch_pdl(linear_fwd_grouped_kernel, blocks, 256, 0, s, x, G, M, K, act, slope));
  } else if (mode == 1) {
    BG_CHECK_CUDA(launch_pdl(linear_bwd_weight_grouped_kernel, blocks, 128, 0, s, G, M, K));
  } else {
    BG_REQUIRE(gx != nullptr, "linear_grouped: gx is required for the input gradient");
    return launch_linear_bwd_input_groups(G, gx, M, K, s);
  }
  BG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace bg
