/*
 * bg_b200.h — C ABI of the B200-native StyleGAN hot path behind BYO-GAN's gan.py interface.
 *
 * The reference has no FFI layer: its hot path is torch library calls inside gan.py.  Each entry point
 * below replaces one such call site (cited as gan.py:line) with a hand-written sm_100a kernel.  The
 * Python host (byo-gan_b200/gan.py) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - the caller owns every buffer (including workspaces); nothing is allocated or freed here;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - feature maps are NHWC bf16 (C multiple of 16 for convolutions, 8 elsewhere);
 *     images at the module boundary are NCHW fp32 as the reference passes them;
 *   - return 0 on success, non-zero on error; bg_last_error() returns a thread-local message;
 *   - re-entrant: no global mutable state besides per-device attribute caches and the bg_set_deterministic switch;
 *   - kernels are launched with programmatic stream serialisation: the next kernel of the stream may become
 *     resident while this one is still running, and every kernel of the library waits (griddepcontrol.wait) for
 *     its predecessor's completion and memory visibility before its first global access, so results are exactly
 *     those of plain stream order.  Kernels of other libraries on the same stream are ordered as usual.
 *     BG_PDL=0 in the environment turns the attribute off.
 */
#ifndef BG_B200_H_
#define BG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* bg_last_error(void);
int bg_abi_version(void);
/* Chain-deterministic reductions (process-wide switch; default off, or BG_DETERMINISTIC=1 in the environment).
 * The reference's results are deterministic on one device (gan.py uses no atomics of its own); the fast path of this
 * library is not: instance-norm statistics (gan.py:59), the AdaIN backward sums and the minibatch-stddev plane
 * (gan.py:283-291) are combined across CTAs with fp32 atomics, and a last-bit difference there is amplified by the bf16
 * roundings and LeakyReLU gates of the ~50 layers behind it.  With the switch on, every reduction whose result FEEDS
 * LATER LAYERS is an ordered sum with one contributing block per output (the fused conv-epilogue statistics run as the
 * stand-alone bg_in_stats pass), so images, critic scores, activation gradients and the R1 image gradient are
 * bit-reproducible.  Leaf sums (weight-gradient split-K partials, bias / noise-weight gradients, loss terms) keep
 * their atomics: ~1e-6 relative order noise that ends in that tensor.  Used by the parity tests; costs a few percent. */
int bg_set_deterministic(int on);
int bg_get_deterministic(void);
/* SMs left free by the persistent / one-wave grids of this library (default 0, or BG_SM_RESERVE in the environment): in a
 * data-parallel run the gradient all-reduce (train.py:71,79's nn.DataParallel reduce, here NCCL) runs UNDER the backward
 * and its CTAs occupy a few SMs; grids sized for all 148 SMs would then spill into a second wave. */
int bg_set_sm_reserve(int sms);
/* Entry points that ACCUMULATE into a caller buffer with atomics (fused statistics, split-K weight gradients, channel sums)
 * zero that buffer first, one small launch each.  A caller that carves such buffers out of ONE slab it has already zeroed
 * (a single memset per backward pass) registers the slab here (device pointer, bytes; bytes = 0 clears; per host thread):
 * the zero fill is then skipped for every target that lies inside the range.  The caller must hand each carved buffer to
 * at most one zero-expecting call. */
int bg_set_prezeroed_range(const void* base, size_t bytes);

/* ---- equalized-lr weight staging (gan.py:14,27,32: weight * sqrt(2/fan_in) every forward) -------------
 * w: fp32 (Cout,Cin,ks,ks).  w_fprop: bf16 [ks*ks][Cout][Cin_pad] or NULL.
 * w_dgrad: bf16 [ks*ks flipped][Cin_pad][Cout] or NULL (same forward kernel then computes dL/dx). */
int bg_pack_weight(const float* w, void* w_fprop, void* w_dgrad, int Cout, int Cin, int Cin_pad, int ks, float coef,
                   void* stream);
/* bg_pack_weight for up to 32 layers in one launch (every conv weight of a network is stale after an optimizer step).
 * All arguments are HOST arrays with one entry per layer; both packs are written for every layer. */
int bg_pack_weight_grouped(const float* const* w, void* const* w_fprop, void* const* w_dgrad, const int* Cout,
                           const int* Cin, const int* Cin_pad, const int* ks, const float* coef, int groups, void* stream);
/* dw_packed: fp32 [ks*ks][Cout][Cin_pad] -> dw: fp32 (Cout,Cin,ks,ks) * coef (+= if accumulate). */
int bg_unpack_wgrad(const float* dw_packed, float* dw, int Cout, int Cin, int Cin_pad, int ks, float coef,
                    int accumulate, void* stream);

/* ---- EqualizedConv2d.forward, 3x3 pad 1 / 1x1 (gan.py:29-38 via :83,:240,:254,:259) ------------------
 * tcgen05 implicit GEMM.  out = gate(act(conv(x, wpack) + bias + noise_w[c] * noise[n,h,w]))
 *   bias, noise, noise_w, gate_src may be NULL.  noise: fp32 (N,1,H,W) (InjectSecondaryNoise, gan.py:52).
 *   act: 0 none, 1 LeakyReLU(slope) (gan.py:86,241,255).  gate_src: bf16 (N,H,W,Cout); the output is
 *   multiplied by (gate_src > 0 ? 1 : slope) — the LeakyReLU backward gate fused into a dgrad pass.
 * Run with w_dgrad as wpack (Cin/Cout swapped) this is autograd's conv input-gradient. */
int bg_conv_fprop(const void* x, const void* wpack, void* out, int N, int H, int W, int Cin, int Cout, int ksize,
                  const float* bias, const float* noise, const float* noise_w, const void* gate_src, int act,
                  float slope, void* stream);
/* bg_conv_fprop plus a per-channel reduction of the (bf16) output fused into the epilogue (no extra pass over
 * the map).  The buffer is zeroed by the call.
 *   stats_mode 1: stats fp32 (N,Cout,2) = [sum_hw out, sum_hw out^2] per sample — the nn.InstanceNorm2d statistics
 *                 of AdaINBlock (gan.py:59,69) for the activation this conv produces;
 *   stats_mode 2: stats fp32 (Cout)     = sum_{n,hw} out — when out is the gradient at a conv output (a dgrad pass),
 *                 this is that conv's bias gradient (autograd convolution_backward, bias term). */
int bg_conv_fprop_stats(const void* x, const void* wpack, void* out, int N, int H, int W, int Cin, int Cout, int ksize,
                        const float* bias, const float* noise, const float* noise_w, const void* gate_src, int act,
                        float slope, float* stats, int stats_mode, void* stream);
/* ---- StyleConvBlock.forward fused end to end (gan.py:89-98 inside StyleGanBlock.forward gan.py:118-127) -------
 * The PREVIOUS layer's AdaIN (gan.py:65-71) never materialises: bg_style_modulate folds its per-sample scale
 * gamma*rstd into a per-sample weight pack wmod bf16 [N][9][Cout][Cin] and its shift, pushed through the conv (with
 * the zero padding accounted for per border class) and added to the conv bias, into btab fp32 [N][9][Cout].
 *   stats: fp32 (N,Cin,2) sums of the previous activation over HW pixels; style: fp32 (N,2*Cin) = [gamma|beta].
 * bg_conv_style_fprop then computes
 *   out = LeakyReLU( conv3x3( [bilinear x2 upsample](x), wmod[n] ) + btab[n][class(h,w)] + noise_w[c]*noise[n,h,w] )
 * with the upsample (gan.py:112,123) built inside the operand feed when upsample=1 (x is then (N,H/2,W/2,Cin)), and
 * accumulates this layer's own instance-norm sums into stats fp32 (N,Cout,2) (zeroed by the call).  H,W >= 16. */
int bg_style_modulate(const float* W, const float* bias, const float* stats, const float* style, void* wmod, float* btab,
                      int N, int Cin, int Cout, int HW, float coef, float eps, void* stream);
int bg_conv_style_fprop(const void* x, const void* wmod, const float* btab, void* out, int N, int H, int W, int Cin,
                        int Cout, int upsample, const float* noise, const float* noise_w, float slope, float* stats,
                        void* stream);
/* toRGB 1x1 conv (gan.py:172-179,218,222) of AdaIN(a) without materialising it; out: fp32 (N,3,H,W). */
int bg_to_rgb_adain(const void* a, const float* stats, const float* style, const float* Wm, const float* bias, float* out,
                    int N, int HW, int C, float coef, float eps, void* stream);
/* conv3x3 -> AvgPool2d(2) -> [LeakyReLU | gate] fused in the epilogue (CriticBlock.conv_2, gan.py:258-262; with
 * act=0 and gate_src (pooled resolution) it is the R1 tangent pass through the same layers).  out: (N,H/2,W/2,Cout).
 * Needs H,W >= 16. */
int bg_conv_pool_fprop(const void* x, const void* wpack, void* out, int N, int H, int W, int Cin, int Cout,
                       const float* bias, const void* gate_src, int act, float slope, void* stream);
/* The same function as bg_conv_pool_fprop computed as ONE 4x4 stride-2 convolution (2.25x fewer MACs): w16 is the
 * 16-tap pack of bg_pack_weight_pool4 (quarter sums of the shifted 3x3 kernel, bf16 [16][Cout][Cin]).  x: (N,H,W,Cin),
 * out / gate_src: (N,H/2,W/2,Cout).  Needs H,W >= 32 and Cin % 32 == 0 (bg_conv_pool4_supported). */
int bg_pack_weight_pool4(const float* w, void* w16, int Cout, int Cin, float coef, void* stream);
int bg_conv_pool4_fprop(const void* x, const void* w16, void* out, int N, int H, int W, int Cin, int Cout, const float* bias,
                        const void* gate_src, int act, float slope, void* stream);
int bg_conv_pool4_supported(int N, int H, int W, int Cin, int Cout);
/* Input gradient of conv3x3 -> AvgPool2d(2) (autograd convolution_backward + avg_pool2d_backward) as the transposed
 * 4x4 stride-2 conv: gpool = gradient at the POOLED map (N,Hp,Wp,Cout), gx = gradient at the conv input
 * (N,2Hp,2Wp,Cin), optionally gated by gate_src (N,2Hp,2Wp,Cin) (LeakyReLU of the layer below) and with
 * bias_grad fp32 (Cin) = sum over gx (that layer's bias gradient).  wt: 16-tile pack of bg_pack_weight_tconv4. */
int bg_pack_weight_tconv4(const float* w, void* wt, int Cout, int Cin, float coef, void* stream);
int bg_conv_pool4_dgrad(const void* gpool, const void* wt, void* gx, int N, int Hp, int Wp, int Cout, int Cin,
                        const void* gate_src, float slope, float* bias_grad, void* stream);
/* Weight gradient of conv3x3 -> AvgPool2d(2) on the same 4x4 stride-2 form: dw16 fp32 [16][Cout][Cin] (overwritten, or
 * += when accumulate) = sum over pooled pixels of gpool[i,j,co] * x[2i+a-1, 2j+b-1, ci]; bg_unpack_wgrad_pool4 applies the
 * adjoint of the quarter-sum pack: dw (Cout,Cin,3,3) (+)= coef/4 * sum_{dy,dx} dw16[(ky+dy)*4 + kx+dx]. */
int bg_conv_pool4_wgrad(const void* x, const void* gpool, float* dw16, int N, int Hp, int Wp, int Cin, int Cout,
                        int accumulate, void* stream);
int bg_unpack_wgrad_pool4(const float* dw16, float* dw, int Cout, int Cin, float coef, int accumulate, void* stream);
/* bg_conv_fprop picks the halo-resident kernel (conv_halo.cu) for 3x3 at H,W >= 16 and the tap-wise TMA kernel
 * (conv_fprop.cu) otherwise; this entry forces the tap-wise kernel (A/B tests, small maps, 1x1). */
int bg_conv_fprop_tapwise(const void* x, const void* wpack, void* out, int N, int H, int W, int Cin, int Cout,
                          int ksize, const float* bias, const float* noise, const float* noise_w, const void* gate_src,
                          int act, float slope, void* stream);

/* ---- weight gradient of the 3x3 conv (autograd convolution_backward / _convolution_double_backward) ---
 * dw_packed: fp32 [9][Cout][Cin] (overwritten, or += when accumulate) = sum_pixels g[p,co] * x[p+tap,ci].
 * accumulate=1 is the second half of the R1 "doubled-K" weight gradient wgrad(x, ybar) + wgrad(v, ghat). */
int bg_conv_wgrad(const void* x, const void* g, float* dw_packed, int N, int H, int W, int Cin, int Cout,
                  int accumulate, void* stream);
/* bg_conv_wgrad picks the shifted-view kernel (conv_wgrad_halo.cu) at W >= 16 and the tap-wise one otherwise;
 * this entry forces the tap-wise kernel (A/B tests, small maps). */
int bg_conv_wgrad_tapwise(const void* x, const void* g, float* dw_packed, int N, int H, int W, int Cin, int Cout,
                          int accumulate, void* stream);

/* ---- LeakyReLU backward gate (gan.py:86,145,241...): out = g * (y > 0 ? 1 : slope); n elements -------- */
int bg_act_gate(const void* g, const void* y, void* out, size_t n, float slope, void* stream);
/* ---- out = ca*a + cb*b (b may be NULL): torch.lerp on feature maps (gan.py:347) and its gradients ----- */
int bg_axpby(const void* a, const void* b, void* out, size_t n, float ca, float cb, void* stream);

/* ---- AvgPool2d(2) -> LeakyReLU (CriticBlock tail, gan.py:258-262) --------------------------------------
 * fwd mode 0: y = lrelu(avg4(u)); mode 1: y = avg4(u) * gate(gate_src) (tangent of the backward).
 * bwd: gu = 0.25 * gate(y) * gy broadcast to the 2x2 window.  u,gu: (N,2Ho,2Wo,C); y,gy: (N,Ho,Wo,C). */
int bg_pool_act_fwd(const void* u, const void* gate_src, void* y, int N, int Ho, int Wo, int C, float slope,
                    int mode, void* stream);
/* csum (optional fp32 (C), zeroed by the call) += sum_{n,h,w} gu[.,c]: the bias gradient of the conv feeding the pool
 * (autograd convolution_backward, bias term), fused so the gradient map is not read again. */
int bg_pool_act_bwd(const void* gy, const void* y, void* gu, int N, int Ho, int Wo, int C, float slope, float* csum,
                    void* stream);

/* ---- nn.Upsample(scale_factor=2, bilinear) (gan.py:112,123) and its adjoint; x: (N,H,W,C) -------------- */
int bg_upsample2x_fwd(const void* x, void* y, int N, int H, int W, int C, void* stream);
int bg_upsample2x_bwd(const void* gy, void* gx, int N, int H, int W, int C, void* stream);

/* ---- per-channel weighted pixel sums (bias / noise-weight / fromRGB / toRGB weight gradients) ----------
 * out: fp32 [(1+nplanes)][C] (overwritten): out[0][c] = sum_p g[p,c]; out[1+j][c] = sum_p g[p,c]*plane_j[p],
 * plane_j[p] = planes[(p/HW)*img_stride + j*plane_stride + p%HW]. */
int bg_channel_wsum(const void* g, const float* planes, float* out, size_t P, int C, int HW, size_t img_stride,
                    size_t plane_stride, int nplanes, void* stream);

/* ---- fromRGB / toRGB 1x1 convolutions (gan.py:172-179, 351-355) ----------------------------------------
 * planes3_to_nhwc: out[p,c] = act(coef * sum_j img[n,j,hw] * Wm[c*ws_c + j*ws_j] + bias[c]) * gate(gate_src[p,c])
 *   (gate_src may be NULL; with it and act=0/bias=NULL this is the R1 tangent pass through fromRGB)
 * nhwc_to_planes3: out[n,j,hw] = coef * sum_c x[p,c] * Wm[c*ws_c + j*ws_j] + bias[j]
 * (weight (C,3,1,1): ws_c=3, ws_j=1;  weight (3,C,1,1): ws_c=1, ws_j=C) */
int bg_planes3_to_nhwc(const float* img, const float* Wm, const float* bias, const void* gate_src, void* out, size_t P,
                       int HW, int C, int ws_c, int ws_j, float coef, int act, float slope, void* stream);
int bg_nhwc_to_planes3(const void* x, const float* Wm, const float* bias, float* out, size_t P, int HW, int C,
                       int ws_c, int ws_j, float coef, void* stream);

/* ---- AdaINBlock (gan.py:55-71): InstanceNorm2d(eps=1e-8, biased var) then gamma*x+beta -----------------
 * stats: fp32 (N,C,2) = [sum a, sum a^2] over H*W (overwritten by bg_in_stats).
 * style: fp32 (N,2C) = [gamma | beta] (output of the style EqualizedLinear, gan.py:66-67).
 * bsums: fp32 (N,C,2) = [sum g, sum g*ahat] (these ARE dL/dbeta and dL/dgamma).
 * bwd_apply: out = gate(a) * gamma*rstd*(g - bsums0/HW - ahat*bsums1/HW); gate!=0 fuses the LeakyReLU
 * backward that precedes AdaIN in StyleConvBlock.forward (gan.py:96-98). */
int bg_in_stats(const void* a, float* stats, int N, int HW, int C, void* stream);
int bg_adain_apply(const void* a, const float* stats, const float* style, void* x, int N, int HW, int C, float eps,
                   void* stream);
int bg_adain_bwd_reduce(const void* g, const void* a, const float* stats, float* bsums, int N, int HW, int C,
                        float eps, void* stream);
/* noise (fp32 (N,1,H,W)) / wsum (fp32 (2,C), zeroed by the call) optional: wsum[0] = sum out = conv bias gradient,
 * wsum[1] = sum out * noise = InjectSecondaryNoise weight gradient (gan.py:52), reduced while out is written. */
int bg_adain_bwd_apply(const void* g, const void* a, const float* stats, const float* style, const float* bsums,
                       void* out, int N, int HW, int C, float eps, float slope, int gate, const float* noise,
                       float* wsum, void* stream);

/* ---- EqualizedLinear (gan.py:16-17): mapping network (gan.py:130-148), AdaIN style FCs (gan.py:60,66), critic
 * head FCs and, on the NCHW-flattened 4x4 map, the critic's 4x4 valid conv (gan.py:245-250).  All fp32.
 * fwd: y[M,N] = act(coef * x[M,K] W[N,K]^T + bias).  The input gradient is the same call on the transposed
 * weight (bg_transpose_f32).  bwd_weight: dW[N,K] (+)= coef * gy^T x;  db[N] (+)= sum_m gy (db may be NULL). */
int bg_linear_fwd(const float* x, const float* W, const float* bias, float* y, int M, int N, int K, float coef, int act,
                  float slope, void* stream);
/* Input gradient of y = coef * x W^T + b straight from the (N, K) row-major weight: gx (M, K) = coef * gy (M, N) @ W.
 * (autograd's addmm backward of gan.py:16-17; needs no transposed copy of W.)  K % 2 == 0. */
int bg_linear_bwd_input(const float* gy, const float* W, float* gx, int M, int N, int K, float coef, void* stream);
int bg_linear_bwd_weight(const float* gy, const float* x, float* dW, float* db, int M, int N, int K, float coef,
                         int accumulate, void* stream);
/* Grouped forms: several small layers in one launch.  W / bias / y / gy / dW / db / x / N / coef are HOST arrays with one
 * entry per layer (<= 16): device pointers, output widths N[g] (multiples of 4) and equalized coefficients.
 *   fwd:        y[g] (M,N[g]) = act(coef[g] * x W[g]^T + bias[g])       layers sharing their input x (M,K): the 2*steps
 *                                                                       AdaIN style FCs all read the latent w (gan.py:60,66)
 *   bwd_weight: dW[g] (N[g],K) = coef[g] * gy[g]^T x[g];  db[g] (N[g]) = sum_m gy[g]   (db or db[g] may be NULL); one
 *               input per layer (the same pointer 2*steps times for the style FCs; the 8 mapping layers, gan.py:130-148)
 *   bwd_input:  gx (M,K) = sum_g coef[g] * gy[g] W[g], W[g] the (N[g],K) weight as stored (no transposed copies);
 *               the n range is split over a thread-block cluster and reduced through DSMEM in rank order. */
int bg_linear_fwd_grouped(const float* x, const float* const* W, const float* const* bias, float* const* y, const int* N,
                          const float* coef, int groups, int M, int K, int act, float slope, void* stream);
int bg_linear_bwd_weight_grouped(const float* const* x, float* const* gy, float* const* dW, float* const* db, const int* N,
                                 const float* coef, int groups, int M, int K, void* stream);
int bg_linear_bwd_input_grouped(float* const* gy, const float* const* W, const int* N, const float* coef, int groups,
                                int M, int K, float* gx, void* stream);
int bg_transpose_f32(const float* in, float* out, int R, int C, void* stream);
int bg_act_gate_f32(const float* g, const float* y, float* out, size_t n, float slope, void* stream);
int bg_axpby_f32(const float* a, const float* b, float* out, size_t n, float ca, float cb, void* stream);

/* ---- learned 4x4 constant + noise + LeakyReLU (StyleConvBlock is_initial, gan.py:81,92,96-97) and its gradient */
int bg_const_noise_act(const float* cst, const float* noise, const float* nw, void* a, int N, int HW, int C,
                       float slope, void* stream);
int bg_const_bwd(const void* g, float* dconst, int N, int HW, int C, void* stream);

/* ---- image-plane fade ops, NCHW fp32 viewed as P = B*3 planes ------------------------------------------
 * avgpool2: F.avg_pool2d(images, 2) (gan.py:345) and its adjoint (gimg (+)= .25*scale*g).
 * up2_lerp: torch.lerp(bilinear_up2(small), large, alpha) (gan.py:213-220); up2_bwd: gsmall = scale * up2^T(g). */
int bg_img_avgpool2(const float* img, float* out, int P, int Ho, int Wo, void* stream);
int bg_img_avgpool2_bwd(const float* g, float* gimg, int P, int Ho, int Wo, float scale, int accumulate, void* stream);
int bg_img_up2_lerp(const float* small, const float* large, float* out, int P, int H, int W, float alpha,
                    void* stream);
int bg_img_up2_bwd(const float* g, float* gsmall, int P, int H, int W, float scale, void* stream);
/* sums[3] = per-colour sum of g (B,3,HW): toRGB bias gradient */
int bg_plane_sums(const float* g, float* sums, int B, int HW, void* stream);

/* ---- nn.Flatten boundary of the critic head (gan.py:247): NHWC bf16 <-> NCHW fp32 (gate_src may be NULL) */
int bg_nhwc_to_nchw_f32(const void* x, float* out, int N, int HW, int C, void* stream);
int bg_nchw_f32_to_nhwc(const float* g, const void* gate_src, void* out, int N, int HW, int C, float slope,
                        void* stream);

/* ---- MiniBatchStdDev (gan.py:273-298), x: (B,HW,C) bf16, G = group size, M = B/G ------------------------
 * fwd : plane[M] = s (v NULL) or the tangent s-dot along v (v given); xpad (B,HW,Cpad) = [x or v | plane | 0].
 * bwd : gx (B,HW,C) = gpad[..., :C] + J^T gs  (+ second-order term with tangent v and gs2 when v/gpad2 given),
 *       gs = per-slot sum of gpad[..., C], gs2 likewise from gpad2.  gs_ws: fp32 workspace of 2*M. */
int bg_mbstd_fwd(const void* x, const void* v, float* plane, void* xpad, int B, int G, int HW, int C, int Cpad,
                 float eps, void* stream);
int bg_mbstd_bwd(const void* x, const void* v, const void* gpad, const void* gpad2, float* gs_ws, void* gx, int B, int G,
                 int HW, int C, int Cpad, float eps, void* stream);

/* ---- loss terms (gan.py:228,396,406): loss[0] = mean softplus(sign*pred); seed[i] = seed_scale * dloss/dpred_i
 * (seed may be NULL).  sumsq: out[0] = scale * sum x^2 (the R1 penalty, gan.py:401-404). */
int bg_logistic_loss(const float* pred, int n, float sign, float* loss, float* seed, float seed_scale, void* stream);
int bg_sumsq(const float* x, size_t n, float scale, float* out, void* stream);
/* WGAN-GP penalty rows (gan.py:383-385: gp = mean_n (||gradient_n||_2 - 1)^2), g: fp32 (B, D):
 * pen[0] = pen_scale * sum_n (r_n - 1)^2, r_n = ||g_n||_2;  v[n] = v_scale * 2 (r_n - 1) / r_n * g_n, i.e. the
 * derivative of the penalty w.r.t. g_n — the direction the second-order pass is taken along. */
int bg_gp_rows(const float* g, int B, size_t D, float pen_scale, float v_scale, float* pen, float* v, void* stream);

/* ---- data feed (train.py:43-50: RandomHorizontalFlip, ToTensor, Normalize((.5,.5,.5),(.5,.5,.5)) on the device) -------
 * src: uint8 (B,H,W,3) device;  flip: uint8 (B) device or NULL (non-zero = mirror that sample along W);
 * out: fp32 (B,3,H,W) = src / 127.5 - 1. */
int bg_image_feed_u8(const void* src_u8, const void* flip_u8, float* out, int B, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* BG_B200_H_ */
