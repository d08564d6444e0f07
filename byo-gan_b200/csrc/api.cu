// extern "C" surface declared in include/bg_b200.h: thin forwarding to the launchers.
#include "common.cuh"
#include "../../include/bg_b200.h"

namespace bg {
int launch_conv_fprop(const void*, const void*, void*, int, int, int, int, int, int, const float*, const float*,
                      const float*, const void*, int, float, cudaStream_t);
int launch_conv_halo(const void*, const void*, void*, int, int, int, int, int, const float*, const float*, const float*,
                     const void*, int, int, float, float*, int, const float*, int, int, cudaStream_t);
bool conv_halo_supported(int, int, int, int, int, int);
bool conv_splitk_supported(int, int, int, int, int, int);
int launch_conv_splitk(const void*, const void*, void*, int, int, int, int, int, const float*, const float*, const float*,
                       const void*, int, float, cudaStream_t);
int launch_conv_wgrad(const void*, const void*, float*, int, int, int, int, int, int, cudaStream_t);
int launch_conv_wgrad_halo(const void*, const void*, float*, int, int, int, int, int, int, int, cudaStream_t);
int launch_unpack_wgrad_pool4(const float*, float*, int, int, float, int, cudaStream_t);
bool conv_wgrad_halo_supported(int, int, int, int, int);
int launch_pack_weight(const float*, void*, void*, int, int, int, int, float, cudaStream_t);
int launch_unpack_wgrad(const float*, float*, int, int, int, int, float, int, cudaStream_t);
int launch_act_gate(const void*, const void*, void*, size_t, float, cudaStream_t);
int launch_axpby(const void*, const void*, void*, size_t, float, float, cudaStream_t);
int launch_pool_act_fwd(const void*, const void*, void*, int, int, int, int, float, int, cudaStream_t);
int launch_pool_act_bwd(const void*, const void*, void*, int, int, int, int, float, float*, cudaStream_t);
int launch_upsample2x_fwd(const void*, void*, int, int, int, int, cudaStream_t);
int launch_upsample2x_bwd(const void*, void*, int, int, int, int, cudaStream_t);
int launch_channel_wsum(const void*, const float*, float*, size_t, int, int, size_t, size_t, int, cudaStream_t);
int launch_planes3_to_nhwc(const float*, const float*, const float*, const void*, void*, size_t, int, int, int, int,
                           float, int, float, cudaStream_t);
int launch_linear_fwd(const float*, const float*, const float*, float*, int, int, int, float, int, float, cudaStream_t);
int launch_linear_bwd_weight(const float*, const float*, float*, float*, int, int, int, float, int, cudaStream_t);
int launch_linear_bwd_input(const float*, const float*, float*, int, int, int, float, cudaStream_t);
int launch_transpose_f32(const float*, float*, int, int, cudaStream_t);
int launch_act_gate_f32(const float*, const float*, float*, size_t, float, cudaStream_t);
int launch_axpby_f32(const float*, const float*, float*, size_t, float, float, cudaStream_t);
int launch_const_noise_act(const float*, const float*, const float*, void*, int, int, int, float, cudaStream_t);
int launch_const_bwd(const void*, float*, int, int, int, cudaStream_t);
int launch_img_avgpool2(const float*, float*, int, int, int, cudaStream_t);
int launch_img_avgpool2_bwd(const float*, float*, int, int, int, float, int, cudaStream_t);
int launch_img_up2_lerp(const float*, const float*, float*, int, int, int, float, cudaStream_t);
int launch_img_up2_bwd(const float*, float*, int, int, int, float, cudaStream_t);
int launch_plane_sums(const float*, float*, int, int, cudaStream_t);
int launch_nhwc_to_nchw_f32(const void*, float*, int, int, int, cudaStream_t);
int launch_nchw_f32_to_nhwc(const float*, const void*, void*, int, int, int, float, cudaStream_t);
int launch_mbstd_fwd(const void*, const void*, float*, void*, int, int, int, int, int, float, cudaStream_t);
int launch_mbstd_bwd(const void*, const void*, const void*, const void*, float*, void*, int, int, int, int, int, float,
                     cudaStream_t);
int launch_logistic_loss(const float*, int, float, float*, float*, float, cudaStream_t);
int launch_sumsq(const float*, size_t, float, float*, cudaStream_t);
int launch_image_feed_u8(const void*, const void*, float*, int, int, int, cudaStream_t);
int launch_gp_rows(const float*, int, size_t, float, float, float*, float*, cudaStream_t);
int launch_nhwc_to_planes3(const void*, const float*, const float*, float*, size_t, int, int, int, int, float,
                           cudaStream_t);
int launch_in_stats(const void*, float*, int, int, int, cudaStream_t);
int launch_adain_apply(const void*, const float*, const float*, void*, int, int, int, float, cudaStream_t);
int launch_adain_bwd_reduce(const void*, const void*, const float*, float*, int, int, int, float, cudaStream_t);
int launch_adain_bwd_apply(const void*, const void*, const float*, const float*, const float*, void*, int, int, int,
                           float, float, int, const float*, float*, cudaStream_t);
int launch_style_modulate(const float*, const float*, const float*, const float*, void*, float*, int, int, int, int, float,
                          float, cudaStream_t);
int launch_to_rgb_adain(const void*, const float*, const float*, const float*, const float*, float*, int, int, int, float,
                        float, cudaStream_t);
int launch_linear_grouped(int, const float*, const float* const*, const float* const*, const float* const*, float* const*,
                          float* const*, float* const*, const int*, const float*, int, int, int, int, float, float*,
                          cudaStream_t);
int launch_pack_weight_grouped(const float* const*, void* const*, void* const*, const int*, const int*, const int*,
                               const int*, const float*, int, cudaStream_t);
int launch_pack_weight_pool4(const float*, void*, int, int, float, cudaStream_t);
int launch_pack_weight_tconv4(const float*, void*, int, int, float, cudaStream_t);
}  // namespace bg

#define S(stream) reinterpret_cast<cudaStream_t>(stream)

extern "C" {

int bg_pack_weight(const float* w, void* wf, void* wd, int Cout, int Cin, int Cin_pad, int ks, float coef,
                   void* stream) {
  return bg::launch_pack_weight(w, wf, wd, Cout, Cin, Cin_pad, ks, coef, S(stream));
}
int bg_pack_weight_grouped(const float* const* w, void* const* wf, void* const* wd, const int* Cout, const int* Cin,
                           const int* Cin_pad, const int* ks, const float* coef, int groups, void* stream) {
  return bg::launch_pack_weight_grouped(w, wf, wd, Cout, Cin, Cin_pad, ks, coef, groups, S(stream));
}
int bg_unpack_wgrad(const float* dwp, float* dw, int Cout, int Cin, int Cin_pad, int ks, float coef, int accumulate,
                    void* stream) {
  return bg::launch_unpack_wgrad(dwp, dw, Cout, Cin, Cin_pad, ks, coef, accumulate, S(stream));
}
int bg_conv_fprop(const void* x, const void* wpack, void* out, int N, int H, int W, int Cin, int Cout, int ksize,
                  const float* bias, const float* noise, const float* noise_w, const void* gate_src, int act,
                  float slope, void* stream) {
  if (bg::conv_halo_supported(N, H, W, Cin, Cout, ksize))
    return bg::launch_conv_halo(x, wpack, out, N, H, W, Cin, Cout, bias, noise, noise_w, gate_src, act, 0, slope,
                                nullptr, 0, nullptr, 0, 0, S(stream));
  // 4x4 / 8x8 maps: split-K over a thread-block cluster with a distributed-shared-memory reduction (conv_splitk.cu)
  if (bg::conv_splitk_supported(N, H, W, Cin, Cout, ksize))
    return bg::launch_conv_splitk(x, wpack, out, N, H, W, Cin, Cout, bias, noise, noise_w, gate_src, act, slope, S(stream));
  return bg::launch_conv_fprop(x, wpack, out, N, H, W, Cin, Cout, ksize, bias, noise, noise_w, gate_src, act, slope,
                               S(stream));
}
int bg_conv_fprop_stats(const void* x, const void* wpack, void* out, int N, int H, int W, int Cin, int Cout, int ksize,
                        const float* bias, const float* noise, const float* noise_w, const void* gate_src, int act,
                        float slope, float* stats, int stats_mode, void* stream) {
  if (stats == nullptr || (stats_mode != 1 && stats_mode != 2)) {
    bg::set_error("conv_fprop_stats: stats must be non-NULL and stats_mode 1 or 2 (got %d)", stats_mode);
    return 2;
  }
  if (bg::conv_halo_supported(N, H, W, Cin, Cout, ksize)) {
    if (stats_mode == 1 && bg::deterministic()) {
      // chain-deterministic mode: the statistics feed the next layer, so they are summed in a fixed order by the
      // stand-alone reduction (one block per sample) instead of the epilogue's cross-CTA fp32 atomics
      const int rc = bg::launch_conv_halo(x, wpack, out, N, H, W, Cin, Cout, bias, noise, noise_w, gate_src, act, 0,
                                          slope, nullptr, 0, nullptr, 0, 0, S(stream));
      return rc != 0 ? rc : bg::launch_in_stats(out, stats, N, H * W, Cout, S(stream));
    }
    return bg::launch_conv_halo(x, wpack, out, N, H, W, Cin, Cout, bias, noise, noise_w, gate_src, act, 0, slope, stats,
                                stats_mode, nullptr, 0, 0, S(stream));
  }
  // small maps (< 16x16): split-K cluster kernel (or the tap-wise one), then the stand-alone reduction over the (tiny) output
  int rc = bg::conv_splitk_supported(N, H, W, Cin, Cout, ksize)
               ? bg::launch_conv_splitk(x, wpack, out, N, H, W, Cin, Cout, bias, noise, noise_w, gate_src, act, slope, S(stream))
               : bg::launch_conv_fprop(x, wpack, out, N, H, W, Cin, Cout, ksize, bias, noise, noise_w, gate_src, act, slope,
                                       S(stream));
  if (rc != 0) return rc;
  if (stats_mode == 1) return bg::launch_in_stats(out, stats, N, H * W, Cout, S(stream));
  return bg::launch_channel_wsum(out, nullptr, stats, (size_t)N * H * W, Cout, H * W, 0, 0, 0, S(stream));
}
int bg_conv_style_fprop(const void* x, const void* wmod, const float* btab, void* out, int N, int H, int W, int Cin,
                        int Cout, int upsample, const float* noise, const float* noise_w, float slope, float* stats,
                        void* stream) {
  if (!bg::conv_halo_supported(N, H, W, Cin, Cout, 3)) {
    bg::set_error("conv_style_fprop: needs a 3x3 layer at H,W >= 16 (got H %d W %d Cin %d Cout %d)", H, W, Cin, Cout);
    return 2;
  }
  if (wmod == nullptr || btab == nullptr) {
    bg::set_error("conv_style_fprop: wmod and btab (from bg_style_modulate) are required");
    return 2;
  }
  if (stats != nullptr && bg::deterministic()) {
    const int rc = bg::launch_conv_halo(x, wmod, out, N, H, W, Cin, Cout, nullptr, noise, noise_w, nullptr, 1, 0, slope,
                                        nullptr, 0, btab, 1, upsample, S(stream));
    return rc != 0 ? rc : bg::launch_in_stats(out, stats, N, H * W, Cout, S(stream));
  }
  return bg::launch_conv_halo(x, wmod, out, N, H, W, Cin, Cout, nullptr, noise, noise_w, nullptr, 1, 0, slope, stats, 1,
                              btab, 1, upsample, S(stream));
}
int bg_style_modulate(const float* W, const float* bias, const float* stats, const float* style, void* wmod, float* btab,
                      int N, int Cin, int Cout, int HW, float coef, float eps, void* stream) {
  return bg::launch_style_modulate(W, bias, stats, style, wmod, btab, N, Cin, Cout, HW, coef, eps, S(stream));
}
int bg_to_rgb_adain(const void* a, const float* stats, const float* style, const float* Wm, const float* bias, float* out,
                    int N, int HW, int C, float coef, float eps, void* stream) {
  return bg::launch_to_rgb_adain(a, stats, style, Wm, bias, out, N, HW, C, coef, eps, S(stream));
}
int bg_conv_pool_fprop(const void* x, const void* wpack, void* out, int N, int H, int W, int Cin, int Cout,
                       const float* bias, const void* gate_src, int act, float slope, void* stream) {
  if (!bg::conv_halo_supported(N, H, W, Cin, Cout, 3)) {
    bg::set_error("conv_pool_fprop: needs a 3x3 layer at H,W >= 16 (got H %d W %d Cin %d Cout %d)", H, W, Cin, Cout);
    return 2;
  }
  return bg::launch_conv_halo(x, wpack, out, N, H, W, Cin, Cout, bias, nullptr, nullptr, gate_src, act, 1, slope,
                              nullptr, 0, nullptr, 0, 0, S(stream));
}
int bg_pack_weight_pool4(const float* w, void* w16, int Cout, int Cin, float coef, void* stream) {
  return bg::launch_pack_weight_pool4(w, w16, Cout, Cin, coef, S(stream));
}
int bg_conv_pool4_fprop(const void* x, const void* w16, void* out, int N, int H, int W, int Cin, int Cout, const float* bias,
                        const void* gate_src, int act, float slope, void* stream) {
  if (!bg_conv_pool4_supported(N, H, W, Cin, Cout)) {
    bg::set_error("conv_pool4_fprop: needs H,W >= 32 (powers of two), Cin %% 32 == 0, Cout %% 16 == 0 (H %d W %d Cin %d Cout %d)",
                  H, W, Cin, Cout);
    return 2;
  }
  return bg::launch_conv_halo(x, w16, out, N, H, W, Cin, Cout, bias, nullptr, nullptr, gate_src, act, 2, slope, nullptr, 0,
                              nullptr, 0, 0, S(stream));
}
int bg_pack_weight_tconv4(const float* w, void* wt, int Cout, int Cin, float coef, void* stream) {
  return bg::launch_pack_weight_tconv4(w, wt, Cout, Cin, coef, S(stream));
}
int bg_conv_pool4_dgrad(const void* gpool, const void* wt, void* gx, int N, int Hp, int Wp, int Cout, int Cin,
                        const void* gate_src, float slope, float* bias_grad, void* stream) {
  if (!bg::conv_halo_supported(N, Hp, Wp, Cout, Cin, 3)) {
    bg::set_error("conv_pool4_dgrad: needs a pooled map >= 16x16 (powers of two), channels %% 16 == 0 (Hp %d Wp %d Cout %d Cin %d)",
                  Hp, Wp, Cout, Cin);
    return 2;
  }
  return bg::launch_conv_halo(gpool, wt, gx, N, Hp, Wp, Cout, Cin, nullptr, nullptr, nullptr, gate_src, 0, 3, slope,
                              bias_grad, 2, nullptr, 0, 0, S(stream));
}
int bg_conv_pool4_wgrad(const void* x, const void* gpool, float* dw16, int N, int Hp, int Wp, int Cin, int Cout,
                        int accumulate, void* stream) {
  if (!bg::conv_wgrad_halo_supported(N, Hp, Wp, Cin, Cout) || Hp < 8 || Wp < 16) {
    bg::set_error("conv_pool4_wgrad: unsupported shape (pooled %d x %d, Cin %d, Cout %d)", Hp, Wp, Cin, Cout);
    return 2;
  }
  return bg::launch_conv_wgrad_halo(x, gpool, dw16, N, Hp, Wp, Cin, Cout, accumulate, 1, S(stream));
}
int bg_unpack_wgrad_pool4(const float* dw16, float* dw, int Cout, int Cin, float coef, int accumulate, void* stream) {
  return bg::launch_unpack_wgrad_pool4(dw16, dw, Cout, Cin, coef, accumulate, S(stream));
}
int bg_conv_pool4_supported(int N, int H, int W, int Cin, int Cout) {
  return (H >= 32 && W >= 32 && Cin % 32 == 0 && bg::conv_halo_supported(N, H / 2, W / 2, Cin, Cout, 3)) ? 1 : 0;
}
int bg_conv_fprop_tapwise(const void* x, const void* wpack, void* out, int N, int H, int W, int Cin, int Cout, int ksize,
                          const float* bias, const float* noise, const float* noise_w, const void* gate_src, int act,
                          float slope, void* stream) {
  return bg::launch_conv_fprop(x, wpack, out, N, H, W, Cin, Cout, ksize, bias, noise, noise_w, gate_src, act, slope,
                               S(stream));
}
int bg_conv_wgrad(const void* x, const void* g, float* dwp, int N, int H, int W, int Cin, int Cout, int accumulate,
                  void* stream) {
  if (bg::conv_wgrad_halo_supported(N, H, W, Cin, Cout))
    return bg::launch_conv_wgrad_halo(x, g, dwp, N, H, W, Cin, Cout, accumulate, 0, S(stream));
  return bg::launch_conv_wgrad(x, g, dwp, N, H, W, Cin, Cout, accumulate, S(stream));
}
int bg_conv_wgrad_tapwise(const void* x, const void* g, float* dwp, int N, int H, int W, int Cin, int Cout,
                          int accumulate, void* stream) {
  return bg::launch_conv_wgrad(x, g, dwp, N, H, W, Cin, Cout, accumulate, S(stream));
}
int bg_act_gate(const void* g, const void* y, void* out, size_t n, float slope, void* stream) {
  return bg::launch_act_gate(g, y, out, n, slope, S(stream));
}
int bg_axpby(const void* a, const void* b, void* out, size_t n, float ca, float cb, void* stream) {
  return bg::launch_axpby(a, b, out, n, ca, cb, S(stream));
}
int bg_pool_act_fwd(const void* u, const void* gate_src, void* y, int N, int Ho, int Wo, int C, float slope, int mode,
                    void* stream) {
  return bg::launch_pool_act_fwd(u, gate_src, y, N, Ho, Wo, C, slope, mode, S(stream));
}
int bg_pool_act_bwd(const void* gy, const void* y, void* gu, int N, int Ho, int Wo, int C, float slope, float* csum,
                    void* stream) {
  return bg::launch_pool_act_bwd(gy, y, gu, N, Ho, Wo, C, slope, csum, S(stream));
}
int bg_upsample2x_fwd(const void* x, void* y, int N, int H, int W, int C, void* stream) {
  return bg::launch_upsample2x_fwd(x, y, N, H, W, C, S(stream));
}
int bg_upsample2x_bwd(const void* gy, void* gx, int N, int H, int W, int C, void* stream) {
  return bg::launch_upsample2x_bwd(gy, gx, N, H, W, C, S(stream));
}
int bg_channel_wsum(const void* g, const float* planes, float* out, size_t P, int C, int HW, size_t img_stride,
                    size_t plane_stride, int nplanes, void* stream) {
  return bg::launch_channel_wsum(g, planes, out, P, C, HW, img_stride, plane_stride, nplanes, S(stream));
}
int bg_planes3_to_nhwc(const float* img, const float* Wm, const float* bias, const void* gate_src, void* out, size_t P,
                       int HW, int C, int ws_c, int ws_j, float coef, int act, float slope, void* stream) {
  return bg::launch_planes3_to_nhwc(img, Wm, bias, gate_src, out, P, HW, C, ws_c, ws_j, coef, act, slope, S(stream));
}
int bg_nhwc_to_planes3(const void* x, const float* Wm, const float* bias, float* out, size_t P, int HW, int C,
                       int ws_c, int ws_j, float coef, void* stream) {
  return bg::launch_nhwc_to_planes3(x, Wm, bias, out, P, HW, C, ws_c, ws_j, coef, S(stream));
}
int bg_in_stats(const void* a, float* stats, int N, int HW, int C, void* stream) {
  return bg::launch_in_stats(a, stats, N, HW, C, S(stream));
}
int bg_adain_apply(const void* a, const float* stats, const float* style, void* x, int N, int HW, int C, float eps,
                   void* stream) {
  return bg::launch_adain_apply(a, stats, style, x, N, HW, C, eps, S(stream));
}
int bg_adain_bwd_reduce(const void* g, const void* a, const float* stats, float* bsums, int N, int HW, int C,
                        float eps, void* stream) {
  return bg::launch_adain_bwd_reduce(g, a, stats, bsums, N, HW, C, eps, S(stream));
}
int bg_adain_bwd_apply(const void* g, const void* a, const float* stats, const float* style, const float* bsums,
                       void* out, int N, int HW, int C, float eps, float slope, int gate, const float* noise,
                       float* wsum, void* stream) {
  return bg::launch_adain_bwd_apply(g, a, stats, style, bsums, out, N, HW, C, eps, slope, gate, noise, wsum, S(stream));
}

int bg_linear_fwd_grouped(const float* x, const float* const* W, const float* const* bias, float* const* y, const int* N,
                          const float* coef, int groups, int M, int K, int act, float slope, void* stream) {
  return bg::launch_linear_grouped(0, x, nullptr, W, bias, y, nullptr, nullptr, N, coef, groups, M, K, act, slope, nullptr,
                                   S(stream));
}
int bg_linear_bwd_weight_grouped(const float* const* x, float* const* gy, float* const* dW, float* const* db, const int* N,
                                 const float* coef, int groups, int M, int K, void* stream) {
  return bg::launch_linear_grouped(1, nullptr, x, reinterpret_cast<const float* const*>(dW), nullptr, gy, dW, db, N, coef,
                                   groups, M, K, 0, 0.f, nullptr, S(stream));
}
int bg_linear_bwd_input_grouped(float* const* gy, const float* const* W, const int* N, const float* coef, int groups,
                                int M, int K, float* gx, void* stream) {
  return bg::launch_linear_grouped(2, nullptr, nullptr, W, nullptr, gy, nullptr, nullptr, N, coef, groups, M, K, 0, 0.f, gx,
                                   S(stream));
}
int bg_linear_fwd(const float* x, const float* W, const float* bias, float* y, int M, int N, int K, float coef, int act,
                  float slope, void* stream) {
  return bg::launch_linear_fwd(x, W, bias, y, M, N, K, coef, act, slope, S(stream));
}
int bg_linear_bwd_input(const float* gy, const float* W, float* gx, int M, int N, int K, float coef, void* stream) {
  return bg::launch_linear_bwd_input(gy, W, gx, M, N, K, coef, S(stream));
}
int bg_linear_bwd_weight(const float* gy, const float* x, float* dW, float* db, int M, int N, int K, float coef,
                         int accumulate, void* stream) {
  return bg::launch_linear_bwd_weight(gy, x, dW, db, M, N, K, coef, accumulate, S(stream));
}
int bg_transpose_f32(const float* in, float* out, int R, int C, void* stream) {
  return bg::launch_transpose_f32(in, out, R, C, S(stream));
}
int bg_act_gate_f32(const float* g, const float* y, float* out, size_t n, float slope, void* stream) {
  return bg::launch_act_gate_f32(g, y, out, n, slope, S(stream));
}
int bg_axpby_f32(const float* a, const float* b, float* out, size_t n, float ca, float cb, void* stream) {
  return bg::launch_axpby_f32(a, b, out, n, ca, cb, S(stream));
}
int bg_const_noise_act(const float* cst, const float* noise, const float* nw, void* a, int N, int HW, int C,
                       float slope, void* stream) {
  return bg::launch_const_noise_act(cst, noise, nw, a, N, HW, C, slope, S(stream));
}
int bg_const_bwd(const void* g, float* dconst, int N, int HW, int C, void* stream) {
  return bg::launch_const_bwd(g, dconst, N, HW, C, S(stream));
}
int bg_img_avgpool2(const float* img, float* out, int P, int Ho, int Wo, void* stream) {
  return bg::launch_img_avgpool2(img, out, P, Ho, Wo, S(stream));
}
int bg_img_avgpool2_bwd(const float* g, float* gimg, int P, int Ho, int Wo, float scale, int accumulate,
                        void* stream) {
  return bg::launch_img_avgpool2_bwd(g, gimg, P, Ho, Wo, scale, accumulate, S(stream));
}
int bg_img_up2_lerp(const float* small, const float* large, float* out, int P, int H, int W, float alpha,
                    void* stream) {
  return bg::launch_img_up2_lerp(small, large, out, P, H, W, alpha, S(stream));
}
int bg_img_up2_bwd(const float* g, float* gsmall, int P, int H, int W, float scale, void* stream) {
  return bg::launch_img_up2_bwd(g, gsmall, P, H, W, scale, S(stream));
}
int bg_plane_sums(const float* g, float* sums, int B, int HW, void* stream) {
  return bg::launch_plane_sums(g, sums, B, HW, S(stream));
}
int bg_nhwc_to_nchw_f32(const void* x, float* out, int N, int HW, int C, void* stream) {
  return bg::launch_nhwc_to_nchw_f32(x, out, N, HW, C, S(stream));
}
int bg_nchw_f32_to_nhwc(const float* g, const void* gate_src, void* out, int N, int HW, int C, float slope,
                        void* stream) {
  return bg::launch_nchw_f32_to_nhwc(g, gate_src, out, N, HW, C, slope, S(stream));
}
int bg_mbstd_fwd(const void* x, const void* v, float* plane, void* xpad, int B, int G, int HW, int C, int Cpad,
                 float eps, void* stream) {
  return bg::launch_mbstd_fwd(x, v, plane, xpad, B, G, HW, C, Cpad, eps, S(stream));
}
int bg_mbstd_bwd(const void* x, const void* v, const void* gpad, const void* gpad2, float* gs_ws, void* gx, int B, int G,
                 int HW, int C, int Cpad, float eps, void* stream) {
  return bg::launch_mbstd_bwd(x, v, gpad, gpad2, gs_ws, gx, B, G, HW, C, Cpad, eps, S(stream));
}
int bg_logistic_loss(const float* pred, int n, float sign, float* loss, float* seed, float seed_scale, void* stream) {
  return bg::launch_logistic_loss(pred, n, sign, loss, seed, seed_scale, S(stream));
}
int bg_sumsq(const float* x, size_t n, float scale, float* out, void* stream) {
  return bg::launch_sumsq(x, n, scale, out, S(stream));
}
int bg_image_feed_u8(const void* src_u8, const void* flip_u8, float* out, int B, int H, int W, void* stream) {
  return bg::launch_image_feed_u8(src_u8, flip_u8, out, B, H, W, S(stream));
}
int bg_gp_rows(const float* g, int B, size_t D, float pen_scale, float v_scale, float* pen, float* v, void* stream) {
  return bg::launch_gp_rows(g, B, D, pen_scale, v_scale, pen, v, S(stream));
}

}  // extern "C"
