// Split-K implicit-GEMM convolution for the SMALL maps (4x4, 8x8; 3x3 pad 1), one thread-block CLUSTER per output tile.
//
// Same contract as conv_fprop.cu (EqualizedConv2d.forward gan.py:29-38 on the low-resolution blocks gan.py:159-161 /
// 325-327, its input gradient on the flipped pack, the R1 tangent pass).  Those layers are 512 x 512 channels on 16..64
// pixels per image: 1 % of the iteration's FLOPs, but the tap-wise kernel needs 40 us for each of its 33 launches (10 % of
// the iteration, profiles/r2_call_times_train256_b32.txt) because an output-stationary CTA has to stream its whole
// K = 9 * Cin operand rows through the L2 -> shared-memory path (~42 B/clk per SM): 128 CTAs x 1.8 MB for 7 MB of operands.
//
// Here the K loop of one (128-pixel, N_t-channel) output tile is split over the S CTAs of a cluster (S <= 8): CTA r runs
// k-blocks [r KB / S, (r + 1) KB / S) into its own TMEM accumulator, so a CTA streams 1 / S of the rows and S times as many
// SMs pull on L2.  The partial accumulators are then reduced THROUGH DISTRIBUTED SHARED MEMORY, without atomics and in a
// fixed order (the results are bit-reproducible, which the chain-deterministic mode needs: these are forward
// activations): every CTA spills its 128 x N_t fp32 partial into its own shared memory (aliased over the drained pipeline
// stages), the cluster synchronises, and CTA r sums rows [128 r / S, 128 (r + 1) / S) of all S partials with
// ld.shared::cluster, applies bias / noise / LeakyReLU / gate and stores the bf16 rows.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM allocator, 2..5 = epilogue (TMEM -> smem partial, then reduction).
#include "common.cuh"

#include <stdlib.h>

namespace bg {

namespace {

constexpr int kEpiWarps = 4;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kMaxStages = 6;
constexpr uint32_t kTmemCols = 256;

struct SplitKParams {
  int N, H, W, Cin, Cout;
  int bw, bh, bn;                  // pixel tile = bw x bh x bn images = 128 rows
  int tiles_w, tiles_h, tiles_n;
  int block_n, n_blocks;
  int kc, k_chunks, k_blocks;      // k-block = (channel chunk, tap)
  int stages, csize;
  int rows_per_cta, cols_per_thread;
  uint32_t a_bytes, b_bytes, stage_bytes, part_stride;   // part_stride: floats per row of the partial buffer
  const float* bias;
  const float* noise;
  const float* noise_w;
  const __nv_bfloat16* gate_src;
  __nv_bfloat16* out;
  int act;
  float slope;
};

__device__ __forceinline__ float4 ld_cluster_f4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr));
  return v;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_splitk_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                   const SplitKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  // [pipeline stages | partial accumulator (aliases the stages once they are drained)] [aux]
  const size_t pipe_bytes = (size_t)p.stages * p.stage_bytes;
  const size_t part_bytes = (size_t)128 * p.part_stride * sizeof(float);
  uint8_t* aux = smem + (pipe_bytes > part_bytes ? pipe_bytes : part_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* done_bar = empty_bar + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  float* part = reinterpret_cast<float*>(smem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int S = p.csize;

  // tile of this cluster
  const int tile = (int)blockIdx.x / S;
  const int nb = tile % p.n_blocks;
  int pt = tile / p.n_blocks;
  const int tw = pt % p.tiles_w;
  pt /= p.tiles_w;
  const int th = pt % p.tiles_h;
  const int tn = pt / p.tiles_h;
  const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn, co0 = nb * p.block_n;
  // this CTA's share of the K loop
  const int kb0 = (int)(((long long)p.k_blocks * crank) / S);
  const int kb1 = (int)(((long long)p.k_blocks * (crank + 1)) / S);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
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
    pdl_launch_dependents();       // only once the TMEM columns are taken (see conv_fprop.cu)
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int kcx = kb / 9, tap = kb - kcx * 9;
        const int ky = tap / 3, kx = tap - ky * 3;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
        mbar_expect_tx(&full_bar[stage], p.a_bytes + p.b_bytes);
        tma_load_4d(&tmap_x, &full_bar[stage], sa, kcx * p.kc, w0 + kx - 1, h0 + ky - 1, n0);
        tma_load_3d(&tmap_w, &full_bar[stage], sa + p.a_bytes, kcx * p.kc, co0, tap);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.block_n, 0, 0);
      const int ksteps = p.kc / 16;
      const uint32_t sbo = 8u * (uint32_t)p.kc * 2u;
      const uint32_t layout = p.kc == 64 ? 2u : (p.kc == 32 ? 4u : 6u);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
        const uint32_t sb = sa + p.a_bytes;
        for (int k = 0; k < ksteps; ++k)
          tc_mma_bf16(tmem_base, umma_desc(sa + k * 32, 16, sbo, layout), umma_desc(sb + k * 32, 16, sbo, layout), idesc,
                      (kb > kb0 || k > 0) ? 1u : 0u);
        tc_commit(&empty_bar[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      tc_commit(done_bar);          // every MMA of this CTA has retired: the stages are drained, the accumulator is final
    }
  } else {
    // ---- phase 1: this CTA's partial accumulator TMEM -> its own shared memory (fp32, row-major, padded rows)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* prow = part + (size_t)row * p.part_stride;
    for (int c = 0; c < p.block_n; c += 16) {
      uint32_t v[16];
      tmem_ld_x16(taddr + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(prow + c + j) =
            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // all S partials are in shared memory and visible cluster-wide

  if (warp >= 2) {
    // ---- phase 2: rows [crank * rows_per_cta, ...) summed over the S partials in CTA order, epilogue, bf16 store
    const int t = (int)threadIdx.x - 64;                     // 0..127
    const int tpr = 128 / p.rows_per_cta;                    // threads per row (= S)
    const int row = (int)crank * p.rows_per_cta + t / tpr;
    const int c_begin = (t % tpr) * p.cols_per_thread;
    const int wi = row % p.bw, hi = (row / p.bw) % p.bh, ni = row / (p.bw * p.bh);
    const int n = n0 + ni, h = h0 + hi, w = w0 + wi;
    const bool valid = n < p.N && h < p.H && w < p.W;
    const size_t pix = ((size_t)n * p.H + h) * p.W + w;
    const float nz = (p.noise != nullptr && valid) ? p.noise[pix] : 0.f;
    const uint32_t local = smem_u32(part + (size_t)row * p.part_stride + c_begin);
    uint32_t src[8];
    for (int s = 0; s < S; ++s) src[s] = mapa_cta(local, (uint32_t)s);
    for (int c = 0; c < p.cols_per_thread; c += 8) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.f;
      for (int s = 0; s < S; ++s) {                          // fixed order: bit-reproducible
        const float4 a = ld_cluster_f4(src[s] + (uint32_t)c * 4u), b = ld_cluster_f4(src[s] + (uint32_t)c * 4u + 16u);
        f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w;
        f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
      }
      if (valid) {
        const int co = co0 + c_begin + c;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float x = f[j];
          if (p.bias != nullptr) x += __ldg(p.bias + co + j);
          if (p.noise != nullptr) x = fmaf(__ldg(p.noise_w + co + j), nz, x);
          if (p.act) x = fmaxf(x, x * p.slope);
          f[j] = x;
        }
        if (p.gate_src != nullptr) {
          const uint4 gv = *reinterpret_cast<const uint4*>(p.gate_src + pix * p.Cout + co);
          const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 g2 = unpack_bf16x2(gw[j]);
            f[2 * j] *= g2.x > 0.f ? 1.f : p.slope;
            f[2 * j + 1] *= g2.y > 0.f ? 1.f : p.slope;
          }
        }
        uint4 o;
        o.x = pack_bf16x2(f[0], f[1]);
        o.y = pack_bf16x2(f[2], f[3]);
        o.z = pack_bf16x2(f[4], f[5]);
        o.w = pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(p.out + pix * p.Cout + co) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // peers may still be reading this CTA's partial
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

// Picks (N tile, cluster size) for a small-map layer; returns false when the split-K kernel does not apply.
static bool plan_splitk(int N, int H, int W, int Cin, int Cout, int* block_n_out, int* csize_out) {
  if (!(H <= 8 && W <= 8 && (H & (H - 1)) == 0 && (W & (W - 1)) == 0 && Cin % 64 == 0 && Cout % 16 == 0)) return false;
  const int bw = W, bh = H, bn = 128 / (bw * bh);
  const int pixel_tiles = (N + bn - 1) / bn;
  const int k_blocks = (Cin / 64) * 9;
  long best_cost = -1;
  int best_n = 0, best_s = 1;
  for (int nt = 256; nt >= 16; nt -= 16) {
    if (Cout % nt != 0) continue;
    const int tiles = pixel_tiles * (Cout / nt);
    for (int s = 8; s >= 1; s >>= 1) {
      if (tiles * s > num_sms() || s > k_blocks) continue;
      if ((128 / s) < 1 || (nt / s) % 8 != 0) continue;        // the reduction gives each thread whole 16-byte chunks
      // time ~ rows streamed per CTA: (128 + N_t) rows for each of its k-blocks
      const long cost = (long)(128 + nt) * ((k_blocks + s - 1) / s);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_n = nt; best_s = s; }
    }
  }
  if (best_cost < 0 || best_s < 2) return false;
  *block_n_out = best_n;
  *csize_out = best_s;
  return true;
}

bool conv_splitk_supported(int N, int H, int W, int Cin, int Cout, int ksize) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("BG_SPLITK"); on = (e && e[0] == '0') ? 0 : 1; }
  int bn, cs;
  return on && ksize == 3 && N > 0 && plan_splitk(N, H, W, Cin, Cout, &bn, &cs);
}

int launch_conv_splitk(const void* x, const void* wpack, void* out, int N, int H, int W, int Cin, int Cout,
                       const float* bias, const float* noise, const float* noise_w, const void* gate_src, int act,
                       float slope, cudaStream_t stream) {
  SplitKParams p;
  memset(&p, 0, sizeof(p));
  int block_n = 0, csize = 1;
  BG_REQUIRE(plan_splitk(N, H, W, Cin, Cout, &block_n, &csize), "conv_splitk: unsupported shape N %d H %d W %d Cin %d Cout %d", N,
             H, W, Cin, Cout);
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.bw = W; p.bh = H; p.bn = 128 / (W * H);
  p.tiles_w = 1; p.tiles_h = 1; p.tiles_n = (N + p.bn - 1) / p.bn;
  p.block_n = block_n; p.n_blocks = Cout / block_n;
  p.kc = 64; p.k_chunks = Cin / 64; p.k_blocks = p.k_chunks * 9;
  p.csize = csize;
  p.rows_per_cta = 128 / csize;
  p.cols_per_thread = block_n / csize;
  p.a_bytes = 128u * 128u;
  p.b_bytes = (uint32_t)block_n * 128u;
  p.stage_bytes = (p.a_bytes + p.b_bytes + 1023u) & ~1023u;
  p.part_stride = (uint32_t)block_n + 4u;
  const size_t part_bytes = (size_t)128 * p.part_stride * sizeof(float);
  const uint32_t aux_bytes = 8 * (2 * kMaxStages + 1) + 64;
  const size_t budget = 227u * 1024u - 1024u - aux_bytes;
  int stages = (int)(budget / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  BG_REQUIRE(stages >= 2 && part_bytes <= budget, "conv_splitk: tile does not fit shared memory");
  p.stages = stages;
  p.bias = bias; p.noise = noise; p.noise_w = noise_w;
  p.gate_src = reinterpret_cast<const __nv_bfloat16*>(gate_src);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.act = act; p.slope = slope;

  CUtensorMap tmx, tmw;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {64u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    if (make_tmap_bf16(&tmx, x, 4, dims, str, box, 128) != 0) return 1;
  }
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 9u};
    uint64_t str[2] = {(uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
    uint32_t box[3] = {64u, (uint32_t)block_n, 1u};
    if (make_tmap_bf16(&tmw, wpack, 3, dims, str, box, 128) != 0) return 1;
  }
  const size_t pipe_bytes = (size_t)p.stages * p.stage_bytes;
  const size_t smem_bytes = (pipe_bytes > part_bytes ? pipe_bytes : part_bytes) + aux_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    BG_CHECK_CUDA(cudaFuncSetAttribute(conv_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int tiles = p.tiles_n * p.n_blocks;
  BG_CHECK_CUDA(launch_pdl_cluster(conv_splitk_kernel, tiles * csize, kThreads, smem_bytes, stream, csize, tmx, tmw, p));
  return 0;
}

}  // namespace bg
