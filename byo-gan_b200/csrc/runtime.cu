// Host-side runtime glue shared by all C-ABI entry points: error slot, device attributes,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency,
// so the library still loads on a CPU-only box for the symbol check).
#include "common.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

namespace bg {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_err; }

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return 1;
}

// SMs the persistent / one-wave grids are sized for = SMs of the device minus a reserve (bg_set_sm_reserve, BG_SM_RESERVE):
// in a data-parallel run NCCL's all-reduce CTAs occupy a few SMs while the backward is still running; a 148-CTA
// persistent grid (one CTA per SM, 200+ KB of shared memory each) then cannot be fully resident and its last CTAs run as
// a second wave.  Sizing the grids for 148 - reserve keeps them one wave next to the collective.
static int g_sm_reserve = -1;
int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  if (g_sm_reserve < 0) {
    const char* e = getenv("BG_SM_RESERVE");
    g_sm_reserve = e ? atoi(e) : 0;
    if (g_sm_reserve < 0) g_sm_reserve = 0;
  }
  int n = cached[dev] - g_sm_reserve;
  n &= ~1;                                       // CTA pairs need an even count
  return n < 16 ? 16 : n;
}
void set_sm_reserve(int n) { g_sm_reserve = n < 0 ? 0 : n; }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("BG_PDL");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

// Chain-deterministic mode (bg_set_deterministic / BG_DETERMINISTIC=1): every reduction whose result feeds later layers
// (instance-norm statistics, the AdaIN backward sums, the minibatch-stddev plane) is summed in a fixed order — one
// contributing block per output, ordered two-stage sums inside it — so images, critic scores and activation gradients are
// bit-reproducible from run to run.  Leaf sums (weight-gradient split-K partials, bias / noise-weight gradients, loss
// terms) keep their fp32 atomics: their order noise (~1e-6 relative) ends in that tensor and is never amplified.
static int g_deterministic = -1;
bool deterministic() {
  if (g_deterministic < 0) {
    const char* e = getenv("BG_DETERMINISTIC");
    g_deterministic = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return g_deterministic != 0;
}
void set_deterministic(int on) { g_deterministic = on ? 1 : 0; }

__global__ void zero_kernel(uint4* __restrict__ p16, size_t n16, uint32_t* __restrict__ tail, int ntail) {
  pdl_prologue();
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p16[i] = z;
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0u;
}

// Pre-zeroed range (bg_set_prezeroed_range): the caller promises that every accumulation target inside it is already zero
// (one bulk memset of a slab it carved them from), so the per-call zero fill — one extra launch in the dependent-launch
// chain per accumulator, ~120 per training iteration — is skipped for targets that lie inside the range.
static thread_local const uint8_t* g_prezero_lo = nullptr;
static thread_local const uint8_t* g_prezero_hi = nullptr;
void set_prezeroed_range(const void* base, size_t bytes) {
  g_prezero_lo = static_cast<const uint8_t*>(base);
  g_prezero_hi = bytes ? g_prezero_lo + bytes : g_prezero_lo;
}

int launch_zero(void* ptr, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return 0;
  {
    const uint8_t* b0 = static_cast<const uint8_t*>(ptr);
    if (g_prezero_lo != nullptr && b0 >= g_prezero_lo && b0 + bytes <= g_prezero_hi) return 0;
  }
  if ((bytes & 3) != 0 || (reinterpret_cast<uintptr_t>(ptr) & 3) != 0) {      // not word sized: plain memset node
    BG_CHECK_CUDA(cudaMemsetAsync(ptr, 0, bytes, stream));
    return 0;
  }
  // head words up to 16-byte alignment go with the tail (at most 3 + 3 words)
  uint8_t* b = static_cast<uint8_t*>(ptr);
  const size_t head = (16 - (reinterpret_cast<uintptr_t>(b) & 15)) & 15;
  if (head > 0 && head <= bytes) {
    BG_CHECK_CUDA(cudaMemsetAsync(b, 0, head, stream));                         // never taken for torch allocations
    b += head;
    bytes -= head;
  }
  const size_t n16 = bytes / 16;
  const int ntail = (int)((bytes - n16 * 16) / 4);
  size_t blocks = (n16 + 255) / 256;
  const size_t cap = (size_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  BG_CHECK_CUDA(launch_pdl(zero_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, reinterpret_cast<uint4*>(b), n16,
                           reinterpret_cast<uint32_t*>(b + n16 * 16), ntail));
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  return make_tmap_bf16_strided(map, base, rank, dims, strides_bytes, box, nullptr, swizzle_bytes);
}

// elem_strides (may be null = all 1): traversal stride per dimension; box[] is then the SPAN in elements and the tile
// written to shared memory has ceil(box / stride) elements along that dimension.
int make_tmap_bf16_strided(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                           int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return 1;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu box %u,%u,%u swz %d)",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
              swizzle_bytes);
    return 1;
  }
  return 0;
}

}  // namespace bg

extern "C" const char* bg_last_error(void) { return bg::last_error(); }
extern "C" int bg_abi_version(void) { return 3; }
extern "C" int bg_set_deterministic(int on) {
  bg::set_deterministic(on);
  return 0;
}
extern "C" int bg_get_deterministic(void) { return bg::deterministic() ? 1 : 0; }
extern "C" int bg_set_prezeroed_range(const void* base, size_t bytes) {
  bg::set_prezeroed_range(base, bytes);
  return 0;
}
extern "C" int bg_set_sm_reserve(int sms) {
  bg::set_sm_reserve(sms);
  return 0;
}
