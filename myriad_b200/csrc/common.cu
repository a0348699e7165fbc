#include "common.h"

#include <string.h>

namespace myr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("MYR_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}
void set_pdl(int v) { g_pdl = v ? 1 : 0; }

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static void* g_trace = nullptr;
void set_trace_buffer(void* buf) { g_trace = buf; }
long long* next_trace_slot() {
  long long* t = reinterpret_cast<long long*>(g_trace);
  if (g_trace) g_trace = reinterpret_cast<char*>(g_trace) + 148 * 6 * sizeof(long long);
  return t;
}

int make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return MYR_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): base=%p rank=%d dims=[%llu,%llu,%llu] stride0=%llu box=[%u,%u,%u]",
              (int)r, base, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
              box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    return MYR_ERR_CUDA;
  }
  return MYR_OK;
}

}  // namespace myr

extern "C" {

int myr_version(void) { return 1; }

/* sizeof() of the argument structs as this library was compiled: bindings check their own layout against it */
size_t myr_abi_sizeof(int32_t which) {
  switch (which) {
    case 0: return sizeof(myr_gemm_args);
    case 1: return sizeof(myr_attn_args);
    case 2: return sizeof(myr_norm_args);
    case 3: return sizeof(myr_rope_args);
    case 4: return sizeof(myr_decode_attn_args);
    case 5: return sizeof(myr_mega_op);
    default: return 0;
  }
}

int myr_last_error(char* buf, size_t n) {
  if (!buf || n == 0) return MYR_ERR_INVALID;
  strncpy(buf, myr::g_err, n - 1);
  buf[n - 1] = 0;
  return MYR_OK;
}

void myr_set_pdl(int32_t enabled) { myr::set_pdl(enabled); }

int myr_device_sm_count(void) { return myr::sm_count(); }

unsigned long long myr_launch_count(void) { return __atomic_load_n(&myr::g_launches, __ATOMIC_RELAXED); }
}
