// C-ABI plumbing shared by all entry points: error text, launch counter, device queries and the
// driver entry point for TMA descriptor encoding (resolved at run time so the library links
// against cudart only and still loads — without computing — on a GPU-less build box).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <mutex>

#include "common.cuh"

namespace iir {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return IIR_ERR_CUDA;
  }
  return IIR_OK;
}

int sm_count() {
  // per device: a process may drive several GPUs (one per rank is the norm, but nothing forbids more)
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 148;  // B200
  }
  const int slot = dev >= 0 && dev < 64 ? dev : 0;
  int n = cache[slot].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = 148;
    }
    cache[slot].store(n, std::memory_order_relaxed);
  }
  return n;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("IIR_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    else
      cudaGetLastError();
  });
  return fn;
}

CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, uint32_t rank, const void* base,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = resolve_encode();
  if (!fn) return CUDA_ERROR_NOT_INITIALIZED;
  cuuint64_t d[5];
  cuuint64_t s[4];
  cuuint32_t b[5];
  cuuint32_t es[5];
  for (uint32_t i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) s[i] = strides_bytes[i];
  }
  return fn(map, dt, rank, const_cast<void*>(base), d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace iir

extern "C" int iir_abi_version(void) { return IIR_ABI_VERSION; }
extern "C" int iir_h16_dtype(void) { return IIR_H16; }
extern "C" const char* iir_last_error(void) { return iir::g_err; }
extern "C" uint64_t iir_launch_count(void) { return iir::g_launches.load(std::memory_order_relaxed); }

extern "C" int iir_memset_zero(void* ptr, int64_t bytes, void* stream) {
  IIR_REQUIRE(ptr && bytes > 0, "iir_memset_zero: bad args");
  cudaError_t e = cudaMemsetAsync(ptr, 0, static_cast<size_t>(bytes), reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { iir::set_error("iir_memset_zero: %s", cudaGetErrorString(e)); return IIR_ERR_CUDA; }
  return IIR_OK;
}
