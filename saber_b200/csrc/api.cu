// saber_b200 — C-ABI plumbing: per-thread error string, version, driver entry point for TMA
// tensor-map encoding (fetched at run time so the library links against cudart only and loads on a
// box without a driver — needed by the CPU-side "library loads and exports every symbol" test).
#include "common.cuh"
#include <stdarg.h>

namespace {
thread_local char g_err[1024] = "";

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
}  // namespace

void sb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* sb_last_error(void) { return g_err; }

extern "C" int sb_version(void) { return 100; }

// Number of SMs of the current device (148 on B200); negative on error.
extern "C" int sb_device_sm_count(void) {
  int dev = 0, n = 0;
  SB_CHECK_CUDA(cudaGetDevice(&dev));
  SB_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return n;
}

// Fails loudly unless the current device is compute capability 10.x (the kernels are sm_100a-only).
extern "C" int sb_require_sm100(void) {
  int dev = 0, major = 0, minor = 0;
  SB_CHECK_CUDA(cudaGetDevice(&dev));
  SB_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  SB_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    sb_set_error("saber_b200 requires an sm_100a device (B200); found sm_%d%d", major, minor);
    return SB_ERR_UNSUPPORTED;
  }
  return SB_OK;
}

static int sb_resolve_encode() {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      sb_set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s",
                   cudaGetErrorString(e));
      return SB_ERR_DRIVER;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  return SB_OK;
}

int sb_make_tmap_nd_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  int rc = sb_resolve_encode();
  if (rc != SB_OK) return rc;
  if (rank < 1 || rank > 5) {
    sb_set_error("sb_make_tmap_nd_bf16: bad rank %d", rank);
    return SB_ERR_ARG;
  }
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstride[i] = strides_bytes[i];
  }
  const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                        gdim, gstride, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sb_set_error("cuTensorMapEncodeTiled (rank %d, swizzle %d) failed (%d): base=%p dims=%llu,%llu,.. box=%u,%u,..",
                 rank, swizzle_bytes, (int)r, base, (unsigned long long)dims[0],
                 (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
    return SB_ERR_DRIVER;
  }
  return SB_OK;
}

// 2-D row-major tensor map of 2-byte (bf16) or 4-byte (fp32) elements with a chosen swizzle (0 / 32 / 64 / 128 bytes): the
// TMA-store side of the GEMM epilogues (box = one warp's 32 x 32 staging tile).
int sb_make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                    uint32_t box_rows, uint32_t box_cols, int swizzle_bytes) {
  {
    int rc = sb_resolve_encode();
    if (rc != SB_OK) return rc;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * static_cast<uint64_t>(elem_bytes)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sb_set_error("cuTensorMapEncodeTiled failed (%d): base=%p elem=%d rows=%llu cols=%llu ld=%llu box=%ux%u swizzle=%d",
                 (int)r, base, elem_bytes, (unsigned long long)rows, (unsigned long long)cols,
                 (unsigned long long)ld_elems, box_rows, box_cols, swizzle_bytes);
    return SB_ERR_DRIVER;
  }
  return SB_OK;
}

int sb_make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                         uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols) {
  {
    int rc = sb_resolve_encode();
    if (rc != SB_OK) return rc;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                        gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sb_set_error("cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu ld=%llu box=%ux%u",
                 (int)r, base, (unsigned long long)rows, (unsigned long long)cols,
                 (unsigned long long)ld_elems, box_rows, box_cols);
    return SB_ERR_DRIVER;
  }
  return SB_OK;
}
