// Host-side helpers shared by the tensor-core kernels: cuTensorMapEncodeTiled through the runtime's driver entry
// point (no link-time dependency on libcuda), NHWC tensor-map views, tile-shape selection.
#pragma once
#include <cuda.h>

#include "common.cuh"

typedef CUresult (*DgEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline DgEncodeTiledFn get_encode() {
  static DgEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) {
      dg_set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<DgEncodeTiledFn>(ptr);
  }
  return fn;
}

// bf16 tensor map, zero fill out of bounds.  dims/strides fastest-first; strides in elements; swizzle_bytes 128 or 32.
static inline int make_map(CUtensorMap* m, const void* base, int rank, const long long* dims,
                           const long long* strides_elems, const int* box, int swizzle_bytes = 128) {
  DgEncodeTiledFn enc = get_encode();
  if (!enc) return DG_ERR_CUDA;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = (cuuint64_t)dims[i];
    bdim[i] = (cuuint32_t)box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = (cuuint64_t)strides_elems[i] * 2;
  }
  const CUtensorMapSwizzle sw = swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    dg_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %lld %lld %lld, box %d %d %d)", (int)r,
                 rank, dims[0], dims[1], rank > 2 ? dims[2] : 0, box[0], box[1], rank > 2 ? box[2] : 0);
    return DG_ERR_CUDA;
  }
  return DG_OK;
}

// parity-split view of an NHWC tensor [B,H,W,C]: dims {2C, W/2, 2, H/2, B}
static inline int make_parity_map(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h,
                                  int box_b) {
  long long dims[5] = {2LL * C, W / 2, 2, H / 2, B};
  long long str[5] = {1, 2LL * C, (long long)W * C, 2LL * W * C, (long long)H * W * C};
  int box[5] = {64, box_w, 1, box_h, box_b};
  return make_map(m, base, 5, dims, str, box);
}
// plain view of an NHWC tensor [B,H,W,C]: dims {C, W, H, B}
static inline int make_nhwc_map(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h,
                                int box_b) {
  long long dims[4] = {C, W, H, B};
  long long str[4] = {1, C, (long long)W * C, (long long)H * W * C};
  int box[4] = {64, box_w, box_h, box_b};
  return make_map(m, base, 4, dims, str, box);
}
static inline int make_weight_map(CUtensorMap* m, const void* base, int rows, int cols, int box_rows) {
  long long dims[2] = {cols, rows};
  long long str[2] = {1, cols};
  int box[2] = {64, box_rows};
  return make_map(m, base, 2, dims, str, box);
}

static inline int num_sms() {
  static int n_dev[kMaxDevices] = {};
  const int dev = current_device();
  int& n = n_dev[dev];
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// split `pixels` (128 or 32) over (W, H, B) of a power-of-two image
static inline void tile_shape(int pixels, int H, int W, int* Wt, int* Ht, int* Bt) {
  *Wt = W < pixels ? W : pixels;
  int rest = pixels / *Wt;
  *Ht = H < rest ? H : rest;
  *Bt = rest / *Ht;
}
