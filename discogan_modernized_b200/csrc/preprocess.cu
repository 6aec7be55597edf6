// Input pipeline on the GPU: the per-image work of the reference's read_images / DiscoGANDataset._load_and_process_image
// (dataset.py:37-73, 238-261) -- crop a half of a side-by-side pair, erode the edge map, bilinear resize, scale to [0,1],
// HWC -> CHW -- done by one kernel per batch on decoded uint8 images that already live in device memory.
//
// Arithmetic is that of the reference's cv2 calls, bit for bit:
//   domain None / 'B' (uint8 image into cv2.resize, INTER_LINEAR): OpenCV's fixed-point bilinear -- 11-bit coefficients
//       saturate_cast<short>(w * 2048) (round half to even), horizontal pass in int32, vertical pass
//       ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2, then float32(v) / 255.f
//   domain 'A' (255 - image, cv2.dilate 3x3, 255 - image: an erosion over the in-bounds 3x3 neighbourhood, carried out in
//       float64 by the reference, then cv2.resize on float64): float coefficients, double multiply-adds WITHOUT fma
//       contraction (numpy / OpenCV round after every operation), float32(v) / 255.f
// Source coordinates: fx = float((dx + 0.5) * scale - 0.5), scale = 1 / (double(dst) / src); sx = floor(fx);
// left / right border: sx < 0 -> (0, w=0); sx >= W-1 -> (W-1, w=0); rows are clamped to [0, H-1] with the weights kept.
// This is HBM-bound byte work: each output pixel reads 4 (x9 for the erosion) source pixels; no tensor cores involved.
#include "common.cuh"

namespace {

struct PreImage {            // one row of the device table (8 x int64)
  long long src;             // uint8 [H][W][3] (device pointer)
  long long H, W;            // full decoded size
  long long x0, cw;          // crop: columns [x0, x0 + cw)
  long long mode;            // 0 = resize only, 1 = erode 3x3 then resize in float64 (domain 'A')
  long long pad0, pad1;
};

__device__ __forceinline__ void src_coord(int d, double scale, int n, int* s, float* f, bool clamp_w) {
  float v = (float)(((double)d + 0.5) * scale - 0.5);
  int i = (int)floorf(v);
  v -= (float)i;
  if (clamp_w) {
    if (i < 0) { v = 0.f; i = 0; }
    if (i >= n - 1) { v = 0.f; i = n - 1; }
  }
  *s = i;
  *f = v;
}

__device__ __forceinline__ int rnd_short(float w) {   // saturate_cast<short>(w * 2048): cvRound = round half to even
  int r = __float2int_rn(w * 2048.f);
  return r > 32767 ? 32767 : (r < -32768 ? -32768 : r);
}

__device__ __forceinline__ int eroded(const unsigned char* img, int H, int W, long long stride, int y, int x, int c) {
  int m = 255;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= W) continue;
      const int v = img[(long long)yy * stride + (long long)xx * 3 + c];
      m = v < m ? v : m;
    }
  }
  return m;
}

__global__ void preprocess_u8_kernel(const PreImage* __restrict__ table, int S, float* __restrict__ out) {
  const PreImage im = table[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * S) return;
  const int dy = idx / S, dx = idx - dy * S;
  const int H = (int)im.H, W = (int)im.cw;
  const long long stride = im.W * 3;
  const unsigned char* img = reinterpret_cast<const unsigned char*>(im.src) + im.x0 * 3;
  const double scale_x = 1.0 / ((double)S / (double)W), scale_y = 1.0 / ((double)S / (double)H);
  int sx, sy;
  float fx, fy;
  src_coord(dx, scale_x, W, &sx, &fx, true);
  src_coord(dy, scale_y, H, &sy, &fy, false);
  const int x1 = sx + 1 < W ? sx + 1 : W - 1;
  const int r0 = sy < 0 ? 0 : (sy > H - 1 ? H - 1 : sy);
  const int r1 = sy + 1 < 0 ? 0 : (sy + 1 > H - 1 ? H - 1 : sy + 1);
  float* o = out + (size_t)blockIdx.y * 3 * S * S + idx;
  if (im.mode == 0) {
    const int a0 = rnd_short(1.f - fx), a1 = rnd_short(fx), b0 = rnd_short(1.f - fy), b1 = rnd_short(fy);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int p00 = img[(long long)r0 * stride + sx * 3 + c], p01 = img[(long long)r0 * stride + x1 * 3 + c];
      const int p10 = img[(long long)r1 * stride + sx * 3 + c], p11 = img[(long long)r1 * stride + x1 * 3 + c];
      const int h0 = p00 * a0 + p01 * a1, h1 = p10 * a0 + p11 * a1;
      const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      const int u = v < 0 ? 0 : (v > 255 ? 255 : v);
      o[(size_t)c * S * S] = __fdiv_rn((float)u, 255.f);
    }
  } else {
    const double a0 = (double)(1.f - fx), a1 = (double)fx, b0 = (double)(1.f - fy), b1 = (double)fy;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double p00 = eroded(img, H, W, stride, r0, sx, c), p01 = eroded(img, H, W, stride, r0, x1, c);
      const double p10 = eroded(img, H, W, stride, r1, sx, c), p11 = eroded(img, H, W, stride, r1, x1, c);
      const double h0 = __dadd_rn(__dmul_rn(p00, a0), __dmul_rn(p01, a1));
      const double h1 = __dadd_rn(__dmul_rn(p10, a0), __dmul_rn(p11, a1));
      const double v = __dadd_rn(__dmul_rn(h0, b0), __dmul_rn(h1, b1));
      o[(size_t)c * S * S] = __fdiv_rn((float)v, 255.f);
    }
  }
}

}  // namespace

extern "C" {

int dg_preprocess_u8(const long long* table, int n, int S, float* out, cudaStream_t stream) {
  DG_CHECK_ARG(table != nullptr && out != nullptr && n >= 0 && S > 0, "preprocess: bad arguments (n=%d S=%d)", n, S);
  if (n == 0) return DG_OK;
  DG_CHECK_ARG(((uintptr_t)table & 7) == 0, "preprocess: table must be 8-byte aligned");
  dim3 grid(dg_ceil_div((long long)S * S, 256), n);
  dg_launch(preprocess_u8_kernel, dg_cfg(grid, 256, 0, stream), reinterpret_cast<const PreImage*>(table), S, out);
  DG_CHECK_LAUNCH("preprocess_u8_kernel");
  return DG_OK;
}

}  // extern "C"
