// Direct (non-GEMM) kernels, sm_100a:
//   * the 3-channel image-side layers: Conv2d(3,64,4,2,1)+LeakyReLU (model.py:8-9,80-81) and
//     ConvTranspose2d(64,3,4,2,1)+Sigmoid (model.py:142-143) -- K=48 / N=3, HBM/latency bound, so they read and
//     write the fp32 NCHW images of the module boundary directly and fuse the activation;
//   * the 4x4 "valid" heads that collapse a 4x4 map to 1x1 (model.py:35,107) and the ConvTranspose2d(100,C,4,1,0)
//     that expands it again (model.py:114): skinny matrix products against the packed weight Wd[Ns][16*C];
//   * slow SIMT reference versions of the tensor-core convolutions, for on-device debugging only
//     (DISCOGAN_B200_CONV=simt); never the product path.
#include "common.cuh"

namespace {

constexpr int kC1 = 64;  // channel count next to the image in every member of the model family

// ------------------------------------------------------------------------------------------------
// 3 -> 64 stride-2 4x4 "down" direct conv.  img is fp32 NCHW [B,3,S,S]; out is bf16 NHWC [B,S/2,S/2,64].
//   SIGGRAD = false: conv1 forward, out = LeakyReLU(conv(img))
//   SIGGRAD = true : dgrad of the final ConvTranspose2d+Sigmoid: the "image" is dout * yimg * (1 - yimg), no act.
// w is the fp32 PyTorch weight [64][3][4][4] in both cases (conv: [co][ci], convT: [ci][co] -- same memory order
// seen from the 64-channel side).
// ------------------------------------------------------------------------------------------------
template <bool SIGGRAD>
__global__ void __launch_bounds__(256)
c3_down_kernel(const float* __restrict__ img, const float* __restrict__ yimg, const float* __restrict__ w,
               bf16* __restrict__ out, int B, int S, int act, float slope) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ float ws[48][kC1];  // [k = c*16 + kh*4 + kw][co]
  for (int i = threadIdx.x; i < 48 * kC1; i += blockDim.x) {
    const int co = i / 48, k = i % 48;
    ws[k][co] = w[i];
  }
  __syncthreads();
  const int So = S >> 1;
  const long long npix = (long long)B * So * So;
  // 4 threads per pixel, 16 channels each; a warp covers 8 pixels x 64 channels = 8 x 128 B contiguous
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pix = gid >> 2;
  const int cg = (int)(gid & 3) * 16;
  if (pix >= npix) return;
  const int wo = (int)(pix % So);
  const int ho = (int)((pix / So) % So);
  const int b = (int)(pix / ((long long)So * So));
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  const size_t plane = (size_t)S * S;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const int y = 2 * ho - 1 + kh;
      if (y < 0 || y >= S) continue;
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        const int x = 2 * wo - 1 + kw;
        if (x < 0 || x >= S) continue;
        const size_t off = ((size_t)b * 3 + c) * plane + (size_t)y * S + x;
        float v = img[off];
        if (SIGGRAD) {
          const float s = yimg[off];
          v *= s * (1.f - s);
        }
        const float4* wr = reinterpret_cast<const float4*>(&ws[c * 16 + kh * 4 + kw][cg]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = wr[i];
          acc[4 * i] += v * t.x;
          acc[4 * i + 1] += v * t.y;
          acc[4 * i + 2] += v * t.z;
          acc[4 * i + 3] += v * t.w;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = act_fwd(acc[i], act, slope);
  bf16* o = out + pix * kC1 + cg;
  *reinterpret_cast<bf16x8*>(o) = pack8(acc);
  *reinterpret_cast<bf16x8*>(o + 8) = pack8(acc + 8);
}

// ------------------------------------------------------------------------------------------------
// 64 -> 3 stride-2 4x4 "up" direct conv.  act64 is bf16 NHWC [B,S/2,S/2,64]; img is fp32 NCHW [B,3,S,S].
//   MASK = false: final ConvTranspose2d forward, img = sigmoid(convT(act64))      (sigmoid != 0)
//   MASK = true : dgrad of conv1+LeakyReLU: input is dy64 * lrelu'(y64), img = d(loss)/d(input image)
// One thread per output pixel; only the 2x2 taps of the pixel's parity contribute.
// ------------------------------------------------------------------------------------------------
template <bool MASK>
__global__ void __launch_bounds__(256)
c3_up_kernel(const bf16* __restrict__ act64, const bf16* __restrict__ y64, const float* __restrict__ w,
             float* __restrict__ img, int B, int S, int sigmoid, float slope, int accumulate) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ float ws[16][kC1][3];  // [kh*4+kw][c64][c3]
  for (int i = threadIdx.x; i < 48 * kC1; i += blockDim.x) {
    const int c64 = i / 48, r = i % 48, c3 = r / 16, t = r % 16;
    ws[t][c64][c3] = w[i];
  }
  __syncthreads();
  const long long npix = (long long)B * S * S;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int x = (int)(pix % S);
  const int y = (int)((pix / S) % S);
  const int b = (int)(pix / ((long long)S * S));
  const int So = S >> 1;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
  for (int th = 0; th < 2; ++th) {
    const int kh = ((y + 1) & 1) + 2 * th;  // kh == y+1 (mod 2)
    const int i = (y + 1 - kh) >> 1;
    if (i < 0 || i >= So) continue;
#pragma unroll
    for (int tw = 0; tw < 2; ++tw) {
      const int kw = ((x + 1) & 1) + 2 * tw;
      const int j = (x + 1 - kw) >> 1;
      if (j < 0 || j >= So) continue;
      const size_t off = (((size_t)b * So + i) * So + j) * kC1;
      const int t = kh * 4 + kw;
#pragma unroll
      for (int c8 = 0; c8 < kC1; c8 += 8) {
        float f[8];
        unpack8(*reinterpret_cast<const bf16x8*>(act64 + off + c8), f);
        if (MASK) {
          float fy[8];
          unpack8(*reinterpret_cast<const bf16x8*>(y64 + off + c8), fy);
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] *= (fy[k] > 0.f ? 1.f : slope);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          a0 += f[k] * ws[t][c8 + k][0];
          a1 += f[k] * ws[t][c8 + k][1];
          a2 += f[k] * ws[t][c8 + k][2];
        }
      }
    }
  }
  if (sigmoid) {
    a0 = 1.f / (1.f + expf(-a0));
    a1 = 1.f / (1.f + expf(-a1));
    a2 = 1.f / (1.f + expf(-a2));
  }
  const size_t plane = (size_t)S * S;
  float* o = img + (size_t)b * 3 * plane + (size_t)y * S + x;
  if (accumulate) {
    o[0] += a0;
    o[plane] += a1;
    o[2 * plane] += a2;
  } else {
    o[0] = a0;
    o[plane] = a1;
    o[2 * plane] = a2;
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad of the two 3-channel layers: dw[c64][c3][kh][kw] += sum_pixels v64[pix][c64] * patch[pix][c3,kh,kw]
//   v64   = bf16 NHWC tensor, optionally masked by lrelu'(y64)      (conv1: dy of conv1 output; convT: its input)
//   patch = 4x4x3 window of the fp32 NCHW image, optionally * yimg*(1-yimg)  (conv1: input image; convT: dout)
// Each block stages 32 pixels in smem, thread (c64, kq) keeps 12 of the 48 patch entries; atomics at the end.
// ------------------------------------------------------------------------------------------------
template <bool MASK, bool SIGGRAD>
__global__ void __launch_bounds__(256)
c3_wgrad_kernel(const bf16* __restrict__ v64, const bf16* __restrict__ y64, const float* __restrict__ img,
                const float* __restrict__ yimg, float* __restrict__ dw, int B, int S, float slope,
                int pix_per_block) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ float sv[32][kC1];
  __shared__ float sp[32][48];
  const int So = S >> 1;
  const long long npix = (long long)B * So * So;
  const int c64 = threadIdx.x & 63;
  const int kq = threadIdx.x >> 6;  // 0..3 -> k in [12*kq, 12*kq+12)
  float acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.f;
  const long long p_begin = (long long)blockIdx.x * pix_per_block;
  const long long p_end = min(npix, p_begin + pix_per_block);
  const size_t plane = (size_t)S * S;
  for (long long p0 = p_begin; p0 < p_end; p0 += 32) {
    // stage v64 (32 x 64) and the patches (32 x 48)
    for (int i = threadIdx.x; i < 32 * kC1; i += blockDim.x) {
      const int pl = i >> 6, c = i & 63;
      const long long pix = p0 + pl;
      float v = 0.f;
      if (pix < p_end) {
        v = __bfloat162float(v64[pix * kC1 + c]);
        if (MASK) v *= (__bfloat162float(y64[pix * kC1 + c]) > 0.f ? 1.f : slope);
      }
      sv[pl][c] = v;
    }
    for (int i = threadIdx.x; i < 32 * 48; i += blockDim.x) {
      const int pl = i / 48, k = i % 48;
      const long long pix = p0 + pl;
      float v = 0.f;
      if (pix < p_end) {
        const int wo = (int)(pix % So);
        const int ho = (int)((pix / So) % So);
        const int b = (int)(pix / ((long long)So * So));
        const int c = k >> 4, kh = (k >> 2) & 3, kw = k & 3;
        const int y = 2 * ho - 1 + kh, x = 2 * wo - 1 + kw;
        if (y >= 0 && y < S && x >= 0 && x < S) {
          const size_t off = ((size_t)b * 3 + c) * plane + (size_t)y * S + x;
          v = img[off];
          if (SIGGRAD) {
            const float s = yimg[off];
            v *= s * (1.f - s);
          }
        }
      }
      sp[pl][k] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int pl = 0; pl < 32; ++pl) {
      const float v = sv[pl][c64];
#pragma unroll
      for (int i = 0; i < 12; ++i) acc[i] += v * sp[pl][12 * kq + i];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) atomicAdd(dw + c64 * 48 + 12 * kq + i, acc[i]);
}

// ------------------------------------------------------------------------------------------------
// FC heads against the packed weight Wd[Ns][K] (K = 16*C), "small" side [B][Ns], "big" side [B][K] bf16.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float ld_small(const T* p);
template <>
__device__ __forceinline__ float ld_small<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_small<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void st_small(T* p, float v);
template <>
__device__ __forceinline__ void st_small<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st_small<bf16>(bf16* p, float v) { *p = __float2bfloat16(v); }

// small[b][n] = sum_k big[b][k] * wd[n][k].  Block = (4 n) x (8 b) outputs, 256 threads split K.
template <typename TS>
__global__ void __launch_bounds__(256)
fc_down_kernel(const bf16* __restrict__ big, const bf16* __restrict__ wd, TS* __restrict__ small, int B, int Ns,
               int K) {
  griddep_launch_dependents();
  griddep_wait();
  const int n0 = blockIdx.x * 4, b0 = blockIdx.y * 8;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int k = threadIdx.x * 8; k < K; k += 256 * 8) {
    float wv[4][8], xv[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (n0 + i < Ns) {
        unpack8(*reinterpret_cast<const bf16x8*>(wd + (size_t)(n0 + i) * K + k), wv[i]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) wv[i][e] = 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (b0 + j < B) {
        unpack8(*reinterpret_cast<const bf16x8*>(big + (size_t)(b0 + j) * K + k), xv);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[i][j] += wv[i][e] * xv[e];
      }
    }
  }
  __shared__ float sh[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = warp_sum(acc[i][j]);
      if (lane == 0) sh[warp][i * 8 + j] = v;
    }
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = 0.f;
#pragma unroll
    for (int wv_ = 0; wv_ < 8; ++wv_) v += sh[wv_][threadIdx.x];
    const int i = threadIdx.x >> 3, j = threadIdx.x & 7;
    if (n0 + i < Ns && b0 + j < B) st_small<TS>(small + (size_t)(b0 + j) * Ns + n0 + i, v);
  }
}

// big[b][k] = sum_n small[b][n] * wd[n][k].  Block = 64 consecutive k (8 threads x 8) x 32 batch rows; it stages its
// [Ns][64] weight slice and its [32][Ns] slice of `small` in shared memory once, so every operand byte comes from L2 once
// per (k slice, batch group) -- the previous version re-read the weight slice for every 4 batch rows (26 MB of L2 reads
// for a 2.6 MB problem at 64x64, 12 us per launch inside the step graph).  Same summation order (n ascending).
template <typename TS>
__global__ void __launch_bounds__(256)
fc_up_kernel(const TS* __restrict__ small, const bf16* __restrict__ wd, bf16* __restrict__ big, int B, int Ns, int K) {
  griddep_launch_dependents();
  griddep_wait();
  extern __shared__ __align__(16) uint8_t fsm[];
  bf16* wsm = reinterpret_cast<bf16*>(fsm);                          // [Ns][64]
  float* ssm = reinterpret_cast<float*>(fsm + (size_t)Ns * 128);     // [32][Ns]
  const int k0 = blockIdx.x * 64, b0 = blockIdx.y * 32;
  for (int i = threadIdx.x; i < Ns * 8; i += 256) {
    const int n = i >> 3, v = i & 7;
    const int kk = k0 + v * 8;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (kk < K) val = *reinterpret_cast<const uint4*>(wd + (size_t)n * K + kk);
    *reinterpret_cast<uint4*>(wsm + n * 64 + v * 8) = val;
  }
  for (int i = threadIdx.x; i < 32 * Ns; i += 256) {
    const int j = i / Ns, n = i - j * Ns;
    ssm[i] = (b0 + j < B) ? ld_small<TS>(small + (size_t)(b0 + j) * Ns + n) : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int k = k0 + tx * 8, b = b0 + ty;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const float* srow = ssm + ty * Ns;
#pragma unroll 4
  for (int n = 0; n < Ns; ++n) {
    float wv[8];
    unpack8(*reinterpret_cast<const bf16x8*>(wsm + n * 64 + tx * 8), wv);
    const float sv = srow[n];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += sv * wv[e];
  }
  if (b < B && k < K) *reinterpret_cast<bf16x8*>(big + (size_t)b * K + k) = pack8(acc);
}

// dw[n][c][tap] = beta*dw + sum_b small[b][n] * big[b][tap*C + c]   (PyTorch layout [Ns][C][4][4], fp32)
// Thread = (n, tap, 8 consecutive c): coalesced 16-byte loads of `big`, 2*C threads per n.
template <typename TS>
__global__ void __launch_bounds__(128)
fc_wgrad_kernel(const TS* __restrict__ small, const bf16* __restrict__ big, float* __restrict__ dw, float beta, int B,
                int Ns, int C) {
  griddep_launch_dependents();
  griddep_wait();
  const int n = blockIdx.y;
  const int cv = C >> 3;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int tap = idx / cv;
  const int c0 = (idx - tap * cv) * 8;
  if (tap >= 16) return;
  const size_t K = (size_t)16 * C;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const bf16* bp = big + (size_t)tap * C + c0;
#pragma unroll 8
  for (int b = 0; b < B; ++b) {
    const float s = ld_small<TS>(small + (size_t)b * Ns + n);
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(bp + (size_t)b * K), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += s * f[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float* d = dw + ((size_t)n * C + c0 + e) * 16 + tap;
    *d = (beta != 0.f ? beta * *d : 0.f) + acc[e];
  }
}

// ------------------------------------------------------------------------------------------------
// SIMT reference convolutions (debug only): same tensors / packed weights as the tensor-core kernels.
// ------------------------------------------------------------------------------------------------
__global__ void simt_down_kernel(const bf16* __restrict__ big, const bf16* __restrict__ wd, bf16* __restrict__ small,
                                 int B, int Hs, int Ws, int Cs, int Cb) {
  griddep_launch_dependents();
  griddep_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * Hs * Ws * Cs;
  if (idx >= total) return;
  const int cs = (int)(idx % Cs);
  long long r = idx / Cs;
  const int wo = (int)(r % Ws);
  r /= Ws;
  const int ho = (int)(r % Hs);
  const int b = (int)(r / Hs);
  const int H = 2 * Hs, W = 2 * Ws;
  float acc = 0.f;
  for (int kh = 0; kh < 4; ++kh) {
    const int y = 2 * ho - 1 + kh;
    if (y < 0 || y >= H) continue;
    for (int kw = 0; kw < 4; ++kw) {
      const int x = 2 * wo - 1 + kw;
      if (x < 0 || x >= W) continue;
      const bf16* xp = big + (((size_t)b * H + y) * W + x) * Cb;
      const bf16* wp = wd + ((size_t)cs * 16 + kh * 4 + kw) * Cb;
      for (int cb = 0; cb < Cb; ++cb) acc += __bfloat162float(xp[cb]) * __bfloat162float(wp[cb]);
    }
  }
  small[idx] = __float2bfloat16(acc);
}
__global__ void simt_up_kernel(const bf16* __restrict__ small, const bf16* __restrict__ wu, bf16* __restrict__ big,
                               int B, int Hs, int Ws, int Cs, int Cb) {
  griddep_launch_dependents();
  griddep_wait();
  const int H = 2 * Hs, W = 2 * Ws;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W * Cb;
  if (idx >= total) return;
  const int cb = (int)(idx % Cb);
  long long r = idx / Cb;
  const int x = (int)(r % W);
  r /= W;
  const int y = (int)(r % H);
  const int b = (int)(r / H);
  float acc = 0.f;
  for (int kh = 0; kh < 4; ++kh) {
    const int t = y + 1 - kh;
    if (t < 0 || (t & 1)) continue;
    const int i = t >> 1;
    if (i >= Hs) continue;
    for (int kw = 0; kw < 4; ++kw) {
      const int u = x + 1 - kw;
      if (u < 0 || (u & 1)) continue;
      const int j = u >> 1;
      if (j >= Ws) continue;
      const bf16* sp = small + (((size_t)b * Hs + i) * Ws + j) * Cs;
      const bf16* wp = wu + ((size_t)cb * 16 + kh * 4 + kw) * Cs;
      for (int cs = 0; cs < Cs; ++cs) acc += __bfloat162float(sp[cs]) * __bfloat162float(wp[cs]);
    }
  }
  big[idx] = __float2bfloat16(acc);
}
__global__ void simt_wgrad_kernel(const bf16* __restrict__ small, const bf16* __restrict__ big, float* __restrict__ dw,
                                  float beta, int B, int Hs, int Ws, int Cs, int Cb) {
  griddep_launch_dependents();
  griddep_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)Cs * Cb * 16;
  if (idx >= total) return;
  const int tap = (int)(idx & 15);
  const int cb = (int)((idx >> 4) % Cb);
  const int cs = (int)((idx >> 4) / Cb);
  const int kh = tap >> 2, kw = tap & 3;
  const int H = 2 * Hs, W = 2 * Ws;
  float acc = 0.f;
  for (int b = 0; b < B; ++b)
    for (int ho = 0; ho < Hs; ++ho) {
      const int y = 2 * ho - 1 + kh;
      if (y < 0 || y >= H) continue;
      for (int wo = 0; wo < Ws; ++wo) {
        const int x = 2 * wo - 1 + kw;
        if (x < 0 || x >= W) continue;
        acc += __bfloat162float(small[(((size_t)b * Hs + ho) * Ws + wo) * Cs + cs]) *
               __bfloat162float(big[(((size_t)b * H + y) * W + x) * Cb + cb]);
      }
    }
  dw[idx] = (beta != 0.f ? beta * dw[idx] : 0.f) + acc;
}

}  // namespace

extern "C" {

int dg_conv_c3_in_fwd(const float* x, const float* w, void* y, int B, int S, float slope, cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && S >= 4 && S % 2 == 0 && x && w && y, "conv_c3_in_fwd: bad args");
  const long long threads = (long long)B * (S / 2) * (S / 2) * 4;
  dg_launch(c3_down_kernel<false>, dg_cfg(dg_ceil_div(threads, 256), 256, 0, stream), x, nullptr, w, (bf16*)y, B, S, DG_ACT_LRELU,
                                                                       slope);
  DG_CHECK_LAUNCH("conv_c3_in_fwd");
  return DG_OK;
}

// Backward of y = LeakyReLU(conv(x, w)): dx (fp32 NCHW, optional, overwritten or accumulated) and dw (+=, atomics).
int dg_conv_c3_in_bwd(const float* x, const float* w, const void* y, const void* dy, float* dx, int dx_accumulate,
                      float* dw, int B, int S, float slope, cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && S >= 4 && S % 2 == 0 && w && y && dy, "conv_c3_in_bwd: bad args");
  if (dx) {
    const long long threads = (long long)B * S * S;
    dg_launch(c3_up_kernel<true>, dg_cfg(dg_ceil_div(threads, 256), 256, 0, stream), (const bf16*)dy, (const bf16*)y, w, dx, B, S, 0,
                                                                      slope, dx_accumulate);
    DG_CHECK_LAUNCH("conv_c3_in_dgrad");
  }
  if (dw) {
    DG_CHECK_ARG(x != nullptr, "conv_c3_in_bwd: x required for wgrad");
    const long long npix = (long long)B * (S / 2) * (S / 2);
    int ppb = (int)((npix + 591) / 592);
    ppb = (ppb + 31) / 32 * 32;
    dg_launch(c3_wgrad_kernel<true, false>, dg_cfg(dg_ceil_div(npix, ppb), 256, 0, stream), (const bf16*)dy, (const bf16*)y, x,
                                                                             nullptr, dw, B, S, slope, ppb);
    DG_CHECK_LAUNCH("conv_c3_in_wgrad");
  }
  return DG_OK;
}

int dg_convT_c3_out_fwd(const void* x, const float* w, float* y, int B, int S, cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && S >= 4 && S % 2 == 0 && x && w && y, "convT_c3_out_fwd: bad args");
  const long long threads = (long long)B * S * S;
  dg_launch(c3_up_kernel<false>, dg_cfg(dg_ceil_div(threads, 256), 256, 0, stream), (const bf16*)x, nullptr, w, y, B, S, 1, 0.f, 0);
  DG_CHECK_LAUNCH("convT_c3_out_fwd");
  return DG_OK;
}

// Backward of y = sigmoid(convT(x, w)): dx (bf16 NHWC, optional) and dw (+=, atomics); dy is d(loss)/dy (fp32 NCHW).
int dg_convT_c3_out_bwd(const void* x, const float* w, const float* y, const float* dy, void* dx, float* dw, int B,
                        int S, cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && S >= 4 && S % 2 == 0 && w && y && dy, "convT_c3_out_bwd: bad args");
  if (dx) {
    const long long threads = (long long)B * (S / 2) * (S / 2) * 4;
    dg_launch(c3_down_kernel<true>, dg_cfg(dg_ceil_div(threads, 256), 256, 0, stream), dy, y, w, (bf16*)dx, B, S, DG_ACT_NONE, 0.f);
    DG_CHECK_LAUNCH("convT_c3_out_dgrad");
  }
  if (dw) {
    DG_CHECK_ARG(x != nullptr, "convT_c3_out_bwd: x required for wgrad");
    const long long npix = (long long)B * (S / 2) * (S / 2);
    int ppb = (int)((npix + 591) / 592);
    ppb = (ppb + 31) / 32 * 32;
    dg_launch(c3_wgrad_kernel<false, true>, dg_cfg(dg_ceil_div(npix, ppb), 256, 0, stream), (const bf16*)x, nullptr, dy, y, dw, B, S,
                                                                             0.f, ppb);
    DG_CHECK_LAUNCH("convT_c3_out_wgrad");
  }
  return DG_OK;
}

// FC heads.  small_f32: the [B][Ns] side is fp32 (Discriminator logit) instead of bf16.
int dg_fc_down(const void* big, const void* wd, void* small, int small_f32, int B, int Ns, int K, cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && Ns > 0 && K > 0 && K % 8 == 0, "fc_down: bad dims");
  dim3 grid(dg_ceil_div(Ns, 4), dg_ceil_div(B, 8));
  if (small_f32)
    dg_launch(fc_down_kernel<float>, dg_cfg(grid, 256, 0, stream), (const bf16*)big, (const bf16*)wd, (float*)small, B, Ns, K);
  else
    dg_launch(fc_down_kernel<bf16>, dg_cfg(grid, 256, 0, stream), (const bf16*)big, (const bf16*)wd, (bf16*)small, B, Ns, K);
  DG_CHECK_LAUNCH("fc_down");
  return DG_OK;
}
int dg_fc_up(const void* small, int small_f32, const void* wd, void* big, int B, int Ns, int K, cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && Ns > 0 && K > 0 && K % 8 == 0, "fc_up: bad dims");
  DG_CHECK_ARG(Ns <= 160, "fc_up: Ns=%d exceeds the staged slice (160 rows)", Ns);
  dim3 grid(dg_ceil_div(K, 64), dg_ceil_div(B, 32));
  const size_t smem = (size_t)Ns * 256;      // [Ns][64] bf16 + [32][Ns] fp32
  if (small_f32)
    dg_launch(fc_up_kernel<float>, dg_cfg(grid, 256, smem, stream), (const float*)small, (const bf16*)wd, (bf16*)big, B, Ns, K);
  else
    dg_launch(fc_up_kernel<bf16>, dg_cfg(grid, 256, smem, stream), (const bf16*)small, (const bf16*)wd, (bf16*)big, B, Ns, K);
  DG_CHECK_LAUNCH("fc_up");
  return DG_OK;
}
int dg_fc_wgrad(const void* small, int small_f32, const void* big, float* dw, float beta, int B, int Ns, int C,
                cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && Ns > 0 && C > 0 && C % 8 == 0 && Ns <= 65535, "fc_wgrad: bad dims");
  dim3 grid(dg_ceil_div(2 * C, 128), Ns);
  if (small_f32)
    dg_launch(fc_wgrad_kernel<float>, dg_cfg(grid, 128, 0, stream), (const float*)small, (const bf16*)big, dw, beta, B, Ns, C);
  else
    dg_launch(fc_wgrad_kernel<bf16>, dg_cfg(grid, 128, 0, stream), (const bf16*)small, (const bf16*)big, dw, beta, B, Ns, C);
  DG_CHECK_LAUNCH("fc_wgrad");
  return DG_OK;
}

// Debug-only SIMT versions of the tensor-core convolutions (same signatures minus the workspace).
int dg_simt_conv4x4s2_fprop(const void* x, const void* wd, void* z, int B, int H, int W, int Cb, int Cs,
                            cudaStream_t stream) {
  const long long total = (long long)B * (H / 2) * (W / 2) * Cs;
  dg_launch(simt_down_kernel, dg_cfg(dg_ceil_div(total, 256), 256, 0, stream), (const bf16*)x, (const bf16*)wd, (bf16*)z, B, H / 2,
                                                               W / 2, Cs, Cb);
  DG_CHECK_LAUNCH("simt_down");
  return DG_OK;
}
int dg_simt_conv4x4s2_dgrad(const void* dz, const void* wu, void* dx, int B, int Hs, int Ws, int Cs, int Cb,
                            cudaStream_t stream) {
  const long long total = (long long)B * Hs * Ws * 4 * Cb;
  dg_launch(simt_up_kernel, dg_cfg(dg_ceil_div(total, 256), 256, 0, stream), (const bf16*)dz, (const bf16*)wu, (bf16*)dx, B, Hs, Ws,
                                                             Cs, Cb);
  DG_CHECK_LAUNCH("simt_up");
  return DG_OK;
}
int dg_simt_conv4x4s2_wgrad(const void* small, const void* big, float* dw, float beta, int B, int Hs, int Ws, int Cs,
                            int Cb, cudaStream_t stream) {
  const long long total = (long long)Cs * Cb * 16;
  dg_launch(simt_wgrad_kernel, dg_cfg(dg_ceil_div(total, 256), 256, 0, stream), (const bf16*)small, (const bf16*)big, dw, beta, B, Hs,
                                                                Ws, Cs, Cb);
  DG_CHECK_LAUNCH("simt_wgrad");
  return DG_OK;
}

}  // extern "C"
