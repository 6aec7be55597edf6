// HBM-bound glue kernels (sm_100a): weight packing, layout conversion, BatchNorm statistics + affine fused with
// LeakyReLU/ReLU (forward and backward), fused GAN-BCE / reconstruction-MSE / feature-matching losses, Adam.
// All activations are NHWC bf16 viewed as a [P, C] matrix (P = B*H*W); reductions are fp32 with a double-precision
// combine; 16-byte vector loads, warp-shuffle / shared-memory block reductions.
// Reference ops replaced: nn.BatchNorm2d + LeakyReLU/ReLU (model.py:12-33,84-140), nn.BCELoss / nn.MSELoss /
// get_fm_loss (image_translation.py:136-168,267-269,349-350), optim.Adam (image_translation.py:275-287).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_dg_launches{0};

void dg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

// ------------------------------------------------------------------------------------------------
// weight packing: W[Cs][Cb][16] fp32 -> Wd[Cs][16][Cb] bf16 (cb fastest), Wu[Cb][16][Cs] bf16 (cs fastest)
// ------------------------------------------------------------------------------------------------
__global__ void pack_wd_kernel(const float* __restrict__ w, bf16* __restrict__ wd, int Cs, int Cb) {
  griddep_launch_dependents();
  griddep_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)Cs * Cb) return;
  const int cb = (int)(idx % Cb);
  const int cs = (int)(idx / Cb);
  const float4* src = reinterpret_cast<const float4*>(w + idx * 16);
  float v[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 t = src[i];
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
#pragma unroll
  for (int t = 0; t < 16; ++t) wd[((size_t)cs * 16 + t) * Cb + cb] = __float2bfloat16(v[t]);
}
__global__ void pack_wu_kernel(const float* __restrict__ w, bf16* __restrict__ wu, int Cs, int Cb) {
  griddep_launch_dependents();
  griddep_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)Cs * Cb) return;
  const int cs = (int)(idx % Cs);
  const int cb = (int)(idx / Cs);
  const float4* src = reinterpret_cast<const float4*>(w + ((size_t)cs * Cb + cb) * 16);
  float v[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 t = src[i];
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
#pragma unroll
  for (int t = 0; t < 16; ++t) wu[((size_t)cb * 16 + t) * Cs + cs] = __float2bfloat16(v[t]);
}

// Multi-tensor version: one launch re-packs every GEMM weight of a network.  table[i] = {w, wd, wu, Cs, Cb, end}
// (pointers as integers, `end` = running total of 2*Cs*Cb items).  First Cs*Cb items of a tensor write Wd (cb fastest),
// the second Cs*Cb write Wu (cs fastest), so both stores stay coalesced.
__global__ void __launch_bounds__(256)
pack_multi_kernel(const long long* __restrict__ table, int n, long long total) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ long long tb[64 * 6];
  for (int i = threadIdx.x; i < n * 6; i += blockDim.x) tb[i] = table[i];
  __syncthreads();
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total;
       it += (long long)gridDim.x * blockDim.x) {
    int t = 0;
    while (it >= tb[t * 6 + 5]) ++t;
    const long long begin = t ? tb[(t - 1) * 6 + 5] : 0;
    const float* w = reinterpret_cast<const float*>(tb[t * 6]);
    bf16* wd = reinterpret_cast<bf16*>(tb[t * 6 + 1]);
    bf16* wu = reinterpret_cast<bf16*>(tb[t * 6 + 2]);
    const int Cs = (int)tb[t * 6 + 3], Cb = (int)tb[t * 6 + 4];
    long long idx = it - begin;
    const long long nn = (long long)Cs * Cb;
    const bool up = idx >= nn;
    if (up) idx -= nn;
    int cs, cb;
    if (!up) {
      cb = (int)(idx % Cb);
      cs = (int)(idx / Cb);
      if (!wd) continue;
    } else {
      cs = (int)(idx % Cs);
      cb = (int)(idx / Cs);
      if (!wu) continue;
    }
    const float4* src = reinterpret_cast<const float4*>(w + ((size_t)cs * Cb + cb) * 16);
    float v[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 q = src[i];
      v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
    if (!up) {
#pragma unroll
      for (int k = 0; k < 16; ++k) wd[((size_t)cs * 16 + k) * Cb + cb] = __float2bfloat16(v[k]);
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) wu[((size_t)cb * 16 + k) * Cs + cs] = __float2bfloat16(v[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// layout: NHWC bf16 [B][HW][C] <-> NCHW fp32 [B][C][HW], 32x32 smem tiles
// ------------------------------------------------------------------------------------------------
__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ x, float* __restrict__ y, int HW, int C) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) tile[i][threadIdx.x] = __bfloat162float(x[((size_t)b * HW + p) * C + c]);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (p < HW && c < C) y[((size_t)b * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, bf16* __restrict__ y, int HW, int C) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (p < HW && c < C) tile[i][threadIdx.x] = x[((size_t)b * C + c) * HW + p];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) y[((size_t)b * HW + p) * C + c] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm over [P][C]
// Block = 256 threads as (cw channel-vectors) x (256/cw rows); VEC = 8 (16-byte loads) or 1 (any C).
// ------------------------------------------------------------------------------------------------
struct BnGeom {
  int cw;          // channel vectors per block row (power of two <= 32)
  int rows_iter;   // 256 / cw
  int gx;          // blocks along channels
  int gy;          // row splits
  int rows_split;  // rows per split (multiple of rows_iter)
};

BnGeom bn_geom(long long P, int C, int vec, int sms, int blocks_per_sm = 4) {
  BnGeom g;
  const int cv = (C + vec - 1) / vec;
  int cw = 1;
  while (cw < cv && cw < 32) cw <<= 1;
  g.cw = cw;
  g.rows_iter = 256 / cw;
  g.gx = (cv + cw - 1) / cw;
  long long want = ((long long)blocks_per_sm * sms + g.gx - 1) / g.gx;
  long long max_splits = (P + g.rows_iter * 4 - 1) / (g.rows_iter * 4);
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  if (want > 2048) want = 2048;
  long long rs = (P + want - 1) / want;
  rs = (rs + g.rows_iter - 1) / g.rows_iter * g.rows_iter;
  g.rows_split = (int)rs;
  g.gy = (int)((P + rs - 1) / rs);
  return g;
}

// "Folded finalize": instead of a separate finalize launch between the reduction and its consumer, reductions add their
// block sums atomically into one zero-initialised accumulator pair per layer (acc_out), and the consumer derives its
// per-channel coefficients from those sums in its prologue (acc); the first row-block also publishes what the finalize
// kernel used to publish (statistics + running statistics, or dgamma / dbeta).  Two launches fewer per BatchNorm layer
// and pass on the dependency chain of the (latency-bound) 64x64 step.
struct BnFold {
  const float* acc;        // consumer: {sum, sum2}[C] (forward) or {sum g', sum g'*xhat}[C] (backward dx); NULL = off
  float* acc_out;          // reduction: add block sums here instead of writing a partial row; NULL = off
  const float* gamma;
  const float* beta;
  float* stats_out;        // forward consumer: {mean, invstd, scale, shift}[C]
  float* running_mean;
  float* running_var;
  float* dgamma;           // backward consumer
  float* dbeta;
  float grad_beta;
  float eps, momentum;
  double inv_P, unbias;    // 1/P and P/(P-1)
};

// forward coefficients of channel c from the accumulated sums (same arithmetic as bn_stats_finalize_kernel)
__device__ __forceinline__ void bn_fold_fwd(const BnFold& f, int C, int c, bool publish, float* sc, float* sh) {
  const double m = (double)f.acc[c] * f.inv_P;
  double var = (double)f.acc[C + c] * f.inv_P - m * m;
  if (var < 0.0) var = 0.0;
  const float is = (float)(1.0 / sqrt(var + (double)f.eps));
  const float scv = f.gamma[c] * is;
  *sc = scv;
  *sh = f.beta[c] - (float)m * scv;
  if (publish) {
    f.stats_out[c] = (float)m;
    f.stats_out[C + c] = is;
    f.stats_out[2 * C + c] = scv;
    f.stats_out[3 * C + c] = *sh;
    if (f.running_mean) {
      f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * (float)m;
      f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * (float)(var * f.unbias);
    }
  }
}

// backward-dx coefficients {gamma*invstd, mean(g'), mean(g'*xhat)} of channel c (bn_bwd_finalize_kernel's arithmetic)
__device__ __forceinline__ void bn_fold_bwd(const BnFold& f, const float* stats, int C, int c, bool publish, float* k0,
                                            float* k1, float* k2) {
  const float sg = f.acc[c], sgx = f.acc[C + c];
  *k0 = f.gamma[c] * stats[C + c];
  *k1 = (float)((double)sg * f.inv_P);
  *k2 = (float)((double)sgx * f.inv_P);
  if (publish && f.dgamma) {
    f.dgamma[c] = (f.grad_beta != 0.f ? f.grad_beta * f.dgamma[c] : 0.f) + sgx;
    f.dbeta[c] = (f.grad_beta != 0.f ? f.grad_beta * f.dbeta[c] : 0.f) + sg;
  }
}

template <int VEC>
__device__ __forceinline__ void load_vec(const bf16* p, float* f) {
  if (VEC == 8) {
    bf16x8 v = *reinterpret_cast<const bf16x8*>(p);
    unpack8(v, f);
  } else {
    f[0] = __bfloat162float(*p);
  }
}
template <int VEC>
__device__ __forceinline__ void store_vec(bf16* p, const float* f) {
  if (VEC == 8) {
    *reinterpret_cast<bf16x8*>(p) = pack8(f);
  } else {
    *p = __float2bfloat16(f[0]);
  }
}

// Block reduction of VEC*2 per-thread values over the row dimension; result written to part{1,2}[split][C].
template <int VEC>
__device__ __forceinline__ void bn_block_reduce_store(float* s1, float* s2, int cw, int rows_iter, int tx, int ty,
                                                      int c, int C, float* part1, float* part2, int split,
                                                      float* acc_out = nullptr) {
  __shared__ float sh1[256 * VEC];
  __shared__ float sh2[256 * VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    sh1[(ty * cw + tx) * VEC + i] = s1[i];
    sh2[(ty * cw + tx) * VEC + i] = s2[i];
  }
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float a = 0.f, b = 0.f;
      for (int r = 0; r < rows_iter; ++r) {
        a += sh1[(r * cw + tx) * VEC + i];
        b += sh2[(r * cw + tx) * VEC + i];
      }
      if (c + i < C) {
        if (acc_out) {
          atomicAdd(acc_out + c + i, a);
          atomicAdd(acc_out + C + c + i, b);
        } else {
          part1[(size_t)split * C + c + i] = a;
          part2[(size_t)split * C + c + i] = b;
        }
      }
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
bn_stats_partial_kernel(const bf16* __restrict__ z, long long P, int C, int cw, int rows_iter, int rows_split,
                        float* __restrict__ part_sum, float* __restrict__ part_sq, float* __restrict__ acc_out) {
  griddep_launch_dependents();
  griddep_wait();
  const int tx = threadIdx.x % cw, ty = threadIdx.x / cw;
  const int c = (blockIdx.x * cw + tx) * VEC;
  float s1[VEC], s2[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) s1[i] = s2[i] = 0.f;
  if (c < C) {
    const long long r0 = (long long)blockIdx.y * rows_split;
    const long long r1 = min(P, r0 + rows_split);
    long long r = r0 + ty;
    for (; r + 3LL * rows_iter < r1; r += 4LL * rows_iter) {   // 4 independent 16-byte loads in flight
      float f[4][VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u) load_vec<VEC>(z + (r + (long long)u * rows_iter) * C + c, f[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          s1[i] += f[u][i];
          s2[i] += f[u][i] * f[u][i];
        }
    }
    for (; r < r1; r += rows_iter) {
      float f[VEC];
      load_vec<VEC>(z + r * C + c, f);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        s1[i] += f[i];
        s2[i] += f[i] * f[i];
      }
    }
  }
  bn_block_reduce_store<VEC>(s1, s2, cw, rows_iter, tx, ty, c, C, part_sum, part_sq, blockIdx.y, acc_out);
}

// mean/invstd + scale/shift for the apply kernel + running statistics (momentum update, unbiased variance).
__global__ void bn_stats_finalize_kernel(const float* __restrict__ part_sum, const float* __restrict__ part_sq,
                                         int splits, long long P, int C, float eps, float momentum,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float* __restrict__ mean, float* __restrict__ invstd,
                                         float* __restrict__ scale, float* __restrict__ shift,
                                         float* __restrict__ running_mean, float* __restrict__ running_var) {
  griddep_launch_dependents();
  griddep_wait();
  // block = 32 channels x 8 split lanes; partial rows are summed 8-way in parallel, then combined in smem
  __shared__ double sh_s[8][32], sh_q[8][32];
  const int cx = threadIdx.x & 31, sy = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  // fp32 running sums per lane (4 independent chains), double only for the final combine: FP64 add chains are an
  // order of magnitude slower on this part
  float fs[4] = {0.f, 0.f, 0.f, 0.f}, fq[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    int i = sy;
    for (; i + 24 < splits; i += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        fs[u] += part_sum[(size_t)(i + 8 * u) * C + c];
        fq[u] += part_sq[(size_t)(i + 8 * u) * C + c];
      }
    }
    for (; i < splits; i += 8) {
      fs[0] += part_sum[(size_t)i * C + c];
      fq[0] += part_sq[(size_t)i * C + c];
    }
  }
  double s = ((double)fs[0] + (double)fs[1]) + ((double)fs[2] + (double)fs[3]);
  double q = ((double)fq[0] + (double)fq[1]) + ((double)fq[2] + (double)fq[3]);
  sh_s[sy][cx] = s;
  sh_q[sy][cx] = q;
  __syncthreads();
  if (sy != 0 || c >= C) return;
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    s += sh_s[i][cx];
    q += sh_q[i][cx];
  }
  const double m = s / (double)P;
  double var = q / (double)P - m * m;
  if (var < 0.0) var = 0.0;
  const float is = (float)(1.0 / sqrt(var + (double)eps));
  mean[c] = (float)m;
  invstd[c] = is;
  const float sc = gamma[c] * is;
  scale[c] = sc;
  shift[c] = beta[c] - (float)m * sc;
  if (running_mean) {
    const double unbiased = P > 1 ? var * (double)P / (double)(P - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// eval mode: scale/shift from running statistics
__global__ void bn_eval_coeffs_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                      float eps, int C, float* __restrict__ scale, float* __restrict__ shift) {
  griddep_launch_dependents();
  griddep_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float is = rsqrtf(running_var[c] + eps);
  const float sc = gamma[c] * is;
  scale[c] = sc;
  shift[c] = beta[c] - running_mean[c] * sc;
}

template <int VEC>
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const bf16* __restrict__ z, bf16* __restrict__ y, long long P, int C, int cw, int rows_iter,
                  int rows_split, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                  float slope, const BnFold fold) {
  griddep_launch_dependents();
  griddep_wait();
  // same (channel vectors) x (rows) block shape as the reductions: a thread keeps its channels' coefficients in
  // registers and streams rows -- no per-element index arithmetic
  const int tx = threadIdx.x % cw, ty = threadIdx.x / cw;
  const int c = (blockIdx.x * cw + tx) * VEC;
  if (c >= C) return;
  float sc[VEC], sh[VEC];
  if (fold.acc) {
    const bool publish = blockIdx.y == 0 && ty == 0;
#pragma unroll
    for (int i = 0; i < VEC; ++i) bn_fold_fwd(fold, C, c + i, publish, &sc[i], &sh[i]);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      sc[i] = scale[c + i];
      sh[i] = shift[c + i];
    }
  }
  const long long r0 = (long long)blockIdx.y * rows_split;
  const long long r1 = min(P, r0 + rows_split);
  long long r = r0 + ty;
  for (; r + 3LL * rows_iter < r1; r += 4LL * rows_iter) {
    float f[4][VEC];
#pragma unroll
    for (int u = 0; u < 4; ++u) load_vec<VEC>(z + (r + (long long)u * rows_iter) * C + c, f[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) f[u][k] = act_fwd(f[u][k] * sc[k] + sh[k], act, slope);
      store_vec<VEC>(y + (r + (long long)u * rows_iter) * C + c, f[u]);
    }
  }
  for (; r < r1; r += rows_iter) {
    float f[VEC];
    load_vec<VEC>(z + r * C + c, f);
#pragma unroll
    for (int k = 0; k < VEC; ++k) f[k] = act_fwd(f[k] * sc[k] + sh[k], act, slope);
    store_vec<VEC>(y + r * C + c, f);
  }
}

// backward pass 1: per-channel sum(g) and sum(g * xhat), g = (dy [+ dy2] [+ coef*bcast]) * act'(y)
// upstream gradient g = dy [+ dy2] [+ coef*bcast[row % bcast_rows]] (before the activation derivative)
template <int VEC>
__device__ __forceinline__ void bn_bwd_load_g(const bf16* dy, const bf16* dy2, const float* bcast, float coef,
                                              long long bcast_rows, long long r, int C, int c, float* g) {
  load_vec<VEC>(dy + r * C + c, g);
  if (dy2) {
    float t[VEC];
    load_vec<VEC>(dy2 + r * C + c, t);
#pragma unroll
    for (int i = 0; i < VEC; ++i) g[i] += t[i];
  }
  if (bcast) {
    const float* bp = bcast + (r % bcast_rows) * C + c;
#pragma unroll
    for (int i = 0; i < VEC; ++i) g[i] += coef * bp[i];
  }
}

// backward pass 1: per-channel sum(g') and sum(g' * xhat), g' = g * act'(pre).  The activation derivative is taken from
// the recomputed pre-activation z*scale+shift (same fp32 expression as the forward), so y is not read.
template <int VEC>
__global__ void __launch_bounds__(256, 2)
bn_bwd_partial_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ dy2, const float* __restrict__ bcast,
                      float coef, long long bcast_rows, const bf16* __restrict__ z, const float* __restrict__ stats,
                      long long P, int C, int cw, int rows_iter, int rows_split, int act, float slope,
                      float* __restrict__ part_g, float* __restrict__ part_gx, float* __restrict__ acc_out) {
  griddep_launch_dependents();
  griddep_wait();
  const int tx = threadIdx.x % cw, ty = threadIdx.x / cw;
  const int c = (blockIdx.x * cw + tx) * VEC;
  float s1[VEC], s2[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) s1[i] = s2[i] = 0.f;
  if (c < C) {
    float mu[VEC], sc[VEC], sh[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      mu[i] = stats[c + i];
      sc[i] = stats[2 * C + c + i];
      sh[i] = stats[3 * C + c + i];
    }
    const long long r0 = (long long)blockIdx.y * rows_split;
    const long long r1 = min(P, r0 + rows_split);
    long long r = r0 + ty;
    for (; r + rows_iter < r1; r += 2LL * rows_iter) {   // two rows (4+ independent 16-byte loads) in flight
      float g[2][VEC], fz[2][VEC];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        bn_bwd_load_g<VEC>(dy, dy2, bcast, coef, bcast_rows, r + (long long)u * rows_iter, C, c, g[u]);
        load_vec<VEC>(z + (r + (long long)u * rows_iter) * C + c, fz[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float gg = g[u][i] * act_grad_from_out(fz[u][i] * sc[i] + sh[i], act, slope);
          s1[i] += gg;
          s2[i] += gg * (fz[u][i] - mu[i]);
        }
    }
    for (; r < r1; r += rows_iter) {
      float g[VEC], fz[VEC];
      bn_bwd_load_g<VEC>(dy, dy2, bcast, coef, bcast_rows, r, C, c, g);
      load_vec<VEC>(z + r * C + c, fz);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float gg = g[i] * act_grad_from_out(fz[i] * sc[i] + sh[i], act, slope);
        s1[i] += gg;
        s2[i] += gg * (fz[i] - mu[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) s2[i] *= stats[C + c + i];   // xhat = (z - mean) * invstd
  }
  bn_block_reduce_store<VEC>(s1, s2, cw, rows_iter, tx, ty, c, C, part_g, part_gx, blockIdx.y, acc_out);
}

// dgamma/dbeta (accumulated into the fp32 grads with `beta_acc`) and the three per-channel dx coefficients
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ part_g, const float* __restrict__ part_gx, int splits,
                                       long long P, int C, const float* __restrict__ gamma,
                                       const float* __restrict__ invstd, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float beta_acc, float* __restrict__ coefs) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ double sh_s[8][32], sh_q[8][32];
  const int cx = threadIdx.x & 31, sy = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float fs[4] = {0.f, 0.f, 0.f, 0.f}, fq[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    int i = sy;
    for (; i + 24 < splits; i += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        fs[u] += part_g[(size_t)(i + 8 * u) * C + c];
        fq[u] += part_gx[(size_t)(i + 8 * u) * C + c];
      }
    }
    for (; i < splits; i += 8) {
      fs[0] += part_g[(size_t)i * C + c];
      fq[0] += part_gx[(size_t)i * C + c];
    }
  }
  double sg = ((double)fs[0] + (double)fs[1]) + ((double)fs[2] + (double)fs[3]);
  double sgx = ((double)fq[0] + (double)fq[1]) + ((double)fq[2] + (double)fq[3]);
  sh_s[sy][cx] = sg;
  sh_q[sy][cx] = sgx;
  __syncthreads();
  if (sy != 0 || c >= C) return;
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    sg += sh_s[i][cx];
    sgx += sh_q[i][cx];
  }
  if (dgamma) {
    dgamma[c] = (beta_acc != 0.f ? beta_acc * dgamma[c] : 0.f) + (float)sgx;
    dbeta[c] = (beta_acc != 0.f ? beta_acc * dbeta[c] : 0.f) + (float)sg;
  }
  coefs[c] = gamma[c] * invstd[c];
  coefs[C + c] = (float)(sg / (double)P);
  coefs[2 * C + c] = (float)(sgx / (double)P);
}

// backward pass 2: dz = gamma*invstd * (g - mean(g) - xhat * mean(g*xhat))
template <int VEC>
__global__ void __launch_bounds__(256, 2)
bn_bwd_dx_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ dy2, const float* __restrict__ bcast, float coef,
                 long long bcast_rows, const bf16* __restrict__ z, const float* __restrict__ stats,
                 const float* __restrict__ coefs, bf16* __restrict__ dz, long long P, int C, int cw, int rows_iter,
                 int rows_split, int act, float slope, const BnFold fold) {
  griddep_launch_dependents();
  griddep_wait();
  const int tx = threadIdx.x % cw, ty = threadIdx.x / cw;
  const int c = (blockIdx.x * cw + tx) * VEC;
  if (c >= C) return;
  // dz = k0*(g' - k1 - (z-mu)*is*k2) = k0*g' + a1*z + a0
  float sc[VEC], sh[VEC], k0[VEC], a1[VEC], a0[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float mu = stats[c + i], is = stats[C + c + i];
    sc[i] = stats[2 * C + c + i];
    sh[i] = stats[3 * C + c + i];
    float k1, k2;
    if (fold.acc) {
      bn_fold_bwd(fold, stats, C, c + i, blockIdx.y == 0 && ty == 0, &k0[i], &k1, &k2);
    } else {
      k0[i] = coefs[c + i];
      k1 = coefs[C + c + i];
      k2 = coefs[2 * C + c + i];
    }
    a1[i] = -k0[i] * k2 * is;
    a0[i] = -k0[i] * k1 - a1[i] * mu;
  }
  const long long r0 = (long long)blockIdx.y * rows_split;
  const long long r1 = min(P, r0 + rows_split);
  long long r = r0 + ty;
  for (; r + rows_iter < r1; r += 2LL * rows_iter) {
    float g[2][VEC], fz[2][VEC];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      bn_bwd_load_g<VEC>(dy, dy2, bcast, coef, bcast_rows, r + (long long)u * rows_iter, C, c, g[u]);
      load_vec<VEC>(z + (r + (long long)u * rows_iter) * C + c, fz[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float gg = g[u][k] * act_grad_from_out(fz[u][k] * sc[k] + sh[k], act, slope);
        g[u][k] = k0[k] * gg + a1[k] * fz[u][k] + a0[k];
      }
      store_vec<VEC>(dz + (r + (long long)u * rows_iter) * C + c, g[u]);
    }
  }
  for (; r < r1; r += rows_iter) {
    float g[VEC], fz[VEC];
    bn_bwd_load_g<VEC>(dy, dy2, bcast, coef, bcast_rows, r, C, c, g);
    load_vec<VEC>(z + r * C + c, fz);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float gg = g[k] * act_grad_from_out(fz[k] * sc[k] + sh[k], act, slope);
      g[k] = k0[k] * gg + a1[k] * fz[k] + a0[k];
    }
    store_vec<VEC>(dz + r * C + c, g);
  }
}


// ------------------------------------------------------------------------------------------------
// Streaming variants of the three big BatchNorm passes for C <= 256 (the layers with the large P): [P, C] is contiguous,
// so a tile of R rows is ONE 1-D bulk copy (cp.async.bulk -> shared memory, mbarrier completion).  Thread 0 keeps
// kBsStages tiles of every input stream in flight per block, i.e. the bytes in flight no longer depend on registers or
// occupancy -- the register-staged kernels above top out at 45-60 % of the HBM copy bandwidth because they cannot
// keep more than ~32 KB per SM outstanding.  Threads read their 16-byte channel vector from the staged tile
// (conflict-free: consecutive threads, consecutive 16 bytes) and store results straight to global memory.
// ------------------------------------------------------------------------------------------------
constexpr int kBsTileBytes = 16384;
constexpr int kBsStages = 3;
enum { BS_FWD = 0, BS_BWD_PARTIAL = 1, BS_BWD_DX = 2 };

struct BnStreamParams {
  const bf16* in[3];   // FWD: {z}; BWD: {dy, z, dy2 or NULL}
  int nstreams;
  bf16* out;           // FWD: y, BWD_DX: dz
  const float* bcast;  // optional broadcast gradient [bcast_rows][C] (fp32), added as coef * bcast[row % bcast_rows]
  float bcast_coef;
  long long bcast_rows;
  const float* stats;  // {mean, invstd, scale, shift}[C]
  const float* coefs;  // BWD_DX: {gamma*invstd, mean(g), mean(g*xhat)}[C]
  float* part0;        // BWD_PARTIAL: per-block partial sums [gridDim.x][C]
  float* part1;
  long long P;
  int C, rows_tile;
  long long tiles;
  int tiles_per_block;
  int act;
  float slope;
  BnFold fold;
};

__device__ __forceinline__ void bulk_load_1d(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem)),
               "l"(reinterpret_cast<uint64_t>(gmem)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256)
bn_stream_kernel(const BnStreamParams p) {
  griddep_launch_dependents();
  griddep_wait();
  extern __shared__ __align__(128) uint8_t bs_smem[];
  __shared__ uint64_t full_bar[kBsStages];
  const int C = p.C, cw = C >> 3, rows_iter = 256 / cw;
  const int tx = threadIdx.x % cw, ty = threadIdx.x / cw;
  const bool active = ty < rows_iter;
  const int c = tx * 8;
  const long long t0 = (long long)blockIdx.x * p.tiles_per_block;
  const int ntiles = (int)min((long long)p.tiles_per_block, p.tiles - t0);
  const int R = p.rows_tile;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kBsStages; ++s) mbar_init(&full_bar[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int i) {
    const long long row0 = (t0 + i) * R;
    const uint32_t bytes = (uint32_t)min((long long)R, p.P - row0) * (uint32_t)C * 2u;
    const int s = i % kBsStages;
    mbar_arrive_expect_tx(&full_bar[s], bytes * (uint32_t)p.nstreams);
    for (int k = 0; k < p.nstreams; ++k)
      bulk_load_1d(bs_smem + (size_t)(s * p.nstreams + k) * kBsTileBytes, p.in[k] + row0 * C, bytes, &full_bar[s]);
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < kBsStages && i < ntiles; ++i) issue(i);

  // per-channel constants of this thread's 8 channels
  float sc[8], sh[8], k0[8], a1[8], a0[8], mu[8], s1[8], s2[8];
  const bool publish = blockIdx.x == 0 && ty == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s1[i] = s2[i] = 0.f;
    mu[i] = k0[i] = a1[i] = a0[i] = 0.f;
    if (MODE == BS_FWD && p.fold.acc) {
      bn_fold_fwd(p.fold, C, c + i, publish, &sc[i], &sh[i]);
      continue;
    }
    sc[i] = p.stats[2 * C + c + i];
    sh[i] = p.stats[3 * C + c + i];
    if (MODE == BS_BWD_PARTIAL) mu[i] = p.stats[c + i];
    if (MODE == BS_BWD_DX) {
      const float m = p.stats[c + i], is = p.stats[C + c + i];
      float k1, k2;
      if (p.fold.acc) {
        bn_fold_bwd(p.fold, p.stats, C, c + i, publish, &k0[i], &k1, &k2);
      } else {
        k0[i] = p.coefs[c + i];
        k1 = p.coefs[C + c + i];
        k2 = p.coefs[2 * C + c + i];
      }
      a1[i] = -k0[i] * k2 * is;                // dz = k0*(g' - k1 - (z-mu)*is*k2) = k0*g' + a1*z + a0
      a0[i] = -k0[i] * k1 - a1[i] * m;
    }
  }

  for (int i = 0; i < ntiles; ++i) {
    const int s = i % kBsStages;
    mbar_wait(&full_bar[s], (uint32_t)(i / kBsStages) & 1u);
    const long long row0 = (t0 + i) * R;
    const int rows = (int)min((long long)R, p.P - row0);
    const uint8_t* st = bs_smem + (size_t)(s * p.nstreams) * kBsTileBytes;
    if (active) {
      for (int r = ty; r < rows; r += rows_iter) {
        const size_t off = ((size_t)r * C + c) * 2;
        float f[8];
        unpack8(*reinterpret_cast<const bf16x8*>(st + off), f);
        if (MODE == BS_FWD) {
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = act_fwd(f[k] * sc[k] + sh[k], p.act, p.slope);
          *reinterpret_cast<bf16x8*>(p.out + (row0 + r) * C + c) = pack8(f);
        } else {
          float z[8];
          unpack8(*reinterpret_cast<const bf16x8*>(st + kBsTileBytes + off), z);
          if (p.nstreams == 3) {
            float t[8];
            unpack8(*reinterpret_cast<const bf16x8*>(st + 2 * kBsTileBytes + off), t);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] += t[k];
          }
          if (p.bcast) {
            const float* bp = p.bcast + ((row0 + r) % p.bcast_rows) * C + c;
            const float4 b0 = *reinterpret_cast<const float4*>(bp), b1 = *reinterpret_cast<const float4*>(bp + 4);
            f[0] += p.bcast_coef * b0.x; f[1] += p.bcast_coef * b0.y; f[2] += p.bcast_coef * b0.z; f[3] += p.bcast_coef * b0.w;
            f[4] += p.bcast_coef * b1.x; f[5] += p.bcast_coef * b1.y; f[6] += p.bcast_coef * b1.z; f[7] += p.bcast_coef * b1.w;
          }
          if (MODE == BS_BWD_PARTIAL) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float gg = f[k] * act_grad_from_out(z[k] * sc[k] + sh[k], p.act, p.slope);
              s1[k] += gg;
              s2[k] += gg * (z[k] - mu[k]);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float gg = f[k] * act_grad_from_out(z[k] * sc[k] + sh[k], p.act, p.slope);
              f[k] = k0[k] * gg + a1[k] * z[k] + a0[k];
            }
            *reinterpret_cast<bf16x8*>(p.out + (row0 + r) * C + c) = pack8(f);
          }
        }
      }
    }
    __syncthreads();   // every thread is done with stage s: refill it
    if (threadIdx.x == 0 && i + kBsStages < ntiles) issue(i + kBsStages);
  }
  if (MODE == BS_BWD_PARTIAL) {
    // block reduction over the row lanes in the (now idle) staging buffer; one partial row per block
    float* sh1 = reinterpret_cast<float*>(bs_smem);
    float* sh2 = sh1 + 256 * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sh1[threadIdx.x * 8 + k] = s1[k];
      sh2[threadIdx.x * 8 + k] = active ? s2[k] * p.stats[C + c + k] : 0.f;   // xhat = (z - mean) * invstd
    }
    __syncthreads();
    if (ty == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float a = 0.f, b = 0.f;
        for (int r = 0; r < rows_iter; ++r) {
          a += sh1[(r * cw + tx) * 8 + k];
          b += sh2[(r * cw + tx) * 8 + k];
        }
        if (p.fold.acc_out) {
          atomicAdd(p.fold.acc_out + c + k, a);
          atomicAdd(p.fold.acc_out + C + c + k, b);
        } else {
          p.part0[(size_t)blockIdx.x * C + c + k] = a;
          p.part1[(size_t)blockIdx.x * C + c + k] = b;
        }
      }
    }
  }
}

// geometry of a streaming launch; returns false when the shape does not qualify
struct BnStreamPlan {
  int grid, tiles_per_block, rows_tile, smem;
  long long tiles;
};
bool bn_stream_plan(long long P, int C, int nstreams, int sms, BnStreamPlan* pl) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("DG_BN_STREAM");
    mode = e ? atoi(e) : 1;
  }
  if (!mode || C % 8 != 0 || C > 256 || C < 8) return false;
  if ((long long)P * C * 2 < (4LL << 20)) return false;   // small tensors are latency-bound: keep the light kernels
  pl->rows_tile = kBsTileBytes / (C * 2);
  pl->tiles = (P + pl->rows_tile - 1) / pl->rows_tile;
  pl->smem = kBsStages * nstreams * kBsTileBytes;
  const int per_sm = (224 * 1024) / (pl->smem + 2 * 1024);
  long long blocks = (long long)sms * (per_sm < 1 ? 1 : (per_sm > 3 ? 3 : per_sm));
  if (blocks > pl->tiles) blocks = pl->tiles;
  pl->tiles_per_block = (int)((pl->tiles + blocks - 1) / blocks);
  pl->grid = (int)((pl->tiles + pl->tiles_per_block - 1) / pl->tiles_per_block);
  return true;
}

template <int MODE>
int bn_stream_launch(const BnStreamParams& p, const BnStreamPlan& pl, cudaStream_t stream) {
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bn_stream_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) {
      dg_set_error("bn stream: cannot raise dynamic smem: %s", cudaGetErrorString(e));
      return DG_ERR_CUDA;
    }
    attr_set = true;
  }
  dg_launch(bn_stream_kernel<MODE>, dg_cfg(pl.grid, 256, pl.smem, stream), p);
  DG_CHECK_LAUNCH("bn_stream_kernel");
  return DG_OK;
}

// ------------------------------------------------------------------------------------------------
// losses
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v) {
  __shared__ float sh[8];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

// GAN BCE on sigmoid(logit): out[0]=dis_loss=0.5*(BCE(pr,1)+BCE(pf,0)), out[1]=gen_loss=BCE(pf,1); log clamp -100
// (nn.BCELoss semantics).  Also writes the probabilities.
__global__ void gan_bce_fwd_kernel(const float* __restrict__ logit_real, const float* __restrict__ logit_fake, int B,
                                   float* __restrict__ p_real, float* __restrict__ p_fake, float* __restrict__ out) {
  griddep_launch_dependents();
  griddep_wait();
  float a = 0.f, b = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float pr = 1.f / (1.f + expf(-logit_real[i]));
    const float pf = 1.f / (1.f + expf(-logit_fake[i]));
    p_real[i] = pr;
    p_fake[i] = pf;
    a += -fmaxf(logf(pr), -100.f);
    b += -fmaxf(logf(1.f - pf), -100.f);
    c += -fmaxf(logf(pf), -100.f);
  }
  a = block_sum_256(a);
  b = block_sum_256(b);
  c = block_sum_256(c);
  if (threadIdx.x == 0) {
    out[0] = 0.5f * (a + b) / (float)B;
    out[1] = c / (float)B;
  }
}
// d/dlogit of (g_dis*dis_loss + g_gen*gen_loss) through BCELoss and Sigmoid exactly as autograd composes them:
// dL/dp = (p - y) / max(p(1-p), 1e-12) / B, then * p(1-p).  (p==0 or 1 in fp32 gives 0, as in the reference.)
__global__ void gan_bce_bwd_kernel(const float* __restrict__ p_real, const float* __restrict__ p_fake, int B,
                                   float g_dis, float g_gen, float* __restrict__ dlogit_real,
                                   float* __restrict__ dlogit_fake) {
  griddep_launch_dependents();
  griddep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float pr = p_real[i], pf = p_fake[i];
  const float invB = 1.f / (float)B;
  const float sr = pr * (1.f - pr), sf = pf * (1.f - pf);
  const float dpr = 0.5f * g_dis * (pr - 1.f) / fmaxf(sr, 1e-12f) * invB;
  const float dpf = (0.5f * g_dis * (pf - 0.f) + g_gen * (pf - 1.f)) / fmaxf(sf, 1e-12f) * invB;
  if (dlogit_real) dlogit_real[i] = dpr * sr;
  if (dlogit_fake) dlogit_fake[i] = dpf * sf;
}

// sigmoid forward/backward on tiny [B] vectors (module API path: prob output of the Discriminator)
__global__ void sigmoid_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int n) {
  griddep_launch_dependents();
  griddep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = 1.f / (1.f + expf(-x[i]));
}
__global__ void sigmoid_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx,
                                   int n) {
  griddep_launch_dependents();
  griddep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = dy[i] * y[i] * (1.f - y[i]);
}

// MSE: partial sums of (a-b)^2 per block, then finalize to mean
__global__ void __launch_bounds__(256)
mse_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float* __restrict__ part) {
  griddep_launch_dependents();
  griddep_wait();
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = a4[i], y = b4[i];
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    s += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const float d = a[i] - b[i];
      s += d * d;
    }
  s = block_sum_256(s);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void sum_finalize_kernel(const float* __restrict__ part, int n, double scale, float* __restrict__ out,
                                    int accumulate) {
  griddep_launch_dependents();
  griddep_wait();
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) s += (double)part[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) {
    const float v = (float)(s * scale);
    out[0] = accumulate ? out[0] + v : v;
  }
}
// da (+)= coef * (a - b)
__global__ void __launch_bounds__(256)
mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float coef,
               float* __restrict__ da, int accumulate) {
  griddep_launch_dependents();
  griddep_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = coef * (a[i] - b[i]);
    da[i] = accumulate ? da[i] + v : v;
  }
}

// Feature matching on NHWC bf16 feats [B][n] (n = H*W*C): d = mean_b real - mean_b fake (kept fp32 for backward),
// block partial sums of d^2.
__global__ void __launch_bounds__(256)
fm_partial_kernel(const bf16* __restrict__ real, const bf16* __restrict__ fake, int B, long long n,
                  float* __restrict__ diff, float* __restrict__ part) {
  griddep_launch_dependents();
  griddep_wait();
  // block = 32 element vectors x 8 batch slices: a thread sums (real - fake) over every 8th image, so the batch loop --
  // pure load latency -- is 8x shorter and 8x more loads are in flight than with one thread per vector (42 -> ~8 us at
  // 64x64, B = 64); the slices are combined through shared memory
  __shared__ float sh[8][32][8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long n8 = n >> 3;
  const float invB = 1.f / (float)B;
  float s = 0.f;
  for (long long i0 = (long long)blockIdx.x * 32; i0 < n8; i0 += (long long)gridDim.x * 32) {
    const long long i = i0 + tx;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    if (i < n8) {
      int b = ty;
      for (; b + 24 < B; b += 32) {   // 8 independent 16-byte loads in flight
        bf16x8 vr[4], vf[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          vr[u] = *reinterpret_cast<const bf16x8*>(real + (size_t)(b + 8 * u) * n + i * 8);
          vf[u] = *reinterpret_cast<const bf16x8*>(fake + (size_t)(b + 8 * u) * n + i * 8);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float fr[8], ff[8];
          unpack8(vr[u], fr);
          unpack8(vf[u], ff);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += fr[k] - ff[k];
        }
      }
      for (; b < B; b += 8) {
        float fr[8], ff[8];
        unpack8(*reinterpret_cast<const bf16x8*>(real + (size_t)b * n + i * 8), fr);
        unpack8(*reinterpret_cast<const bf16x8*>(fake + (size_t)b * n + i * 8), ff);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += fr[k] - ff[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sh[ty][tx][k] = acc[k];
    __syncthreads();
    if (ty == 0 && i < n8) {
      float d[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += sh[j][tx][k];
        d[k] = t * invB;
        s += d[k] * d[k];
      }
      if (diff) {
        *reinterpret_cast<float4*>(diff + i * 8) = make_float4(d[0], d[1], d[2], d[3]);
        *reinterpret_cast<float4*>(diff + i * 8 + 4) = make_float4(d[4], d[5], d[6], d[7]);
      }
    }
    __syncthreads();
  }
  s = block_sum_256(s);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
// dfeat[b][i] = coef * diff[i]  (bf16, broadcast over batch)
__global__ void __launch_bounds__(256)
fm_bwd_kernel(const float* __restrict__ diff, int B, long long n, float coef, bf16* __restrict__ dfeat) {
  griddep_launch_dependents();
  griddep_wait();
  const long long n8 = n >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float d[8];
    const float4 a = *reinterpret_cast<const float4*>(diff + i * 8);
    const float4 b4 = *reinterpret_cast<const float4*>(diff + i * 8 + 4);
    d[0] = coef * a.x; d[1] = coef * a.y; d[2] = coef * a.z; d[3] = coef * a.w;
    d[4] = coef * b4.x; d[5] = coef * b4.y; d[6] = coef * b4.z; d[7] = coef * b4.w;
    const bf16x8 v = pack8(d);
    for (int b = 0; b < B; ++b) *reinterpret_cast<bf16x8*>(dfeat + (size_t)b * n + i * 8) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Adam (coupled L2 weight decay, torch.optim.Adam op order), flat fp32 buffers
// ------------------------------------------------------------------------------------------------
// state = {step count, 1 - beta1^step, sqrt(1 - beta2^step), unused}: kept on the device so a captured CUDA graph
// of the train step advances the bias corrections on every replay.
__global__ void adam_tick_kernel(float* __restrict__ state, float b1, float b2) {
  griddep_launch_dependents();
  griddep_wait();
  const double step = (double)state[0] + 1.0;
  state[0] = (float)step;
  state[1] = (float)(1.0 - pow((double)b1, step));
  state[2] = (float)sqrt(1.0 - pow((double)b2, step));
}

__global__ void __launch_bounds__(256)
adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
            long long n4, float lr, float b1, float b2, float eps, float wd, const float* __restrict__ state,
            float grad_scale) {
  griddep_launch_dependents();
  griddep_wait();
  const float step_size = lr / state[1];
  const float bc2_sqrt = state[2];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pi = p[i], mi = m[i], vi = v[i];
    const float4 gr = g[i];
    float* pp = reinterpret_cast<float*>(&pi);
    float* mp = reinterpret_cast<float*>(&mi);
    float* vp = reinterpret_cast<float*>(&vi);
    const float* gp = reinterpret_cast<const float*>(&gr);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gi = gp[k] * grad_scale + wd * pp[k];
      mp[k] = b1 * mp[k] + (1.f - b1) * gi;        // exp_avg.lerp_(grad, 1-beta1)
      vp[k] = b2 * vp[k] + (1.f - b2) * gi * gi;   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
      const float denom = sqrtf(vp[k]) / bc2_sqrt + eps;
      pp[k] = pp[k] - step_size * (mp[k] / denom);
    }
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
}

int ew_grid(long long work_items, int sms) {
  long long blocks = (work_items + 255) / 256;
  const long long cap = (long long)sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int sms() {
  static int n_dev[kMaxDevices] = {};
  const int dev = current_device();
  int& n = n_dev[dev];
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

extern "C" {

const char* dg_last_error(void) { return g_err; }
int dg_version(void) { return 2; }
#ifndef DG_SOURCE_HASH
#define DG_SOURCE_HASH "unknown"
#endif
const char* dg_source_hash(void) { return DG_SOURCE_HASH; }
long long dg_launch_count(void) { return (long long)g_dg_launches.load(std::memory_order_relaxed); }

// Fails loudly unless the current device is a Blackwell sm_100 part with a loadable kernel image.
int dg_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    dg_set_error("no CUDA device: %s", cudaGetErrorString(e));
    return DG_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    dg_set_error("discogan_modernized_b200 kernels are built for sm_100a only; device is sm_%d%d", major, minor);
    return DG_ERR_ARCH;
  }
  return DG_OK;
}

int dg_pack_weights(const float* w, void* wd, void* wu, int Cs, int Cb, cudaStream_t stream) {
  DG_CHECK_ARG(Cs > 0 && Cb > 0 && w, "pack_weights: bad args");
  const long long n = (long long)Cs * Cb;
  if (wd) {
    dg_launch(pack_wd_kernel, dg_cfg(dg_ceil_div(n, 256), 256, 0, stream), w, (bf16*)wd, Cs, Cb);
    DG_CHECK_LAUNCH("pack_wd");
  }
  if (wu) {
    dg_launch(pack_wu_kernel, dg_cfg(dg_ceil_div(n, 256), 256, 0, stream), w, (bf16*)wu, Cs, Cb);
    DG_CHECK_LAUNCH("pack_wu");
  }
  return DG_OK;
}

// table: device int64 [n][6] = {w ptr, wd ptr (or 0), wu ptr (or 0), Cs, Cb, running end of 2*Cs*Cb items}; n <= 64
int dg_pack_weights_multi(const long long* table, int n, long long total_items, cudaStream_t stream) {
  DG_CHECK_ARG(table && n > 0 && n <= 64 && total_items > 0, "pack_weights_multi: bad args");
  dg_launch(pack_multi_kernel, dg_cfg(ew_grid(total_items, sms()), 256, 0, stream), table, n, total_items);
  DG_CHECK_LAUNCH("pack_weights_multi");
  return DG_OK;
}

int dg_nhwc_bf16_to_nchw_f32(const void* x, float* y, int B, int HW, int C, cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && HW > 0 && C > 0 && B <= 65535, "nhwc_to_nchw: bad dims");
  dim3 grid(dg_ceil_div(HW, 32), dg_ceil_div(C, 32), B), block(32, 8);
  dg_launch(nhwc_to_nchw_kernel, dg_cfg(grid, block, 0, stream), (const bf16*)x, y, HW, C);
  DG_CHECK_LAUNCH("nhwc_to_nchw");
  return DG_OK;
}
int dg_nchw_f32_to_nhwc_bf16(const float* x, void* y, int B, int HW, int C, cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && HW > 0 && C > 0 && B <= 65535, "nchw_to_nhwc: bad dims");
  dim3 grid(dg_ceil_div(HW, 32), dg_ceil_div(C, 32), B), block(32, 8);
  dg_launch(nchw_to_nhwc_kernel, dg_cfg(grid, block, 0, stream), x, (bf16*)y, HW, C);
  DG_CHECK_LAUNCH("nchw_to_nhwc");
  return DG_OK;
}

// scratch layout for BN calls: floats [2 * splits_max * C]; query with dg_bn_scratch_floats
size_t dg_bn_scratch_floats(long long P, int C) {
  const int vec = (C % 8 == 0) ? 8 : 1;
  BnGeom g = bn_geom(P, C, vec, sms());
  const int rows = g.gy > 3 * sms() ? g.gy : 3 * sms();   // the streaming backward writes one partial row per block
  return (size_t)2 * rows * C;
}

// Training-mode statistics of z[P][C]: mean, invstd, apply coefficients (scale, shift) and the running-stat update
// (running_* may be NULL).  stats = float[4*C] = {mean, invstd, scale, shift}.
int dg_bn_stats(const void* z, long long P, int C, const float* gamma, const float* beta, float eps, float momentum,
                float* stats, float* running_mean, float* running_var, float* scratch, cudaStream_t stream) {
  DG_CHECK_ARG(P > 0 && C > 0 && z && stats && scratch, "bn_stats: bad args");
  const int vec = (C % 8 == 0) ? 8 : 1;
  BnGeom g = bn_geom(P, C, vec, sms());
  float* ps = scratch;
  float* pq = scratch + (size_t)g.gy * C;
  dim3 grid(g.gx, g.gy);
  if (vec == 8)
    dg_launch(bn_stats_partial_kernel<8>, dg_cfg(grid, 256, 0, stream), (const bf16*)z, P, C, g.cw, g.rows_iter, g.rows_split, ps, pq,
              (float*)nullptr);
  else
    dg_launch(bn_stats_partial_kernel<1>, dg_cfg(grid, 256, 0, stream), (const bf16*)z, P, C, g.cw, g.rows_iter, g.rows_split, ps, pq,
              (float*)nullptr);
  DG_CHECK_LAUNCH("bn_stats_partial");
  dg_launch(bn_stats_finalize_kernel, dg_cfg(dg_ceil_div(C, 32), 256, 0, stream), ps, pq, g.gy, P, C, eps, momentum, gamma, beta,
                                                                    stats, stats + C, stats + 2 * C, stats + 3 * C,
                                                                    running_mean, running_var);
  DG_CHECK_LAUNCH("bn_stats_finalize");
  return DG_OK;
}

// Finish statistics whose partial sums were produced elsewhere (the fused conv epilogue): part = float[2*rows*C]
// {sum rows, sum-of-squares rows}.  Same outputs / running-stat update as dg_bn_stats.
int dg_bn_stats_finalize(const float* part, int rows, long long P, int C, const float* gamma, const float* beta,
                         float eps, float momentum, float* stats, float* running_mean, float* running_var,
                         cudaStream_t stream) {
  DG_CHECK_ARG(part && rows > 0 && P > 0 && C > 0 && stats, "bn_stats_finalize: bad args");
  dg_launch(bn_stats_finalize_kernel, dg_cfg(dg_ceil_div(C, 32), 256, 0, stream), part, part + (size_t)rows * C, rows, P, C, eps,
                                                                   momentum, gamma, beta, stats, stats + C,
                                                                   stats + 2 * C, stats + 3 * C, running_mean,
                                                                   running_var);
  DG_CHECK_LAUNCH("bn_stats_finalize");
  return DG_OK;
}

int dg_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                      float eps, int C, float* stats, cudaStream_t stream) {
  DG_CHECK_ARG(C > 0 && stats, "bn_eval_coeffs: bad args");
  dg_launch(bn_eval_coeffs_kernel, dg_cfg(dg_ceil_div(C, 128), 128, 0, stream), gamma, beta, running_mean, running_var, eps, C,
                                                                 stats + 2 * C, stats + 3 * C);
  DG_CHECK_LAUNCH("bn_eval_coeffs");
  return DG_OK;
}

// sums of z[P][C] added into the zero-initialised accumulator pair acc[2][C] (folded-finalize producer for layers whose
// convolution could not fuse the statistics: the 4x4 valid heads)
int dg_bn_stats_acc(const void* z, long long P, int C, float* acc, cudaStream_t stream) {
  DG_CHECK_ARG(P > 0 && C > 0 && z && acc, "bn_stats_acc: bad args");
  const int vec = (C % 8 == 0) ? 8 : 1;
  BnGeom g = bn_geom(P, C, vec, sms());
  dim3 grid(g.gx, g.gy);
  if (vec == 8)
    dg_launch(bn_stats_partial_kernel<8>, dg_cfg(grid, 256, 0, stream), (const bf16*)z, P, C, g.cw, g.rows_iter, g.rows_split,
              (float*)nullptr, (float*)nullptr, acc);
  else
    dg_launch(bn_stats_partial_kernel<1>, dg_cfg(grid, 256, 0, stream), (const bf16*)z, P, C, g.cw, g.rows_iter, g.rows_split,
              (float*)nullptr, (float*)nullptr, acc);
  DG_CHECK_LAUNCH("bn_stats_partial");
  return DG_OK;
}

static int bn_act_fwd_impl(const void* z, void* y, long long P, int C, const float* stats, int act, float slope,
                           const BnFold& fold, cudaStream_t stream);

// y = act(z * scale + shift)
int dg_bn_act_fwd(const void* z, void* y, long long P, int C, const float* stats, int act, float slope,
                  cudaStream_t stream) {
  DG_CHECK_ARG(P > 0 && C > 0 && z && y && stats, "bn_act_fwd: bad args");
  BnFold fold = {};
  return bn_act_fwd_impl(z, y, P, C, stats, act, slope, fold, stream);
}

// y = act(BN_train(z)) with the statistics taken from accumulated sums acc[2][C] = {sum z, sum z^2} (filled by the
// producing convolution's epilogue in accumulator mode, or by dg_bn_stats_acc): no finalize launch.  Also writes
// stats[4][C] = {mean, invstd, scale, shift} for the backward pass and updates the running statistics (may be NULL).
int dg_bn_act_fwd_acc(const void* z, void* y, long long P, int C, const float* acc, const float* gamma, const float* beta,
                      float eps, float momentum, float* stats, float* running_mean, float* running_var, int act,
                      float slope, cudaStream_t stream) {
  DG_CHECK_ARG(P > 0 && C > 0 && z && y && acc && gamma && beta && stats, "bn_act_fwd_acc: bad args");
  BnFold fold = {};
  fold.acc = acc;
  fold.gamma = gamma;
  fold.beta = beta;
  fold.stats_out = stats;
  fold.running_mean = running_mean;
  fold.running_var = running_mean ? running_var : nullptr;
  fold.eps = eps;
  fold.momentum = momentum;
  fold.inv_P = 1.0 / (double)P;
  fold.unbias = P > 1 ? (double)P / (double)(P - 1) : 1.0;
  return bn_act_fwd_impl(z, y, P, C, stats, act, slope, fold, stream);
}

static int bn_act_fwd_impl(const void* z, void* y, long long P, int C, const float* stats, int act, float slope,
                           const BnFold& fold, cudaStream_t stream) {
  BnStreamPlan pl;
  if (((uintptr_t)z & 15) == 0 && ((uintptr_t)y & 15) == 0 && bn_stream_plan(P, C, 1, sms(), &pl)) {
    BnStreamParams sp = {};
    sp.in[0] = (const bf16*)z;
    sp.nstreams = 1;
    sp.out = (bf16*)y;
    sp.stats = stats;
    sp.P = P;
    sp.C = C;
    sp.rows_tile = pl.rows_tile;
    sp.tiles = pl.tiles;
    sp.tiles_per_block = pl.tiles_per_block;
    sp.act = act;
    sp.slope = slope;
    sp.fold = fold;
    return bn_stream_launch<BS_FWD>(sp, pl, stream);
  }
  const int vec = (C % 8 == 0) ? 8 : 1;
  BnGeom g = bn_geom(P, C, vec, sms(), 8);
  dim3 grid(g.gx, g.gy);
  if (vec == 8)
    dg_launch(bn_act_fwd_kernel<8>, dg_cfg(grid, 256, 0, stream), (const bf16*)z, (bf16*)y, P, C, g.cw, g.rows_iter, g.rows_split,
                                                   stats + 2 * C, stats + 3 * C, act, slope, fold);
  else
    dg_launch(bn_act_fwd_kernel<1>, dg_cfg(grid, 256, 0, stream), (const bf16*)z, (bf16*)y, P, C, g.cw, g.rows_iter, g.rows_split,
                                                   stats + 2 * C, stats + 3 * C, act, slope, fold);
  DG_CHECK_LAUNCH("bn_act_fwd");
  return DG_OK;
}

// Backward of y = act(BN_train(z)).  Upstream gradient g = dy (+ dy2) (+ bcast_coef * bcast[row % bcast_rows]).
// dgamma/dbeta: fp32 grads, new = grad_beta*old + value (grad_beta 0 overwrites).  coefs: float[3*C] scratch.
static int bn_act_bwd_impl(const void* dy, const void* dy2, const float* bcast, float bcast_coef, long long bcast_rows,
                           const void* z, const float* stats, const float* gamma, long long P, int C, int act,
                           float slope, float* dgamma, float* dbeta, float grad_beta, void* dz, float* coefs,
                           float* scratch, float* acc2, cudaStream_t stream);

int dg_bn_act_bwd(const void* dy, const void* dy2, const float* bcast, float bcast_coef, long long bcast_rows,
                  const void* y, const void* z, const float* stats, const float* gamma, long long P, int C, int act,
                  float slope, float* dgamma, float* dbeta, float grad_beta, void* dz, float* coefs, float* scratch,
                  cudaStream_t stream) {
  (void)y;  // the activation derivative is recomputed from z and the statistics; y is accepted for API symmetry
  DG_CHECK_ARG(P > 0 && C > 0 && dy && z && stats && dz && coefs && scratch, "bn_act_bwd: bad args");
  DG_CHECK_ARG(!bcast || bcast_rows > 0, "bn_act_bwd: bcast_rows must be positive");
  return bn_act_bwd_impl(dy, dy2, bcast, bcast_coef, bcast_rows, z, stats, gamma, P, C, act, slope, dgamma, dbeta, grad_beta,
                         dz, coefs, scratch, nullptr, stream);
}

// Same backward in two launches instead of three: the reduction adds {sum g', sum g'*xhat} into the zero-initialised
// accumulator pair acc2[2][C], the dx kernel derives its coefficients from it and writes dgamma / dbeta.
int dg_bn_act_bwd_acc(const void* dy, const void* dy2, const float* bcast, float bcast_coef, long long bcast_rows,
                      const void* z, const float* stats, const float* gamma, long long P, int C, int act, float slope,
                      float* dgamma, float* dbeta, float grad_beta, void* dz, float* acc2, cudaStream_t stream) {
  DG_CHECK_ARG(P > 0 && C > 0 && dy && z && stats && dz && acc2 && gamma, "bn_act_bwd_acc: bad args");
  DG_CHECK_ARG(!bcast || bcast_rows > 0, "bn_act_bwd_acc: bcast_rows must be positive");
  DG_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), "bn_act_bwd_acc: dgamma and dbeta go together");
  return bn_act_bwd_impl(dy, dy2, bcast, bcast_coef, bcast_rows, z, stats, gamma, P, C, act, slope, dgamma, dbeta, grad_beta,
                         dz, nullptr, nullptr, acc2, stream);
}

static int bn_act_bwd_impl(const void* dy, const void* dy2, const float* bcast, float bcast_coef, long long bcast_rows,
                           const void* z, const float* stats, const float* gamma, long long P, int C, int act,
                           float slope, float* dgamma, float* dbeta, float grad_beta, void* dz, float* coefs,
                           float* scratch, float* acc2, cudaStream_t stream) {
  BnFold fold = {};
  if (acc2) {
    fold.gamma = gamma;
    fold.dgamma = dgamma;
    fold.dbeta = dbeta;
    fold.grad_beta = grad_beta;
    fold.inv_P = 1.0 / (double)P;
  }
  BnStreamPlan pl;
  if ((((uintptr_t)dy | (uintptr_t)dy2 | (uintptr_t)z | (uintptr_t)dz | (uintptr_t)bcast) & 15) == 0 &&
      bn_stream_plan(P, C, dy2 ? 3 : 2, sms(), &pl)) {
    BnStreamParams sp = {};
    sp.in[0] = (const bf16*)dy;
    sp.in[1] = (const bf16*)z;
    sp.in[2] = (const bf16*)dy2;
    sp.nstreams = dy2 ? 3 : 2;
    sp.bcast = bcast;
    sp.bcast_coef = bcast_coef;
    sp.bcast_rows = bcast_rows;
    sp.stats = stats;
    sp.coefs = coefs;
    sp.part0 = scratch;
    sp.part1 = scratch ? scratch + (size_t)pl.grid * C : nullptr;
    sp.P = P;
    sp.C = C;
    sp.rows_tile = pl.rows_tile;
    sp.tiles = pl.tiles;
    sp.tiles_per_block = pl.tiles_per_block;
    sp.act = act;
    sp.slope = slope;
    sp.fold = fold;
    sp.fold.acc_out = acc2;
    int rc = bn_stream_launch<BS_BWD_PARTIAL>(sp, pl, stream);
    if (rc) return rc;
    if (!acc2) {
      dg_launch(bn_bwd_finalize_kernel, dg_cfg(dg_ceil_div(C, 32), 256, 0, stream), (const float*)sp.part0,
                (const float*)sp.part1, pl.grid, P, C, gamma, stats + C, dgamma, dbeta, grad_beta, coefs);
      DG_CHECK_LAUNCH("bn_bwd_finalize");
    }
    sp.fold.acc_out = nullptr;
    sp.fold.acc = acc2;
    sp.out = (bf16*)dz;
    return bn_stream_launch<BS_BWD_DX>(sp, pl, stream);
  }
  const int vec = (C % 8 == 0) ? 8 : 1;
  BnGeom g = bn_geom(P, C, vec, sms());
  float* pg = scratch;
  float* pgx = scratch ? scratch + (size_t)g.gy * C : nullptr;
  dim3 grid(g.gx, g.gy);
  const float* invstd = stats + C;
  BnFold fold_dx = fold;
  fold_dx.acc = acc2;
  if (vec == 8)
    dg_launch(bn_bwd_partial_kernel<8>, dg_cfg(grid, 256, 0, stream), (const bf16*)dy, (const bf16*)dy2, bcast, bcast_coef, bcast_rows,
                                                       (const bf16*)z, stats, P, C, g.cw, g.rows_iter, g.rows_split, act,
                                                       slope, pg, pgx, acc2);
  else
    dg_launch(bn_bwd_partial_kernel<1>, dg_cfg(grid, 256, 0, stream), (const bf16*)dy, (const bf16*)dy2, bcast, bcast_coef, bcast_rows,
                                                       (const bf16*)z, stats, P, C, g.cw, g.rows_iter, g.rows_split, act,
                                                       slope, pg, pgx, acc2);
  DG_CHECK_LAUNCH("bn_bwd_partial");
  if (!acc2) {
    dg_launch(bn_bwd_finalize_kernel, dg_cfg(dg_ceil_div(C, 32), 256, 0, stream), pg, pgx, g.gy, P, C, gamma, invstd, dgamma, dbeta,
                                                                   grad_beta, coefs);
    DG_CHECK_LAUNCH("bn_bwd_finalize");
  }
  BnGeom g2 = bn_geom(P, C, vec, sms(), 8);
  dim3 grid2(g2.gx, g2.gy);
  if (vec == 8)
    dg_launch(bn_bwd_dx_kernel<8>, dg_cfg(grid2, 256, 0, stream), (const bf16*)dy, (const bf16*)dy2, bcast, bcast_coef, bcast_rows,
                                                   (const bf16*)z, stats, coefs, (bf16*)dz, P, C, g2.cw, g2.rows_iter,
                                                   g2.rows_split, act, slope, fold_dx);
  else
    dg_launch(bn_bwd_dx_kernel<1>, dg_cfg(grid2, 256, 0, stream), (const bf16*)dy, (const bf16*)dy2, bcast, bcast_coef, bcast_rows,
                                                   (const bf16*)z, stats, coefs, (bf16*)dz, P, C, g2.cw, g2.rows_iter,
                                                   g2.rows_split, act, slope, fold_dx);
  DG_CHECK_LAUNCH("bn_bwd_dx");
  return DG_OK;
}

int dg_gan_bce_fwd(const float* logit_real, const float* logit_fake, int B, float* p_real, float* p_fake, float* out2,
                   cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && logit_real && logit_fake && p_real && p_fake && out2, "gan_bce_fwd: bad args");
  dg_launch(gan_bce_fwd_kernel, dg_cfg(1, 256, 0, stream), logit_real, logit_fake, B, p_real, p_fake, out2);
  DG_CHECK_LAUNCH("gan_bce_fwd");
  return DG_OK;
}
int dg_gan_bce_bwd(const float* p_real, const float* p_fake, int B, float g_dis, float g_gen, float* dlogit_real,
                   float* dlogit_fake, cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && p_real && p_fake, "gan_bce_bwd: bad args");
  dg_launch(gan_bce_bwd_kernel, dg_cfg(dg_ceil_div(B, 128), 128, 0, stream), p_real, p_fake, B, g_dis, g_gen, dlogit_real,
                                                              dlogit_fake);
  DG_CHECK_LAUNCH("gan_bce_bwd");
  return DG_OK;
}
int dg_sigmoid_fwd(const float* x, float* y, int n, cudaStream_t stream) {
  DG_CHECK_ARG(n > 0, "sigmoid_fwd: bad args");
  dg_launch(sigmoid_fwd_kernel, dg_cfg(dg_ceil_div(n, 128), 128, 0, stream), x, y, n);
  DG_CHECK_LAUNCH("sigmoid_fwd");
  return DG_OK;
}
int dg_sigmoid_bwd(const float* y, const float* dy, float* dx, int n, cudaStream_t stream) {
  DG_CHECK_ARG(n > 0, "sigmoid_bwd: bad args");
  dg_launch(sigmoid_bwd_kernel, dg_cfg(dg_ceil_div(n, 128), 128, 0, stream), y, dy, dx, n);
  DG_CHECK_LAUNCH("sigmoid_bwd");
  return DG_OK;
}

// out[0] = mean((a-b)^2); scratch: float[dg_reduce_scratch_floats()]
size_t dg_reduce_scratch_floats(void) { return (size_t)sms() * 8; }

int dg_mse_fwd(const float* a, const float* b, long long n, float* out, float* scratch, cudaStream_t stream) {
  DG_CHECK_ARG(n > 0 && a && b && out && scratch, "mse_fwd: bad args");
  DG_CHECK_ARG(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0, "mse_fwd: inputs must be 16-byte aligned");
  const int grid = ew_grid(n / 4 + 1, sms());
  dg_launch(mse_partial_kernel, dg_cfg(grid, 256, 0, stream), a, b, n, scratch);
  DG_CHECK_LAUNCH("mse_partial");
  dg_launch(sum_finalize_kernel, dg_cfg(1, 32, 0, stream), scratch, grid, 1.0 / (double)n, out, 0);
  DG_CHECK_LAUNCH("mse_finalize");
  return DG_OK;
}
// da (+)= g * 2/n * (a-b)
int dg_mse_bwd(const float* a, const float* b, long long n, float g, float* da, int accumulate, cudaStream_t stream) {
  DG_CHECK_ARG(n > 0 && a && b && da, "mse_bwd: bad args");
  dg_launch(mse_bwd_kernel, dg_cfg(ew_grid(n, sms()), 256, 0, stream), a, b, n, g * 2.f / (float)n, da, accumulate);
  DG_CHECK_LAUNCH("mse_bwd");
  return DG_OK;
}

// out[0] (+)= mean_i (mean_b real - mean_b fake)^2 over n = H*W*C; diff (fp32 [n]) optional, kept for the backward.
int dg_fm_fwd(const void* real, const void* fake, int B, long long n, float* diff, float* out, int accumulate,
              float* scratch, cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && n > 0 && n % 8 == 0 && real && fake && out && scratch, "fm_fwd: bad args (n must be a multiple of 8)");
  long long blocks = (n / 8 + 31) / 32;
  if (blocks > (long long)sms() * 8) blocks = (long long)sms() * 8;      // = dg_reduce_scratch_floats()
  const int grid = (int)blocks;
  dg_launch(fm_partial_kernel, dg_cfg(grid, 256, 0, stream), (const bf16*)real, (const bf16*)fake, B, n, diff, scratch);
  DG_CHECK_LAUNCH("fm_partial");
  dg_launch(sum_finalize_kernel, dg_cfg(1, 32, 0, stream), scratch, grid, 1.0 / (double)n, out, accumulate);
  DG_CHECK_LAUNCH("fm_finalize");
  return DG_OK;
}
// dfeat_fake[b][i] = -g * 2/(n*B) * diff[i]   (use +g for the real branch)
int dg_fm_bwd(const float* diff, int B, long long n, float g, void* dfeat, cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && n > 0 && n % 8 == 0 && diff && dfeat, "fm_bwd: bad args");
  dg_launch(fm_bwd_kernel, dg_cfg(ew_grid(n / 8, sms()), 256, 0, stream), diff, B, n, -g * 2.f / ((float)n * (float)B), (bf16*)dfeat);
  DG_CHECK_LAUNCH("fm_bwd");
  return DG_OK;
}

// One Adam step over flat fp32 buffers (n a multiple of 4, 16-byte aligned).  `state` is a device float[4]
// {steps taken, bias corrections}; zero it once, every call advances it.
// The two halves of dg_adam_step, for callers that update a parameter buffer piecewise (one call per gradient bucket as its
// all-reduce completes): dg_adam_tick advances the step count / bias corrections ONCE per optimiser step, dg_adam_apply
// updates a range with the current state.
int dg_adam_tick(float* state, float beta1, float beta2, cudaStream_t stream) {
  DG_CHECK_ARG(state != nullptr, "adam_tick: bad args");
  dg_launch(adam_tick_kernel, dg_cfg(1, 1, 0, stream), state, beta1, beta2);
  DG_CHECK_LAUNCH("adam_tick");
  return DG_OK;
}
int dg_adam_apply(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, const float* state, float grad_scale, cudaStream_t stream) {
  DG_CHECK_ARG(n > 0 && n % 4 == 0 && p && g && m && v && state, "adam_apply: bad args (n must be a multiple of 4)");
  DG_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adam_apply: unaligned buffers");
  dg_launch(adam_kernel, dg_cfg(ew_grid(n / 4, sms()), 256, 0, stream), (float4*)p, (const float4*)g, (float4*)m, (float4*)v, n / 4, lr,
                                                         beta1, beta2, eps, weight_decay, state, grad_scale);
  DG_CHECK_LAUNCH("adam_apply");
  return DG_OK;
}

int dg_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                 float eps, float weight_decay, float* state, float grad_scale, cudaStream_t stream) {
  DG_CHECK_ARG(n > 0 && n % 4 == 0 && p && g && m && v && state, "adam_step: bad args (n must be a multiple of 4)");
  DG_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adam_step: unaligned buffers");
  dg_launch(adam_tick_kernel, dg_cfg(1, 1, 0, stream), state, beta1, beta2);
  DG_CHECK_LAUNCH("adam_tick");
  dg_launch(adam_kernel, dg_cfg(ew_grid(n / 4, sms()), 256, 0, stream), (float4*)p, (const float4*)g, (float4*)m, (float4*)v, n / 4, lr,
                                                         beta1, beta2, eps, weight_decay, state, grad_scale);
  DG_CHECK_LAUNCH("adam_step");
  return DG_OK;
}

}  // extern "C"
