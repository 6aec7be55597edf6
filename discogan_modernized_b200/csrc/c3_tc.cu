// Tensor-core versions of the image-side 3-channel layers (sm_100a, tcgen05 + TMA).
//
// Conv2d(3,64,4,2,1) (model.py:8,80) has K = 48 and ConvTranspose2d(64,3,4,2,1) (model.py:142) has N = 3: on CUDA
// cores they were 10x off the HBM roofline (profiles/r01_launches_512x512_b32_eager.csv).  Here the fp32 NCHW image is
// first repacked once into a zero-padded NHWC4 bf16 image [B,S+2,S+2,4] (8 B per pixel); then
//   * "down" (conv1 forward / dgrad of the final ConvTranspose2d): for output pixel (ho,wo) and kernel row kh the
//     4 kernel columns x 4 channels are 16 contiguous bf16 = 32 bytes of the padded image, so a TMA map with
//     OVERLAPPING windows {16, So, 4, So, B} (pixel stride 8 elements) delivers, per kh, a [128 pixels x 32 B]
//     SWIZZLE_32B K-major slab = exactly one UMMA K=16 step.  GEMM M = pixels, N = 64, K = 4 x 16; the 8 KB weight
//     stays resident in shared memory; persistent CTAs, 4 TMEM accumulator stages.
//   * wgrad of both layers: dW[c64][k] = sum_pixels V[p][c64] * patch[p][k]: V (NHWC 64-channel tensor) is the
//     MN-major SWIZZLE_128B A operand, the same patch slabs are the MN-major SWIZZLE_32B B operand; split-K over
//     pixel chunks, partial 64x64 tiles reduced by a second kernel into the PyTorch layout [64][3][4][4].
// The "up" direction (final ConvTranspose2d forward / conv1 dgrad) runs on conv_gemm_kernel with N padded to 16
// (gemm_tc.cu, image epilogue).
#include <cuda.h>

#include "common.cuh"
#include "tma_host.cuh"

namespace {

constexpr int kThreads = 192;
constexpr uint32_t kSw32 = 6;  // UMMA layout type SWIZZLE_32B

// ------------------------------------------------------------------------------------------------
// fp32 NCHW image (optionally * y(1-y)) -> zero-padded NHWC4 bf16 [B][S+2][S+2][4]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
img_pad_nhwc4_kernel(const float* __restrict__ img, const float* __restrict__ img2, const float* __restrict__ yimg,
                     bf16* __restrict__ out, int B, int S) {
  griddep_launch_dependents();
  griddep_wait();
  const int Sp = S + 2;
  const long long total = (long long)B * Sp * Sp;
  const size_t plane = (size_t)S * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xp = (int)(i % Sp);
    const int yp = (int)((i / Sp) % Sp);
    const int b = (int)(i / ((long long)Sp * Sp));
    float v[3] = {0.f, 0.f, 0.f};
    if (xp >= 1 && xp <= S && yp >= 1 && yp <= S) {
      const size_t off = (size_t)b * 3 * plane + (size_t)(yp - 1) * S + (xp - 1);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float t = img[off + c * plane];
        if (img2) t += img2[off + c * plane];      // two gradient contributions summed on the fly
        if (yimg) {
          const float s = yimg[off + c * plane];
          t *= s * (1.f - s);
        }
        v[c] = t;
      }
    }
    uint2 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], 0.f);
    *reinterpret_cast<uint2*>(out + i * 4) = o;
  }
}

// w fp32 [64][3][4][4] -> wc bf16 [64][kh*16 + kw*4 + c] (c = 3 zero) and wu3 bf16 [16][9 shifts][64]: the "up" GEMM
// computes all four output parities of a pixel at once -- row n = parity * 3 + c (parity = py * 2 + px; rows 12..15 zero),
// K = (pixel shift (di, dj) in {-1,0,1}^2) x 64 input channels; an entry is W[cs][c][kh][kw] of the tap that parity uses
// at that shift (py = 0: di 0 -> kh 1, di -1 -> kh 3; py = 1: di +1 -> kh 0, di 0 -> kh 2; same for px / kw) or zero.
// 9 pixel-tile loads and MMA groups per tile instead of 4 parities x 4 taps = 16.
__global__ void c3_pack_weights_kernel(const float* __restrict__ w, bf16* __restrict__ wc, bf16* __restrict__ wu3) {
  griddep_launch_dependents();
  griddep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (wc && i < 64 * 64) {
    const int c64 = i >> 6, k = i & 63;
    const int kh = k >> 4, kw = (k >> 2) & 3, c = k & 3;
    wc[i] = __float2bfloat16(c < 3 ? w[c64 * 48 + c * 16 + kh * 4 + kw] : 0.f);
  }
  if (wu3 && i < 16 * 16 * 64) {
    float v = 0.f;
    if (i < 16 * 9 * 64) {
      const int cs = i & 63, shift = (i >> 6) % 9, n = i / (9 * 64);
      const int par = n / 3, c = n - par * 3;
      if (par < 4) {
        const int py = par >> 1, px = par & 1, di = shift / 3 - 1, dj = shift % 3 - 1;
        const int kh = py == 0 ? (di == 0 ? 1 : (di == -1 ? 3 : -1)) : (di == 1 ? 0 : (di == 0 ? 2 : -1));
        const int kw = px == 0 ? (dj == 0 ? 1 : (dj == -1 ? 3 : -1)) : (dj == 1 ? 0 : (dj == 0 ? 2 : -1));
        if (kh >= 0 && kw >= 0) v = w[cs * 48 + c * 16 + kh * 4 + kw];
      }
    }
    wu3[i] = __float2bfloat16(v);
  }
}

// ------------------------------------------------------------------------------------------------
// down: out[pix][64] = act(sum_k patch[pix][k] * wc[c64][k])
// ------------------------------------------------------------------------------------------------
constexpr int kDnStages = 6;
constexpr int kDnSlab = 128 * 32;         // one kernel row: 128 pixels x 32 B
constexpr int kDnStageBytes = 4 * kDnSlab;  // 16 KB
constexpr int kDnWBytes = 4 * 64 * 32;    // resident weights: 4 slabs of [64 rows x 32 B]
constexpr int kDnAcc = 4;                 // TMEM accumulator stages of 64 columns

struct C3DownParams {
  int B, So, Wt, Ht, Bt, tiles_w, tiles_h, tiles_b, num_tiles, act;
  float slope;
  bf16* out;
};

__global__ void __launch_bounds__(kThreads, 1)
c3_down_tc_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmW,
                  const C3DownParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sw = smem;
  uint8_t* sst = smem + kDnWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sst + kDnStages * kDnStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kDnStages;
  uint64_t* tfull_bar = bars + 2 * kDnStages;
  uint64_t* tempty_bar = tfull_bar + kDnAcc;
  uint64_t* w_bar = tempty_bar + kDnAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmP);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < kDnStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kDnAcc; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // programmatic dependent launch: dependents may be scheduled once every CTA of this grid holds its TMEM columns;
  // nothing above touches global memory, everything below runs after the predecessor grid has completed
  griddep_launch_dependents();
  griddep_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_bar, kDnWBytes);
      for (int kh = 0; kh < 4; ++kh) tma_load_2d(sw + kh * 2048, &tmW, w_bar, kh * 16, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int r = tile;
        const int w0 = (r % p.tiles_w) * p.Wt;
        r /= p.tiles_w;
        const int h0 = (r % p.tiles_h) * p.Ht;
        const int b0 = (r / p.tiles_h) * p.Bt;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&full_bar[stage], kDnStageBytes);
        uint8_t* sa = sst + stage * kDnStageBytes;
        for (int kh = 0; kh < 4; ++kh) tma_load_5d(sa + kh * kDnSlab, &tmP, &full_bar[stage], 0, w0, kh, h0, b0);
        if (++stage == kDnStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      mbar_wait(w_bar, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(sst + stage * kDnStageBytes);
        const uint32_t sb = smem_u32(sw);
#pragma unroll
        for (int kh = 0; kh < 4; ++kh) {
          const uint64_t da = make_sdesc(sa + kh * kDnSlab, 16, 256, kSw32);
          const uint64_t db = make_sdesc(sb + kh * 2048, 16, 256, kSw32);
          umma_bf16(tmem_base + (uint32_t)(acc * 64), da, db, idesc, kh > 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tfull_bar[acc]);
        if (++stage == kDnStages) {
          stage = 0;
          phase ^= 1u;
        }
        if (++acc == kDnAcc) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int r = tile;
      const int w0 = (r % p.tiles_w) * p.Wt;
      r /= p.tiles_w;
      const int h0 = (r % p.tiles_h) * p.Ht;
      const int b0 = (r / p.tiles_h) * p.Bt;
      const int wl = row % p.Wt, hl = (row / p.Wt) % p.Ht, bl = row / (p.Wt * p.Ht);
      const int b = b0 + bl;
      bf16* orow = p.out + (((size_t)b * p.So + h0 + hl) * p.So + w0 + wl) * 64;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 64);
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t rr[32];
        tmem_ld_32x32(taddr + c, rr);
        tmem_ld_wait();
        if (b < p.B) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = act_fwd(__uint_as_float(rr[8 * j + e]), p.act, p.slope);
            *reinterpret_cast<bf16x8*>(orow + c + 8 * j) = pack8(f);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == kDnAcc) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad: ws[split][c64][k] = sum over the split's pixel chunks of V[p][c64] * patch[p][k]
// ------------------------------------------------------------------------------------------------
constexpr int kWgStages = 8;
constexpr int kWgA = 2 * 32 * 128;   // A: 32 pixels x 64 channels (4 KB) + an unused second 64-row block (M padded to 128)
constexpr int kWgB = 4 * 32 * 32;    // B: 4 kernel-row slabs of [32 pixels x 32 B]
constexpr int kWgStageBytes = kWgA + kWgB;  // 12 KB

struct C3WgradParams {
  int B, So, Wt, Ht, Bt, chunks_w, chunks_h, chunks_b, total_chunks, splits, chunks_per_split;
  float* ws;  // [split][64][64]
};

__global__ void __launch_bounds__(kThreads, 1)
c3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmP,
                   const C3WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kWgStages;
  uint64_t* tfull_bar = bars + 2 * kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x;
  const int chunk_begin = split * p.chunks_per_split;
  const int chunk_end = min(p.total_chunks, chunk_begin + p.chunks_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmP);
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // programmatic dependent launch: dependents may be scheduled once every CTA of this grid holds its TMEM columns;
  // nothing above touches global memory, everything below runs after the predecessor grid has completed
  griddep_launch_dependents();
  griddep_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ch = chunk_begin; ch < chunk_end; ++ch) {
        int r = ch;
        const int w0 = (r % p.chunks_w) * p.Wt;
        r /= p.chunks_w;
        const int h0 = (r % p.chunks_h) * p.Ht;
        const int b0 = (r / p.chunks_h) * p.Bt;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = smem + stage * kWgStageBytes;
        uint8_t* sb = sa + kWgA;
        mbar_arrive_expect_tx(&full_bar[stage], 4096 + kWgB);
        tma_load_4d(sa, &tmV, &full_bar[stage], 0, w0, h0, b0);
        for (int kh = 0; kh < 4; ++kh) tma_load_5d(sb + kh * 1024, &tmP, &full_bar[stage], 0, w0, kh, h0, b0);
        if (++stage == kWgStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int ch = chunk_begin; ch < chunk_end; ++ch) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kWgStageBytes);
        const uint32_t sb = sa + kWgA;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t da = make_sdesc_sw128(sa + ks * 2048, 4096, 1024);        // MN-major, 2 blocks of 64 rows
          const uint64_t db = make_sdesc(sb + ks * 512, 1024, 256, kSw32);         // MN-major, 4 blocks of 16 columns
          umma_bf16(tmem_base, da, db, idesc, (ch > chunk_begin || ks > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == kWgStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(tfull_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t rr[32];
      tmem_ld_32x32(taddr + c, rr);
      tmem_ld_wait();
      if (row < 64) {
        float* dst = p.ws + ((size_t)split * 64 + row) * 64 + c;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(dst + 4 * j) = make_uint4(rr[4 * j], rr[4 * j + 1], rr[4 * j + 2], rr[4 * j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// dw[c64][c][kh][kw] = beta*dw + sum_split ws[split][c64][kh*16 + kw*4 + c]
__global__ void c3_wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, float beta, int splits) {
  griddep_launch_dependents();
  griddep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 48) return;
  const int c64 = i / 48, r = i % 48, c = r >> 4, kh = (r >> 2) & 3, kw = r & 3;
  const int k = kh * 16 + kw * 4 + c;
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += ws[((size_t)s * 64 + c64) * 64 + k];
  dw[i] = (beta != 0.f ? beta * dw[i] : 0.f) + acc;
}

// overlapping-window view of the padded NHWC4 image: dims {16, So, 4, So, B}
int make_patch_map(CUtensorMap* m, const void* xp, int B, int S, int box_w, int box_h, int box_b) {
  const long long Sp = S + 2, So = S / 2;
  long long dims[5] = {16, So, 4, So, B};
  long long str[5] = {1, 8, Sp * 4, 2 * Sp * 4, Sp * Sp * 4};
  int box[5] = {16, box_w, 1, box_h, box_b};
  return make_map(m, xp, 5, dims, str, box, 32);
}

}  // namespace

extern "C" {

int dg_c3_pack_weights(const float* w, void* wc, void* wu3, cudaStream_t stream) {
  DG_CHECK_ARG(w && (wc || wu3), "c3_pack_weights: bad args");
  dg_launch(c3_pack_weights_kernel, dg_cfg(64, 256, 0, stream), w, (bf16*)wc, (bf16*)wu3);
  DG_CHECK_LAUNCH("c3_pack_weights");
  return DG_OK;
}

// fp32 NCHW [B,3,S,S] (+ img2 when != NULL) (times yimg*(1-yimg) when yimg != NULL) -> bf16 [B,S+2,S+2,4], zero
// border and 4th channel
int dg_img_pad_nhwc4(const float* img, const float* img2, const float* yimg, void* out, int B, int S,
                     cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && S >= 4 && img && out, "img_pad_nhwc4: bad args");
  const long long total = (long long)B * (S + 2) * (S + 2);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  dg_launch(img_pad_nhwc4_kernel, dg_cfg((int)blocks, 256, 0, stream), img, img2, yimg, (bf16*)out, B, S);
  DG_CHECK_LAUNCH("img_pad_nhwc4");
  return DG_OK;
}

// y[B,S/2,S/2,64] = act(conv4x4s2(xp, wc)); act: 0 none, 1 LeakyReLU(slope)
int dg_c3_down_tc(const void* xp, const void* wc, void* y, int B, int S, int act, float slope, cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && is_pow2(S) && S >= 8 && xp && wc && y, "c3_down_tc: bad args");
  C3DownParams p;
  p.B = B;
  p.So = S / 2;
  tile_shape(128, p.So, p.So, &p.Wt, &p.Ht, &p.Bt);
  p.tiles_w = p.So / p.Wt;
  p.tiles_h = p.So / p.Ht;
  p.tiles_b = dg_ceil_div(B, p.Bt);
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  p.act = act;
  p.slope = slope;
  p.out = (bf16*)y;
  CUtensorMap tmP, tmW;
  int rc = make_patch_map(&tmP, xp, B, S, p.Wt, p.Ht, p.Bt);
  if (rc) return rc;
  long long wd[2] = {64, 64}, wstr[2] = {1, 64};
  int wbox[2] = {16, 64};
  rc = make_map(&tmW, wc, 2, wd, wstr, wbox, 32);
  if (rc) return rc;
  const int smem_bytes = kDnWBytes + kDnStages * kDnStageBytes + 1024 + 256;
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(c3_down_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) {
      dg_set_error("c3_down_tc: cannot raise dynamic smem: %s", cudaGetErrorString(e));
      return DG_ERR_CUDA;
    }
    attr_set = true;
  }
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  dg_launch(c3_down_tc_kernel, dg_cfg(grid, kThreads, smem_bytes, stream), tmP, tmW, p);
  DG_CHECK_LAUNCH("c3_down_tc_kernel");
  return DG_OK;
}

size_t dg_c3_wgrad_workspace(int B, int S) {
  (void)B;
  (void)S;
  return (size_t)148 * 2 * 64 * 64 * sizeof(float);
}

// dw[64][3][4][4] = beta*dw + sum_pixels v64[p][c64] * patch(xp)[p][c,kh,kw]
int dg_c3_wgrad_tc(const void* v64, const void* xp, float* dw, float beta, int B, int S, void* ws, size_t ws_bytes,
                   cudaStream_t stream) {
  DG_CHECK_ARG(B > 0 && is_pow2(S) && S >= 8 && v64 && xp && dw && ws, "c3_wgrad_tc: bad args");
  C3WgradParams p;
  p.B = B;
  p.So = S / 2;
  tile_shape(32, p.So, p.So, &p.Wt, &p.Ht, &p.Bt);
  p.chunks_w = p.So / p.Wt;
  p.chunks_h = p.So / p.Ht;
  p.chunks_b = dg_ceil_div(B, p.Bt);
  p.total_chunks = p.chunks_w * p.chunks_h * p.chunks_b;
  int splits = num_sms();
  const int max_splits = p.total_chunks / 8;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.chunks_per_split = dg_ceil_div(p.total_chunks, splits);
  p.splits = dg_ceil_div(p.total_chunks, p.chunks_per_split);
  DG_CHECK_ARG(ws_bytes >= (size_t)p.splits * 64 * 64 * sizeof(float), "c3_wgrad_tc: workspace too small");
  p.ws = (float*)ws;
  CUtensorMap tmV, tmP;
  int rc = make_nhwc_map(&tmV, v64, B, p.So, p.So, 64, p.Wt, p.Ht, p.Bt);
  if (rc) return rc;
  rc = make_patch_map(&tmP, xp, B, S, p.Wt, p.Ht, p.Bt);
  if (rc) return rc;
  const int smem_bytes = kWgStages * kWgStageBytes + 1024 + 256;
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(c3_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) {
      dg_set_error("c3_wgrad_tc: cannot raise dynamic smem: %s", cudaGetErrorString(e));
      return DG_ERR_CUDA;
    }
    attr_set = true;
  }
  dg_launch(c3_wgrad_tc_kernel, dg_cfg(p.splits, kThreads, smem_bytes, stream), tmV, tmP, p);
  DG_CHECK_LAUNCH("c3_wgrad_tc_kernel");
  dg_launch(c3_wgrad_reduce_kernel, dg_cfg(12, 256, 0, stream), p.ws, dw, beta, p.splits);
  DG_CHECK_LAUNCH("c3_wgrad_reduce_kernel");
  return DG_OK;
}

}  // extern "C"
