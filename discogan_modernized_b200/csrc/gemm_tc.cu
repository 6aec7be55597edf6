// Implicit-GEMM 4x4 stride-2 convolution family on tcgen05 / TMEM, operands fed by TMA (sm_100a).
//
// Replaces the cuDNN calls behind nn.Conv2d(ci,co,4,2,1) / nn.ConvTranspose2d(ci,co,4,2,1) forward, dgrad and
// wgrad (reference model.py:11-31, 84-103, 118-138).
//
// Tensors are NHWC bf16.  A conv layer connects a "big" tensor [B,2Hs,2Ws,Cb] and a "small" tensor [B,Hs,Ws,Cs];
// the PyTorch weight seen as conv weight is W[Cs][Cb][4][4] (for ConvTranspose2d the same layout: [in=Cs][out=Cb]).
//   DOWN (conv fprop / convT dgrad):  small[b,ho,wo,cs] = sum_{kh,kw,cb} big[b,2ho-1+kh,2wo-1+kw,cb] W[cs][cb][kh][kw]
//        GEMM M = B*Hs*Ws, N = Cs, K = 16*Cb.  A tiles come straight from `big` through a 5-D TMA map that
//        splits H and W into (coarse, parity): tap (kh,kw) is a shifted box in one parity plane, zero-filled by
//        TMA outside the image (padding=1).  B = packed weights Wd[Cs][tap][Cb].
//   UP   (conv dgrad / convT fprop):  big[b,2i+py,2j+px,cb] = sum_{cs, 2x2 taps of that parity} small[...] W
//        four parity sub-GEMMs, M = B*Hs*Ws, N = Cb, K = 4*Cs, no zero insertion.  B = packed Wu[Cb][tap][Cs].
//   WGRAD: dW[cs][cb][tap] = sum_{pixels} small[p,cs] * big[p shifted by tap, cb]
//        GEMM M = Cs, N = Cb per tap, K = pixels; both operands MN-major straight from NHWC via TMA; 8 taps
//        accumulate side by side in the 512 TMEM columns; split-K partial tiles go to a workspace and a
//        second kernel reduces them into the fp32 gradient in PyTorch layout.
//
// Kernel shape: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+TMEM owner), warps 2-5 = epilogue.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tma_host.cuh"

namespace {

constexpr int kThreads = 192;
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;               // bf16 elements = 128 B = one swizzle row
constexpr int kATileBytes = kBlockM * 128;  // 16 KB
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;           // TMEM columns between the two accumulator stages
constexpr int kMaxStages = 8;

// folded affine + activation, branch-free (see act_fwd in common.cuh): one fma, one select
__device__ __forceinline__ float affine_nslope(int act, float slope) { return act_nslope(act, slope); }
__device__ __forceinline__ float affine_act(float v, float sc, float sh, float nslope) {
  const float t = fmaf(v, sc, sh);
  return t > 0.f ? t : t * nslope;
}

struct ConvGemmParams {
  int mode;  // 0 = DOWN, 1 = UP
  int B, Hm, Wm;  // M-side (small tensor) spatial dims
  int Wt, Ht, Bt;  // M tile = Wt*Ht*Bt = 128 pixels
  int tiles_w, tiles_h, tiles_b;
  int n_tiles, block_n;
  int Ck;       // channels of the K-side tensor
  int cpk;      // Ck / 64
  int k_iters;  // taps * cpk
  int N;        // output channels
  int Ho, Wo;   // output spatial dims
  int num_tiles;
  int num_stages;
  bf16* out;
  // optional epilogue extras
  const bf16* mask;   // same shape as out: out *= (mask > 0 ? 1 : mask_slope)  (LeakyReLU derivative of a BN-less layer)
  float mask_slope;
  float* img;         // epilogue "image": 3 real output channels written as fp32 NCHW planes instead of `out`
  int img_sigmoid, img_accumulate;
  int img4;           // image mode: one tile covers all four output parities (N = 4 x 3 columns, K = 9 pixel shifts)
  const float* aff_scale;   // eval-mode BatchNorm folded into the epilogue: out = act(acc * aff_scale[n] + aff_shift[n])
  const float* aff_shift;   // (fp32, then the single bf16 rounding); NULL = off
  int aff_act;
  float aff_slope;
  int stat_atomic;    // 1: stat_part is one zero-initialised accumulator row pair [2][N]; CTAs add their sums atomically
  float* stat_part;   // BatchNorm statistics fused in the epilogue: per-CTA partial sums [2][gridDim.x][N] of the fp32
                      // accumulators (sum, sum of squares) over the rows this CTA produced; NULL = off
  int splits, kps;    // split-K: work item = (tile, split); a split covers K-iterations [split*kps, (split+1)*kps)
  float* ws;          // split-K: fp32 [output pixels][N] accumulator (zeroed by the launcher), NULL when splits == 1
  int nacc;           // K-interleaved accumulators per tile (1, 2 or 4; nacc * block_n <= 256): back-to-back tcgen05.mma
                      // into ONE accumulator serialise on its read-modify-write, which a narrow (N <= 128) MMA is too
                      // short to hide; the k-substeps of a stage round-robin over nacc column blocks that the
                      // epilogue sums
  int debug;          // timing experiments (env DG_GEMM_DEBUG, results are garbage): 1 = no MMAs, 2 = no A loads,
                      // 3 = no B loads, 4 = no loads at all, 5 = no loads and A operand from TMEM,
                      // 6 = role-swapped kernel without the mask prefetch (results stay correct)
};

// Sum the 32 values each lane holds for 32 columns over the 32 lanes (rows) of the warp: afterwards v[0] of lane l is
// the column-(l) total.  Butterfly reduce-scatter, 31 shuffles.
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int d = 16, n = 16; d >= 1; d >>= 1, n >>= 1) {
    const bool upper = (lane & d) != 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < n) {
        const float send = upper ? v[j] : v[j + n];
        const float keep = upper ? v[j + n] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, d);
      }
    }
  }
  return v[0];
}

struct TileCoord {
  int par, nt, b0, h0, w0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvGemmParams& p, int tile) {
  TileCoord t;
  t.nt = tile % p.n_tiles;
  int r = tile / p.n_tiles;
  int tw = r % p.tiles_w;
  r /= p.tiles_w;
  int th = r % p.tiles_h;
  r /= p.tiles_h;
  int tb = r % p.tiles_b;
  t.par = r / p.tiles_b;
  t.w0 = tw * p.Wt;
  t.h0 = th * p.Ht;
  t.b0 = tb * p.Bt;
  return t;
}

// 32 accumulator columns of this thread's row, summed over the K-interleaved accumulators
__device__ __forceinline__ void load_acc32(uint32_t taddr, int c, int nacc, int block_n, uint32_t (&r)[32]) {
  tmem_ld_32x32(taddr + c, r);
  tmem_ld_wait();
  for (int j = 1; j < nacc; ++j) {
    uint32_t t[32];
    tmem_ld_32x32(taddr + j * block_n + c, t);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(t[e]));
  }
}

// tile index of a work unit: unit u of a CTA pair = M tiles (2*mp, 2*mp+1) of N tile nt, u = mp * n_tiles + nt
template <int kCta>
__device__ __forceinline__ int unit_tile(const ConvGemmParams& p, int u, uint32_t crank) {
  if (kCta == 1) return u;
  const int nt = u % p.n_tiles, mp = u / p.n_tiles;
  return (2 * mp + (int)crank) * p.n_tiles + nt;
}

// kCta = 2: launched as clusters of two CTAs that own two consecutive 128-row M tiles of the same N tile; one
// tcgen05.mma.cta_group::2 (M = 256) issued by the leader feeds both.  Each CTA loads its own A tile and HALF of the B
// rows, so the bytes an SM has to ingest per FLOP drop from (128+N) to (128+N/2) rows per k-step -- the L2->SM path is
// what bounds these GEMMs.
template <int kCta>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int stage_bytes = kATileBytes + (p.block_n / kCta) * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.num_stages * stage_bytes);
  uint64_t* full_bar = bars;                    // [stages]
  uint64_t* empty_bar = bars + kMaxStages;      // [stages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_stat = reinterpret_cast<float*>(bars + 32);   // [2][N] when p.stat_part

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = kCta == 2 ? cluster_ctarank() : 0u;
  // work units: (tile, split) for one CTA, (pair of M tiles, split) for a CTA pair
  const int unit0 = kCta == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_stride = kCta == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int unit_tiles = p.num_tiles / kCta;
  const int num_units = unit_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4 * kCta);   // the leader's MMA thread waits for the epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (kCta == 2)
      tmem_alloc_2cta(tmem_slot, kTmemCols);
    else
      tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (kCta == 2) cluster_sync_all();   // both CTAs' barriers and TMEM exist before any cross-CTA traffic
  // programmatic dependent launch: dependents may be scheduled once every CTA of this grid holds its TMEM columns;
  // nothing above touches global memory, everything below runs after the predecessor grid has completed
  griddep_launch_dependents();
  griddep_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int nrow0 = (int)crank * (p.block_n / kCta);   // this CTA's share of the B rows
      for (int w = unit0; w < num_units; w += unit_stride) {
        const TileCoord t = decode_tile(p, unit_tile<kCta>(p, w % unit_tiles, crank));
        const int k_begin = (w / unit_tiles) * p.kps, k_end = min(p.k_iters, k_begin + p.kps);
        const int py = t.par >> 1, px = t.par & 1;
        for (int it = k_begin; it < k_end; ++it) {
          const int tap_i = it / p.cpk;
          const int c0 = (it - tap_i * p.cpk) * kBlockK;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + kATileBytes;
          // pair: the leader's barrier counts the bytes of both CTAs' loads
          const bool la = p.debug != 2 && p.debug < 4, lb = p.debug != 3 && p.debug < 4;   // timing experiments
          if (crank == 0) {
            const uint32_t bytes = (la ? kATileBytes : 0) + (lb ? (uint32_t)(stage_bytes - kATileBytes) : 0);
            if (bytes)
              mbar_arrive_expect_tx(&full_bar[stage], bytes * kCta);
            else
              mbar_arrive(&full_bar[stage]);
          }
          if (p.mode == 0) {
            const int kh = tap_i >> 2, kw = tap_i & 3;
            const int dh = ((kh + 1) >> 1) - 1, ph = (kh + 1) & 1;
            const int dw = ((kw + 1) >> 1) - 1, pw = (kw + 1) & 1;
            if (kCta == 2) {
              if (la) tma_load_5d_2cta(sa, &tmA, &full_bar[stage], pw * p.Ck + c0, t.w0 + dw, ph, t.h0 + dh, t.b0);
              if (lb) tma_load_2d_2cta(sb, &tmB, &full_bar[stage], tap_i * p.Ck + c0, t.nt * p.block_n + nrow0);
            } else {
              if (la) tma_load_5d(sa, &tmA, &full_bar[stage], pw * p.Ck + c0, t.w0 + dw, ph, t.h0 + dh, t.b0);
              if (lb) tma_load_2d(sb, &tmB, &full_bar[stage], tap_i * p.Ck + c0, t.nt * p.block_n);
            }
          } else if (p.img4) {   // all parities at once: K runs over the 3 x 3 pixel shifts (weights packed to match)
            const int di = tap_i / 3 - 1, dj = tap_i % 3 - 1;
            if (la) tma_load_4d(sa, &tmA, &full_bar[stage], c0, t.w0 + dj, t.h0 + di, t.b0);
            if (lb) tma_load_2d(sb, &tmB, &full_bar[stage], tap_i * p.Ck + c0, 0);
          } else {
            const int th = tap_i >> 1, tw = tap_i & 1;
            const int kh = py == 0 ? (th ? 3 : 1) : (th ? 2 : 0);
            const int di = py == 0 ? (th ? -1 : 0) : (th ? 0 : 1);
            const int kw = px == 0 ? (tw ? 3 : 1) : (tw ? 2 : 0);
            const int dj = px == 0 ? (tw ? -1 : 0) : (tw ? 0 : 1);
            if (kCta == 2) {
              if (la) tma_load_4d_2cta(sa, &tmA, &full_bar[stage], c0, t.w0 + dj, t.h0 + di, t.b0);
              if (lb) tma_load_2d_2cta(sb, &tmB, &full_bar[stage], (kh * 4 + kw) * p.Ck + c0, t.nt * p.block_n + nrow0);
            } else {
              if (la) tma_load_4d(sa, &tmA, &full_bar[stage], c0, t.w0 + dj, t.h0 + di, t.b0);
              if (lb) tma_load_2d(sb, &tmB, &full_bar[stage], (kh * 4 + kw) * p.Ck + c0, t.nt * p.block_n);
            }
          }
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {
      const uint32_t idesc = make_idesc_bf16(kBlockM * kCta, p.block_n, 0, 0);
      // descriptors differ only in the 14-bit start-address field: build the constant part once and add the
      // (stage, k) offset per MMA so the single issuing thread spends a handful of instructions per tcgen05.mma
      const uint64_t desc_base = make_sdesc_sw128(0, 16, 1024);
      const uint32_t smem0 = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = unit0; w < num_units; w += unit_stride) {
        const int k_begin = (w / unit_tiles) * p.kps, k_end = min(p.k_iters, k_begin + p.kps);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccStride);
        for (int it = k_begin; it < k_end; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem0 + (uint32_t)(stage * stage_bytes);
          const uint64_t da0 = desc_base | (uint64_t)((sa & 0x3FFFFu) >> 4);
          const uint64_t db0 = desc_base | (uint64_t)(((sa + kATileBytes) & 0x3FFFFu) >> 4);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            if (p.debug == 1) break;
            const int j = k & (p.nacc - 1);   // accumulator of this k-substep
            const uint32_t accum = (it > k_begin || k >= p.nacc) ? 1u : 0u;
            const uint32_t d = d_tmem + (uint32_t)(j * p.block_n);
            if (kCta == 2)
              umma_bf16_2cta(d, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc, accum);
            else if (p.debug == 5)   // timing experiment: A operand from (uninitialised) TMEM columns
              umma_bf16_ts(d, tmem_base + 448u + (uint32_t)(8 * k), db0 + (uint64_t)(2 * k), idesc, accum);
            else
              umma_bf16(d, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc, accum);
          }
          if (kCta == 2)
            umma_commit_2cta(&empty_bar[stage], 3);   // frees the stage in both CTAs
          else
            umma_commit(&empty_bar[stage]);
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (kCta == 2)
          umma_commit_2cta(&tfull_bar[acc], 3);
        else
          umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // epilogue: TMEM lane quarter of this warp is (warp % 4)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const float aff_nslope = affine_nslope(p.aff_act, p.aff_slope);
    int acc = 0;
    uint32_t acc_phase = 0;
    if (p.stat_part) {
      for (int i = threadIdx.x - 64; i < 2 * p.N; i += 128) s_stat[i] = 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    for (int w = unit0; w < num_units; w += unit_stride) {
      const TileCoord t = decode_tile(p, unit_tile<kCta>(p, w % unit_tiles, crank));
      const int wl = row % p.Wt;
      const int hl = (row / p.Wt) % p.Ht;
      const int bl = row / (p.Wt * p.Ht);
      const int b = t.b0 + bl;
      int oy, ox;
      if (p.mode == 0) {
        oy = t.h0 + hl;
        ox = t.w0 + wl;
      } else {
        oy = 2 * (t.h0 + hl) + (t.par >> 1);
        ox = 2 * (t.w0 + wl) + (t.par & 1);
      }
      const size_t opix = ((size_t)b * p.Ho + oy) * p.Wo + ox;
      bf16* orow = p.out + opix * p.N + (size_t)t.nt * p.block_n;
      const bool valid = b < p.B;
      // masked epilogue (always N = 64): fetch the row's 128 mask bytes before waiting for the accumulator so
      // the DRAM latency overlaps the MMAs of this tile
      const bool premask = p.mask != nullptr && p.block_n == 64;
      uint4 mlo[4], mhi[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) mlo[j] = mhi[j] = make_uint4(0u, 0u, 0u, 0u);
      if (premask && valid) {
        const uint4* mp = reinterpret_cast<const uint4*>(p.mask + opix * p.N + (size_t)t.nt * p.block_n);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mlo[j] = mp[j];
          mhi[j] = mp[4 + j];
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccStride);
      if (p.ws != nullptr) {
        // split-K: accumulate this split's partial tile into the fp32 workspace with 16-byte vector reductions
        float* wrow = p.ws + opix * p.N + (size_t)t.nt * p.block_n;
        for (int c = 0; c < p.block_n; c += 32) {
          uint32_t r[32];
          load_acc32(taddr, c, p.nacc, p.block_n, r);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              red_add_f32x4(wrow + c + 4 * j, __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                            __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          }
        }
      } else if (p.img != nullptr) {
        // 3-channel image epilogue (N padded to 16): fp32 NCHW planes, optional sigmoid / accumulate
        uint32_t r[32];
        tmem_ld_32x32(taddr, r);   // columns >= 16 are never written by the MMA and are ignored
        tmem_ld_wait();
        if (valid) {
          const size_t plane = (size_t)p.Ho * p.Wo;
          if (p.img4) {   // columns parity * 3 + c of this row = the 2 x 2 output pixels of small pixel (h0 + hl, w0 + wl)
#pragma unroll
            for (int par = 0; par < 4; ++par) {
              float* o = p.img + (size_t)b * 3 * plane + (size_t)(2 * (t.h0 + hl) + (par >> 1)) * p.Wo + 2 * (t.w0 + wl) + (par & 1);
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                float v = __uint_as_float(r[par * 3 + c]);
                if (p.img_sigmoid) v = 1.f / (1.f + expf(-v));
                if (p.img_accumulate) v += o[c * plane];
                o[c * plane] = v;
              }
            }
          } else {
            float* o = p.img + (size_t)b * 3 * plane + (size_t)oy * p.Wo + ox;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float v = __uint_as_float(r[c]);
              if (p.img_sigmoid) v = 1.f / (1.f + expf(-v));
              if (p.img_accumulate) v += o[c * plane];
              o[c * plane] = v;
            }
          }
        }
      } else {
        const bf16* mrow = p.mask ? p.mask + opix * p.N + (size_t)t.nt * p.block_n : nullptr;
        for (int c = 0; c < p.block_n; c += 32) {
          uint32_t r[32];
          load_acc32(taddr, c, p.nacc, p.block_n, r);
          if (p.stat_part) {   // rows beyond the batch are exact zeros (TMA zero fill), so no masking is needed
            float v[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
            const float cs = warp_column_sums(v, lane);
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]) * __uint_as_float(r[e]);
            const float cq = warp_column_sums(v, lane);
            const int n = t.nt * p.block_n + c + lane;
            atomicAdd(&s_stat[n], cs);
            atomicAdd(&s_stat[p.N + n], cq);
          }
          if (valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[8 * j + e]);
              if (mrow) {
                float m[8];
                if (premask) {
                  const bool lo = c == 0;   // c is 0 or 32 here
                  const uint32_t w0 = lo ? mlo[j].x : mhi[j].x, w1 = lo ? mlo[j].y : mhi[j].y;
                  const uint32_t w2 = lo ? mlo[j].z : mhi[j].z, w3 = lo ? mlo[j].w : mhi[j].w;
                  m[0] = __uint_as_float(w0 << 16); m[1] = __uint_as_float(w0 & 0xffff0000u);
                  m[2] = __uint_as_float(w1 << 16); m[3] = __uint_as_float(w1 & 0xffff0000u);
                  m[4] = __uint_as_float(w2 << 16); m[5] = __uint_as_float(w2 & 0xffff0000u);
                  m[6] = __uint_as_float(w3 << 16); m[7] = __uint_as_float(w3 & 0xffff0000u);
                } else {
                  unpack8(*reinterpret_cast<const bf16x8*>(mrow + c + 8 * j), m);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] *= (m[e] > 0.f ? 1.f : p.mask_slope);
              }
              if (p.aff_scale) {   // same addresses for every thread of the warp: broadcast loads
                const int n0 = t.nt * p.block_n + c + 8 * j;
                const float4 s0 = *reinterpret_cast<const float4*>(p.aff_scale + n0);
                const float4 s1 = *reinterpret_cast<const float4*>(p.aff_scale + n0 + 4);
                const float4 h0 = *reinterpret_cast<const float4*>(p.aff_shift + n0);
                const float4 h1 = *reinterpret_cast<const float4*>(p.aff_shift + n0 + 4);
                const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = affine_act(f[e], sc[e], sh[e], aff_nslope);
              }
              *reinterpret_cast<bf16x8*>(orow + c + 8 * j) = pack8(f);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCta == 2)
          mbar_arrive_cluster(&tempty_bar[acc], 0);   // the accumulator stage is reused by the leader's MMA thread
        else
          mbar_arrive(&tempty_bar[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (p.stat_part) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (p.stat_atomic) {
        for (int i = threadIdx.x - 64; i < 2 * p.N; i += 128) atomicAdd(p.stat_part + i, s_stat[i]);
      } else {
        float* ps = p.stat_part + (size_t)blockIdx.x * p.N;
        float* pq = p.stat_part + ((size_t)gridDim.x + blockIdx.x) * p.N;
        for (int i = threadIdx.x - 64; i < p.N; i += 128) {
          ps[i] = s_stat[i];
          pq[i] = s_stat[p.N + i];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kCta == 2) cluster_sync_all();   // the pair's MMAs, commits and remote arrivals are all done
  if (warp == 1) {
    tc_fence_after();
    if (kCta == 2)
      tmem_dealloc_2cta(tmem_base, kTmemCols);
    else
      tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// Role-swapped variant for layers with <= 128 output channels.  A tcgen05.mma costs ~130 cycles + 0.3 * N regardless
// of N (measured: MMA-only rate 1450 / 830 / 510 TFLOP/s for N = 256 / 128 / 64), so a 128 x 64 tile can never run at
// more than a third of the tensor peak.  Here the WEIGHTS are the M operand (M = Cout = 64 or 128 rows) and 256 PIXELS
// (two consecutive 128-pixel tiles) are the N operand: every MMA is N = 256 wide.  The accumulator is D^T -- TMEM lane =
// output channel, column = pixel -- so a thread owns one channel: BatchNorm partial sums are plain register sums, and
// the 32 lanes of a warp store 32 consecutive channels (64 contiguous bytes) of one pixel per instruction.
// M = 64: the accumulator occupies lanes 0-15 of each 32-lane TMEM quarter (16 data paths per warp).
// ------------------------------------------------------------------------------------------------
constexpr int kSwapN = 256;
constexpr int kSwapTrStride = 36;   // floats per row of the epilogue transpose tile (32 + 4 padding)

// kAff: folded eval-mode BatchNorm epilogue; kStats: BatchNorm sums of the accumulators.  Compile-time switches: this
// kernel's epilogue (TMEM -> smem transpose -> 16-byte stores) is its bottleneck for N = 64, and every instruction added
// to the per-element loop shows (a run-time select between v and act(v*sc+sh) cost the training path 25-60 %).
template <bool kAff, bool kStats>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_swap_kernel(const __grid_constant__ CUtensorMap tmPix, const __grid_constant__ CUtensorMap tmW,
                      const ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int w_bytes = p.N * 128;                       // weight tile: Cout rows x 64 k
  const int stage_bytes = 2 * kATileBytes + w_bytes;   // 256 pixel rows + the weight rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.num_stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tfull_bar = bars + 2 * kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_units = p.num_tiles >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmPix);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_launch_dependents();
  griddep_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        const TileCoord t0 = decode_tile(p, 2 * u), t1 = decode_tile(p, 2 * u + 1);   // same parity plane
        const int py = t0.par >> 1, px = t0.par & 1;
        for (int it = 0; it < p.k_iters; ++it) {
          const int tap_i = it / p.cpk;
          const int c0 = (it - tap_i * p.cpk) * kBlockK;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sp = smem + stage * stage_bytes;
          uint8_t* sw = sp + 2 * kATileBytes;
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
          if (p.mode == 0) {
            const int kh = tap_i >> 2, kw = tap_i & 3;
            const int dh = ((kh + 1) >> 1) - 1, ph = (kh + 1) & 1;
            const int dw = ((kw + 1) >> 1) - 1, pw = (kw + 1) & 1;
            tma_load_5d(sp, &tmPix, &full_bar[stage], pw * p.Ck + c0, t0.w0 + dw, ph, t0.h0 + dh, t0.b0);
            tma_load_5d(sp + kATileBytes, &tmPix, &full_bar[stage], pw * p.Ck + c0, t1.w0 + dw, ph, t1.h0 + dh, t1.b0);
            tma_load_2d(sw, &tmW, &full_bar[stage], tap_i * p.Ck + c0, 0);
          } else {
            const int th = tap_i >> 1, tw = tap_i & 1;
            const int kh = py == 0 ? (th ? 3 : 1) : (th ? 2 : 0);
            const int di = py == 0 ? (th ? -1 : 0) : (th ? 0 : 1);
            const int kw = px == 0 ? (tw ? 3 : 1) : (tw ? 2 : 0);
            const int dj = px == 0 ? (tw ? -1 : 0) : (tw ? 0 : 1);
            tma_load_4d(sp, &tmPix, &full_bar[stage], c0, t0.w0 + dj, t0.h0 + di, t0.b0);
            tma_load_4d(sp + kATileBytes, &tmPix, &full_bar[stage], c0, t1.w0 + dj, t1.h0 + di, t1.b0);
            tma_load_2d(sw, &tmW, &full_bar[stage], (kh * 4 + kw) * p.Ck + c0, 0);
          }
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(p.N, kSwapN, 0, 0);
      const uint64_t desc_base = make_sdesc_sw128(0, 16, 1024);
      const uint32_t smem0 = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccStride);
        for (int it = 0; it < p.k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sp = smem0 + (uint32_t)(stage * stage_bytes);
          const uint64_t dpix = desc_base | (uint64_t)((sp & 0x3FFFFu) >> 4);
          const uint64_t dwt = desc_base | (uint64_t)(((sp + 2 * kATileBytes) & 0x3FFFFu) >> 4);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16(d_tmem, dwt + (uint64_t)(2 * k), dpix + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    const int q = warp & 3;
    const bool m64 = p.N == 64;
    const int ch = m64 ? q * 16 + lane : q * 32 + lane;   // this thread's output channel
    const bool ch_ok = !m64 || lane < 16;
    const int lw = 31 - __clz(p.Wt), lh = 31 - __clz(p.Ht);   // tile extents are powers of two
    // per-warp transpose tile [32 pixels][32 channels] fp32, rows padded to 144 bytes (conflict-free 16-byte reads)
    float* tr = reinterpret_cast<float*>(smem + p.num_stages * stage_bytes + 1024) + (warp - 2) * 32 * kSwapTrStride;
    float ssum = 0.f, ssq = 0.f;
    const float aff_nslope = affine_nslope(p.aff_act, p.aff_slope);
    const float asc = (kAff && ch < p.N) ? p.aff_scale[ch] : 1.f, ash = (kAff && ch < p.N) ? p.aff_shift[ch] : 0.f;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      const TileCoord tA = decode_tile(p, 2 * u), tB = decode_tile(p, 2 * u + 1);
      // element offset of (this lane's pixel of chunk c, this warp's first channel); -1 beyond the batch
      auto row_offset = [&](int c) -> long long {
        const bool first = c < 128;
        const int row = (c & 127) + lane;
        const int wl = row & (p.Wt - 1), hl = (row >> lw) & (p.Ht - 1), bl = row >> (lw + lh);
        const int b = (first ? tA.b0 : tB.b0) + bl;
        int oy = (first ? tA.h0 : tB.h0) + hl, ox = (first ? tA.w0 : tB.w0) + wl;
        if (p.mode == 1) {
          oy = 2 * oy + (tA.par >> 1);   // both tiles of a unit lie in the same parity plane
          ox = 2 * ox + (tA.par & 1);
        }
        if (b >= p.B) return -1;
        return (long long)((((size_t)b * p.Ho + oy) * p.Wo + ox) * p.N + (m64 ? q * 16 : q * 32));
      };
      const int nvec = m64 ? 2 : 4;
      // masked dgrad: the mask rows are fetched ahead of their use -- L2 prefetch of all eight chunk rows while the
      // unit's MMAs are still running, and the 16-byte vectors of chunk c+1 are loaded while chunk c is processed
      uint4 mcur[4], mnext[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) mcur[j] = mnext[j] = make_uint4(0u, 0u, 0u, 0u);
      const bool pre = p.mask != nullptr && p.debug != 6;
      if (pre) {
        for (int c = 0; c < kSwapN; c += 32) {
          const long long o = row_offset(c);
          if (o >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.mask + o));
        }
        const long long o0 = row_offset(0);
        if (o0 >= 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < nvec) mcur[j] = *reinterpret_cast<const uint4*>(p.mask + o0 + 8 * j);
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccStride);
      for (int c = 0; c < kSwapN; c += 32) {
        if (pre && c + 32 < kSwapN) {
          const long long on = row_offset(c + 32);
          if (on >= 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < nvec) mnext[j] = *reinterpret_cast<const uint4*>(p.mask + on + 8 * j);
          }
        }
        uint32_t r[32];
        tmem_ld_32x32(taddr + c, r);
        tmem_ld_wait();
        // (1) this thread's channel, 32 pixels -> statistics in registers, values into the warp's transpose tile
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float v = __uint_as_float(r[e]);
          if (kStats) {
            ssum += v;          // rows beyond the batch are exact zeros (TMA zero fill)
            ssq += v * v;
          }
          if (ch_ok) tr[e * kSwapTrStride + lane] = kAff ? affine_act(v, asc, ash, aff_nslope) : v;
        }
        __syncwarp();
        // (2) this thread's pixel, the warp's 32 (16) channels -> 16-byte vector stores
        const long long o = row_offset(c);
        if (o >= 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < nvec) {
              const float4 lo = *reinterpret_cast<const float4*>(tr + lane * kSwapTrStride + 8 * j);
              const float4 hi = *reinterpret_cast<const float4*>(tr + lane * kSwapTrStride + 8 * j + 4);
              float f[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
              if (p.mask) {
                float m[8];
                const bf16x8 mv = pre ? mcur[j] : *reinterpret_cast<const bf16x8*>(p.mask + o + 8 * j);
                unpack8(mv, m);
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] *= (m[k] > 0.f ? 1.f : p.mask_slope);
              }
              *reinterpret_cast<bf16x8*>(p.out + o + 8 * j) = pack8(f);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) mcur[j] = mnext[j];
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (kStats && p.stat_part && ch_ok) {
      if (p.stat_atomic) {
        atomicAdd(p.stat_part + ch, ssum);
        atomicAdd(p.stat_part + p.N + ch, ssq);
      } else {
        p.stat_part[(size_t)blockIdx.x * p.N + ch] = ssum;
        p.stat_part[((size_t)gridDim.x + blockIdx.x) * p.N + ch] = ssq;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// split-K finish: fp32 workspace -> bf16 output (8 elements per thread), optionally through the folded eval-mode
// BatchNorm affine + activation (channel = element index mod N)
__global__ void __launch_bounds__(256)
splitk_finish_kernel(const float* __restrict__ ws, bf16* __restrict__ out, long long n8, int N,
                     const float* __restrict__ aff_scale, const float* __restrict__ aff_shift, int aff_act,
                     float aff_slope) {
  griddep_launch_dependents();
  griddep_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(ws + i * 8);
    const float4 b = *reinterpret_cast<const float4*>(ws + i * 8 + 4);
    float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (aff_scale) {
      const int n0 = (int)((i * 8) % N);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        f[e] = affine_act(f[e], aff_scale[n0 + e], aff_shift[n0 + e], affine_nslope(aff_act, aff_slope));
    }
    *reinterpret_cast<bf16x8*>(out + i * 8) = pack8(f);
  }
}

// split-K finish that also accumulates the BatchNorm sums of the fp32 values it converts: acc[2][N] += {sum, sum of squares}
// over rows.  Block = (N/8 channel vectors) x (256 / (N/8) rows); a thread keeps its 8 channels, strides over rows.
__global__ void __launch_bounds__(256)
splitk_finish_stats_kernel(const float* __restrict__ ws, bf16* __restrict__ out, long long rows, int N,
                           float* __restrict__ acc) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ float sh1[256 * 8];
  __shared__ float sh2[256 * 8];
  const int cw = N >> 3, rows_iter = 256 / cw;
  const int tx = threadIdx.x % cw, ty = threadIdx.x / cw;
  const int c = tx * 8;
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  if (ty < rows_iter) {
    for (long long r = (long long)blockIdx.x * rows_iter + ty; r < rows; r += (long long)gridDim.x * rows_iter) {
      const float4 a = *reinterpret_cast<const float4*>(ws + r * N + c);
      const float4 b = *reinterpret_cast<const float4*>(ws + r * N + c + 4);
      const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      *reinterpret_cast<bf16x8*>(out + r * N + c) = pack8(f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s1[i] += f[i];
        s2[i] += f[i] * f[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sh1[threadIdx.x * 8 + i] = s1[i];
    sh2[threadIdx.x * 8 + i] = s2[i];
  }
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = 0.f, b = 0.f;
      for (int r = 0; r < rows_iter; ++r) {
        a += sh1[(r * cw + tx) * 8 + i];
        b += sh2[(r * cw + tx) * 8 + i];
      }
      atomicAdd(acc + c + i, a);
      atomicAdd(acc + N + c + i, b);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// WGRAD
// ------------------------------------------------------------------------------------------------
constexpr int kWgKC = 32;                       // pixels (GEMM-K) per pipeline stage
constexpr int kWgBoxBytes = kWgKC * 128;        // one 64-channel x 32-pixel box = 4 KB
constexpr int kWgStageBytes = 10 * kWgBoxBytes;  // 2 boxes of `small` (128 cs) + 8 taps of `big` (64 cb)
constexpr int kWgStages = 5;

struct WgradParams {
  int B, Hs, Ws;
  int Wt, Ht, Bt;  // K chunk = Wt*Ht*Bt = 32 pixels of the small tensor
  int chunks_w, chunks_h, chunks_b, total_chunks;
  int Cs, Cb, m_tiles, n_tiles;
  int splits, chunks_per_split;
  float* ws;  // [split][m_tile][n_tile][half][128][512]
  float* dw;  // direct mode (splits == 1): the epilogue writes dw = beta*dw + acc in PyTorch layout itself
  float beta;
  int direct;
  int cluster;  // 1: launched as clusters of 2 CTAs = two m-tiles of the same (n-tile, tap half, split); the 8 tap boxes
                // of `big` are identical for both, so each CTA fetches 4 of them and TMA-multicasts them to the pair
  int debug;  // timing experiments only: 1 = skip the MMAs, 2 = skip the TMA loads (results are garbage)
  int share;  // 0: one TMA box per tap.  1/2: taps that differ by a one-pixel shift inside the same parity plane share a
              // 33-pixel box and are addressed with a 128-byte row offset (1: descriptor base_offset = row phase, 2: 0)
};
constexpr int kWgSlot = 5 * 1024;  // 33 pixels x 128 B, padded to the 1024-byte swizzle repeat

__global__ void __launch_bounds__(kThreads, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmBig,
                  const __grid_constant__ CUtensorMap tmBig33, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kWgStages;
  uint64_t* tfull_bar = bars + 2 * kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int half, nt, mt;
  uint32_t crank = 0;
  if (p.cluster) {
    crank = cluster_ctarank();
    int v = blockIdx.x >> 1;
    const int mp = v % (p.m_tiles >> 1);
    v /= (p.m_tiles >> 1);
    half = v & 1;
    nt = v >> 1;
    mt = 2 * mp + (int)crank;
  } else {
    int w = blockIdx.x;
    half = w & 1;
    w >>= 1;
    nt = w % p.n_tiles;
    mt = w / p.n_tiles;
  }
  const uint16_t cmask = p.cluster ? (uint16_t)3 : (uint16_t)1;
  const int split = blockIdx.y;
  const int chunk_begin = split * p.chunks_per_split;
  const int chunk_end = min(p.total_chunks, chunk_begin + p.chunks_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmS);
    tma_prefetch_desc(&tmBig);
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], p.cluster ? 2 : 1);   // a stage is free once BOTH CTAs of the pair have consumed it
    }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (p.cluster) cluster_sync_all();   // the peer's barriers are initialised before any multicast traffic
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
        const int cw = r % p.chunks_w;
        r /= p.chunks_w;
        const int chh = r % p.chunks_h;
        const int cb_ = r / p.chunks_h;
        const int w0 = cw * p.Wt, h0 = chh * p.Ht, b0 = cb_ * p.Bt;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = smem + stage * kWgStageBytes;
        uint8_t* sb = sa + 2 * kWgBoxBytes;
        if (p.debug == 2) {
          mbar_arrive(&full_bar[stage]);
          if (++stage == kWgStages) {
            stage = 0;
            phase ^= 1u;
          }
          continue;
        }
        mbar_arrive_expect_tx(&full_bar[stage],
                              p.share ? (uint32_t)(2 * kWgBoxBytes + 4 * 33 * 128) : (uint32_t)kWgStageBytes);
        tma_load_4d(sa, &tmS, &full_bar[stage], mt * 128, w0, h0, b0);
        tma_load_4d(sa + kWgBoxBytes, &tmS, &full_bar[stage], mt * 128 + 64, w0, h0, b0);
        if (p.share) {
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            const int kh = half * 2 + (s4 >> 1), pw = s4 & 1;
            const int dh = ((kh + 1) >> 1) - 1, ph = (kh + 1) & 1;
            tma_load_5d(sb + s4 * kWgSlot, &tmBig33, &full_bar[stage], pw * p.Cb + nt * 64, w0 + (pw ? -1 : 0), ph,
                        h0 + dh, b0);
          }
        } else
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int tap = half * 8 + t;
          const int kh = tap >> 2, kw = tap & 3;
          const int dh = ((kh + 1) >> 1) - 1, ph = (kh + 1) & 1;
          const int dw = ((kw + 1) >> 1) - 1, pw = (kw + 1) & 1;
          if (!p.cluster)
            tma_load_5d(sb + t * kWgBoxBytes, &tmBig, &full_bar[stage], pw * p.Cb + nt * 64, w0 + dw, ph, h0 + dh, b0);
          else if ((uint32_t)(t & 1) == crank)
            tma_load_5d_mc(sb + t * kWgBoxBytes, &tmBig, &full_bar[stage], pw * p.Cb + nt * 64, w0 + dw, ph, h0 + dh,
                           b0, cmask);
        }
        if (++stage == kWgStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // One tcgen05.mma covers FOUR taps: their 64-channel boxes sit 4 KB apart in shared memory, which is exactly an
      // MN-major B operand with N = 256 and LBO = 4 KB, so a K-step needs 2 instructions instead of 8 (the single
      // issuing thread was the bottleneck at N = 64: 32 tensor-clocks of work per instruction).
      const uint32_t idesc = make_idesc_bf16(128, p.share ? 64 : 256, 1, 1);
      const uint64_t desc_base = make_sdesc_sw128(0, kWgBoxBytes, 1024);
      int stage = 0;
      uint32_t phase = 0;
      for (int ch = chunk_begin; ch < chunk_end; ++ch) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kWgStageBytes);
        const uint32_t sb = sa + 2 * kWgBoxBytes;
#pragma unroll
        for (int ks = 0; ks < kWgKC / 16; ++ks) {
          if (p.debug == 1 && ch > chunk_begin) break;
          const uint32_t accum = (ch > chunk_begin || ks > 0) ? 1u : 0u;
          // MN-major SW128: LBO = distance between 64-wide MN blocks, SBO = distance between 8-row K groups
          const uint64_t da = desc_base | (uint64_t)(((sa + ks * 2048) & 0x3FFFFu) >> 4);
          if (p.share) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const int kw = t & 3, pw = (kw + 1) & 1, dw = ((kw + 1) >> 1) - 1;
              const uint32_t rowoff = (uint32_t)(dw - (pw ? -1 : 0));
              const uint32_t addr = sb + (uint32_t)(((t >> 2) * 2 + pw) * kWgSlot) + rowoff * 128u + ks * 2048u;
              // a start address that is not 1024-byte aligned works with base_offset 0: the swizzle XOR is taken from
              // the absolute shared-memory address bits (verified on B200, DG_WGRAD_SHARE=2)
              umma_bf16(tmem_base + (uint32_t)(t * 64), da, desc_base | (uint64_t)((addr & 0x3FFFFu) >> 4), idesc, accum);
            }
          } else {
#pragma unroll
            for (int tq = 0; tq < 2; ++tq) {
              const uint32_t addr = sb + (uint32_t)(tq * 4 * kWgBoxBytes) + ks * 2048u;
              umma_bf16(tmem_base + (uint32_t)(tq * 256), da, desc_base | (uint64_t)((addr & 0x3FFFFu) >> 4), idesc, accum);
            }
          }
        }
        if (p.cluster)
          umma_commit_mc(&empty_bar[stage], cmask);
        else
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
    if (p.direct) {
      // dw[cs][cb][half*8 .. half*8+7]: the 8 taps of one (cs, cb) sit in 8 TMEM column blocks of this lane
      float* drow = p.dw + (((size_t)(mt * 128 + row)) * p.Cb + (size_t)nt * 64) * 16 + half * 8;
      for (int c8 = 0; c8 < 64; c8 += 8) {
        uint32_t r[8][8];
#pragma unroll
        for (int t = 0; t < 8; ++t) tmem_ld_32x8(taddr + (uint32_t)(t * 64 + c8), r[t]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float* d = drow + (size_t)(c8 + j) * 16;
          float4 o0 = make_float4(__uint_as_float(r[0][j]), __uint_as_float(r[1][j]), __uint_as_float(r[2][j]),
                                  __uint_as_float(r[3][j]));
          float4 o1 = make_float4(__uint_as_float(r[4][j]), __uint_as_float(r[5][j]), __uint_as_float(r[6][j]),
                                  __uint_as_float(r[7][j]));
          if (p.beta == 1.f) {   // accumulate: vector reductions in L2 keep the epilogue free of load round trips
            red_add_f32x4(d, o0.x, o0.y, o0.z, o0.w);
            red_add_f32x4(d + 4, o1.x, o1.y, o1.z, o1.w);
            continue;
          }
          if (p.beta != 0.f) {
            const float4 a0 = *reinterpret_cast<const float4*>(d);
            const float4 a1 = *reinterpret_cast<const float4*>(d + 4);
            o0.x += p.beta * a0.x; o0.y += p.beta * a0.y; o0.z += p.beta * a0.z; o0.w += p.beta * a0.w;
            o1.x += p.beta * a1.x; o1.y += p.beta * a1.y; o1.z += p.beta * a1.z; o1.w += p.beta * a1.w;
          }
          *reinterpret_cast<float4*>(d) = o0;
          *reinterpret_cast<float4*>(d + 4) = o1;
        }
      }
    } else {
      float* dst = p.ws + ((((size_t)split * p.m_tiles + mt) * p.n_tiles + nt) * 2 + half) * (size_t)(128 * 512) +
                   (size_t)row * 512;
      for (int c = 0; c < 512; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 v = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          *reinterpret_cast<uint4*>(dst + c + 4 * j) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.cluster) cluster_sync_all();   // neither CTA leaves while the peer can still arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad GEMM as CTA pairs (default for layers with Cs % 256 == 0; DG_WGRAD_PAIR=0 falls back to the multicast cluster
// kernel above).  Validated on B200 against tests/test_kernels_gpu.py -k wgrad; +35-45 % on the mid layers at 512x512.  Two CTAs (m-tiles 2mp, 2mp+1 of one n-tile / tap half / split) issue M = 256 tcgen05.mma.cta_group::2:
// each CTA loads its own 128 cs rows of `small` and only FOUR of the eight tap boxes of `big` -- the N = 256 operand of
// MMA group g (taps 4g..4g+3) is split across the pair, CTA r supplying taps 4g+2r, 4g+2r+1 -- so an SM ingests 24 KB
// per 32-pixel chunk instead of 40 KB (the existing cluster mode multicasts the boxes: same L2 reads, but every SM still
// receives all 40 KB).  Accumulator layout, epilogue, workspace and reduce kernel are those of wgrad_gemm_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kWpStageBytes = 6 * kWgBoxBytes;   // 2 boxes of `small` + 4 tap boxes of `big`
constexpr int kWpStages = 8;                     // 8 x 24 KB

__global__ void __launch_bounds__(kThreads, 1)
wgrad_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmBig,
                       const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWpStages * kWpStageBytes);
  uint64_t* full_bar = bars;                 // used in the leader CTA only
  uint64_t* empty_bar = bars + kWpStages;
  uint64_t* tfull_bar = bars + 2 * kWpStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  int v = blockIdx.x >> 1;
  const int mp = v % (p.m_tiles >> 1);
  v /= (p.m_tiles >> 1);
  const int half = v & 1;
  const int nt = v >> 1;
  const int mt = 2 * mp + (int)crank;
  const int split = blockIdx.y;
  const int chunk_begin = split * p.chunks_per_split;
  const int chunk_end = min(p.total_chunks, chunk_begin + p.chunks_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmS);
    tma_prefetch_desc(&tmBig);
    for (int s = 0; s < kWpStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();
  griddep_launch_dependents();
  griddep_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ch = chunk_begin; ch < chunk_end; ++ch) {
        int r = ch;
        const int cw = r % p.chunks_w;
        r /= p.chunks_w;
        const int chh = r % p.chunks_h;
        const int cb_ = r / p.chunks_h;
        const int w0 = cw * p.Wt, h0 = chh * p.Ht, b0 = cb_ * p.Bt;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = smem + stage * kWpStageBytes;
        uint8_t* sb = sa + 2 * kWgBoxBytes;
        if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(2 * kWpStageBytes));   // both CTAs' bytes
        tma_load_4d_2cta(sa, &tmS, &full_bar[stage], mt * 128, w0, h0, b0);
        tma_load_4d_2cta(sa + kWgBoxBytes, &tmS, &full_bar[stage], mt * 128 + 64, w0, h0, b0);
#pragma unroll
        for (int t4 = 0; t4 < 4; ++t4) {
          const int t = (t4 >> 1) * 4 + (int)crank * 2 + (t4 & 1);   // this CTA's share of MMA group t4 >> 1
          const int tap = half * 8 + t;
          const int kh = tap >> 2, kw = tap & 3;
          const int dh = ((kh + 1) >> 1) - 1, ph = (kh + 1) & 1;
          const int dw = ((kw + 1) >> 1) - 1, pw = (kw + 1) & 1;
          tma_load_5d_2cta(sb + t4 * kWgBoxBytes, &tmBig, &full_bar[stage], pw * p.Cb + nt * 64, w0 + dw, ph, h0 + dh, b0);
        }
        if (++stage == kWpStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {
      const uint32_t idesc = make_idesc_bf16(256, 256, 1, 1);
      const uint64_t desc_base = make_sdesc_sw128(0, kWgBoxBytes, 1024);
      int stage = 0;
      uint32_t phase = 0;
      for (int ch = chunk_begin; ch < chunk_end; ++ch) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kWpStageBytes);
        const uint32_t sb = sa + 2 * kWgBoxBytes;
#pragma unroll
        for (int ks = 0; ks < kWgKC / 16; ++ks) {
          const uint32_t accum = (ch > chunk_begin || ks > 0) ? 1u : 0u;
          const uint64_t da = desc_base | (uint64_t)(((sa + ks * 2048) & 0x3FFFFu) >> 4);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint32_t addr = sb + (uint32_t)(g * 2 * kWgBoxBytes) + ks * 2048u;
            umma_bf16_2cta(tmem_base + (uint32_t)(g * 256), da, desc_base | (uint64_t)((addr & 0x3FFFFu) >> 4), idesc, accum);
          }
        }
        umma_commit_2cta(&empty_bar[stage], 3);
        if (++stage == kWpStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit_2cta(tfull_bar, 3);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (p.direct) {
      float* drow = p.dw + (((size_t)(mt * 128 + row)) * p.Cb + (size_t)nt * 64) * 16 + half * 8;
      for (int c8 = 0; c8 < 64; c8 += 8) {
        uint32_t r[8][8];
#pragma unroll
        for (int t = 0; t < 8; ++t) tmem_ld_32x8(taddr + (uint32_t)(t * 64 + c8), r[t]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float* d = drow + (size_t)(c8 + j) * 16;
          float4 o0 = make_float4(__uint_as_float(r[0][j]), __uint_as_float(r[1][j]), __uint_as_float(r[2][j]),
                                  __uint_as_float(r[3][j]));
          float4 o1 = make_float4(__uint_as_float(r[4][j]), __uint_as_float(r[5][j]), __uint_as_float(r[6][j]),
                                  __uint_as_float(r[7][j]));
          if (p.beta == 1.f) {
            red_add_f32x4(d, o0.x, o0.y, o0.z, o0.w);
            red_add_f32x4(d + 4, o1.x, o1.y, o1.z, o1.w);
            continue;
          }
          if (p.beta != 0.f) {
            const float4 a0 = *reinterpret_cast<const float4*>(d);
            const float4 a1 = *reinterpret_cast<const float4*>(d + 4);
            o0.x += p.beta * a0.x; o0.y += p.beta * a0.y; o0.z += p.beta * a0.z; o0.w += p.beta * a0.w;
            o1.x += p.beta * a1.x; o1.y += p.beta * a1.y; o1.z += p.beta * a1.z; o1.w += p.beta * a1.w;
          }
          *reinterpret_cast<float4*>(d) = o0;
          *reinterpret_cast<float4*>(d + 4) = o1;
        }
      }
    } else {
      float* dst = p.ws + ((((size_t)split * p.m_tiles + mt) * p.n_tiles + nt) * 2 + half) * (size_t)(128 * 512) +
                   (size_t)row * 512;
      for (int c = 0; c < 512; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 vv = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          *reinterpret_cast<uint4*>(dst + c + 4 * j) = vv;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

// dw[cs][cb][tap] = beta*dw + sum_split ws[...]; one thread per (cs, cb, half) writes 8 consecutive taps.
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, float beta, int Cs, int Cb,
                                    int m_tiles, int n_tiles, int splits) {
  griddep_launch_dependents();
  griddep_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)Cs * Cb * 2;
  if (idx >= total) return;
  const int cb = (int)(idx % Cb);
  long long r = idx / Cb;
  const int half = (int)(r & 1);
  const int cs = (int)(r >> 1);
  const int mt = cs >> 7, row = cs & 127, nt = cb >> 6, cbl = cb & 63;
  float acc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[t] = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float* src = ws + ((((size_t)s * m_tiles + mt) * n_tiles + nt) * 2 + half) * (size_t)(128 * 512) +
                       (size_t)row * 512 + cbl;
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] += src[t * 64];
  }
  float* d = dw + ((size_t)cs * Cb + cb) * 16 + half * 8;
  float4 o0, o1;
  if (beta != 0.f) {
    o0 = *reinterpret_cast<float4*>(d);
    o1 = *reinterpret_cast<float4*>(d + 4);
    o0.x = beta * o0.x + acc[0]; o0.y = beta * o0.y + acc[1]; o0.z = beta * o0.z + acc[2]; o0.w = beta * o0.w + acc[3];
    o1.x = beta * o1.x + acc[4]; o1.y = beta * o1.y + acc[5]; o1.z = beta * o1.z + acc[6]; o1.w = beta * o1.w + acc[7];
  } else {
    o0 = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o1 = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  *reinterpret_cast<float4*>(d) = o0;
  *reinterpret_cast<float4*>(d + 4) = o1;
}

// ------------------------------------------------------------------------------------------------
// host side (tensor-map helpers live in tma_host.cuh)
// ------------------------------------------------------------------------------------------------
// Per-call options (dg_conv_opts in the header; NULL = defaults).  The library keeps no mutable launch state: the split-K
// workspace and the tiling overrides travel with the call, kernel attributes and the SM count are cached per device.
struct ConvGemmExtras {
  float* splitk_ws = nullptr;  // fp32 workspace for split-K; split-K is off without it
  size_t splitk_ws_bytes = 0;
  int force_bn = 0;            // test hook: N tile (0 = heuristic; 1 = the role-swapped kernel wherever it is eligible)
  int force_pair = -1;         // test hook: CTA pairs (-1 = default / DG_GEMM_PAIR)
  const void* mask = nullptr;
  float mask_slope = 0.f;
  float* img = nullptr;  // when set: Cb (mode 1 output channels) is the padded 16 and only 3 planes are written
  int img_sigmoid = 0, img_accumulate = 0;
  const float* aff_scale = nullptr;   // folded eval-mode BatchNorm: out = act(acc * scale[n] + shift[n])
  const float* aff_shift = nullptr;
  int aff_act = 0;
  float aff_slope = 0.f;
  float* stat_part = nullptr;  // [2][grid][N] partial BatchNorm sums, or (stat_atomic) [2][N] zero-initialised accumulators
  int stat_atomic = 0;
  int* grid_out = nullptr;     // plan query: receives the grid size, nothing is launched
};

int launch_conv_gemm(int mode, const void* a, const void* wpacked, void* out, int B, int Hs, int Ws, int Cs, int Cb,
                     cudaStream_t stream, const ConvGemmExtras& ex = ConvGemmExtras()) {
  DG_CHECK_ARG(B > 0 && is_pow2(Hs) && is_pow2(Ws), "conv gemm: B=%d Hs=%d Ws=%d must be positive / powers of two", B,
               Hs, Ws);
  const bool img_mode = ex.img != nullptr;
  DG_CHECK_ARG(Cs % 64 == 0 && Cs >= 64 && ((Cb % 64 == 0 && Cb >= 64) || (img_mode && mode == 1 && Cb == 16)),
               "conv gemm: Cs=%d Cb=%d must be multiples of 64", Cs, Cb);
  DG_CHECK_ARG(((uintptr_t)a & 15) == 0 && ((uintptr_t)wpacked & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
                   ((uintptr_t)ex.mask & 15) == 0,
               "conv gemm: pointers must be 16-byte aligned");
  ConvGemmParams p;
  p.mode = mode;
  p.B = B;
  p.Hm = Hs;
  p.Wm = Ws;
  tile_shape(kBlockM, Hs, Ws, &p.Wt, &p.Ht, &p.Bt);
  p.tiles_w = Ws / p.Wt;
  p.tiles_h = Hs / p.Ht;
  p.tiles_b = dg_ceil_div(B, p.Bt);
  const int N = mode == 0 ? Cs : Cb;
  p.N = N;
  p.Ck = mode == 0 ? Cb : Cs;
  p.cpk = p.Ck / 64;
  p.img4 = img_mode ? 1 : 0;   // image mode: all four parities per tile, K over the 3 x 3 pixel shifts (dg_c3_pack_weights)
  p.k_iters = (mode == 0 ? 16 : (p.img4 ? 9 : 4)) * p.cpk;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b * ((mode == 0 || p.img4) ? 1 : 4);
  // N tile: measured cost of one k-iteration of a 128 x bn x 64 tile is ~{904, 782, 678} cycles for bn = {256, 128, 64}
  // on every layer shape (profiles/README.md), i.e. wide tiles are worth more than filling the last SMs: pick the bn
  // that minimises waves(bn) * cost(bn) on the persistent grid
  int bn = img_mode ? 16 : 64;
  static int bn_policy = -1;   // DG_GEMM_BNPOLICY=0: round-1 rule (smallest tile count that still fills the SMs), A/B only
  if (bn_policy < 0) {
    const char* e = getenv("DG_GEMM_BNPOLICY");
    bn_policy = e ? atoi(e) : 1;
  }
  if (!img_mode && bn_policy == 0) {
    const int cands[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i) {
      if (N % cands[i]) continue;
      bn = cands[i];
      if ((long long)m_tiles * (N / bn) >= num_sms()) break;
    }
  } else if (!img_mode) {
    const int cands[3] = {256, 128, 64};
    const int cost[3] = {904, 782, 678};
    long long best = -1;
    for (int i = 0; i < 3; ++i) {
      if (N % cands[i]) continue;
      const long long tiles = (long long)m_tiles * (N / cands[i]);
      const long long t = ((tiles + num_sms() - 1) / num_sms()) * cost[i];
      if (best < 0 || t < best) {
        best = t;
        bn = cands[i];
      }
    }
  }
  // SM-starved big GEMMs (few M tiles, long K): prefer the widest N tile (best bytes/FLOP) and fill the machine with
  // split-K instead of shrinking the tile
  // split-K thresholds: minimum problem size and minimum k-iterations per split.  Measured on the 64^2 step
  // (B = 64): 1 GFLOP / 16 k-iterations gives 2.04 ms/step against 2.16 without split-K on the 4-GFLOP deep layers;
  // 8 k-iterations per split loses it again to the reduction traffic
  static double sk_gflop = -1.0;
  static int sk_mink = 16;
  if (sk_gflop < 0.0) {
    const char* e = getenv("DG_SPLITK_GFLOP");
    sk_gflop = e ? atof(e) : 1.0;
    const char* e2 = getenv("DG_SPLITK_MINK");
    sk_mink = e2 ? atoi(e2) : 16;
    if (sk_mink < 1) sk_mink = 1;
  }
  const double gflop_all = 2.0 * B * Hs * Ws * (double)Cs * Cb * 16 * 1e-9;
  const int k_iters_all = (mode == 0 ? 16 : 4) * ((mode == 0 ? Cb : Cs) / 64);
  if (!img_mode && !ex.mask && ex.splitk_ws && gflop_all > sk_gflop && k_iters_all >= 2 * sk_mink && N % 256 == 0 &&
      (long long)m_tiles * (N / 256) * 2 <= num_sms())
    bn = 256;
  if (ex.force_bn > 1 && !img_mode && N % ex.force_bn == 0) bn = ex.force_bn;
  // narrow layers (<= 128 output channels): role-swapped kernel, weights as M, 256 pixels as N
  static int swap_mode = -1;
  if (swap_mode < 0) {
    const char* e = getenv("DG_GEMM_SWAP");
    swap_mode = e ? atoi(e) : 1;
  }
  const int tiles_plane = p.tiles_w * p.tiles_h * p.tiles_b;
  const bool swap_ok = !img_mode && (N == 64 || N == 128) && tiles_plane % 2 == 0;
  const bool use_swap = swap_ok && ex.force_bn != 64 && ex.force_bn != 128 &&
                        (ex.force_bn == 1 || (swap_mode && m_tiles / 2 >= (num_sms() * 7) / 8));
  if (use_swap) bn = N;
  p.block_n = bn;
  p.n_tiles = N / bn;
  p.num_tiles = m_tiles * p.n_tiles;
  p.Ho = mode == 0 ? Hs : 2 * Hs;
  p.Wo = mode == 0 ? Ws : 2 * Ws;
  p.out = reinterpret_cast<bf16*>(out);
  p.mask = reinterpret_cast<const bf16*>(ex.mask);
  p.mask_slope = ex.mask_slope;
  p.img = ex.img;
  p.img_sigmoid = ex.img_sigmoid;
  p.img_accumulate = ex.img_accumulate;
  p.stat_part = ex.stat_part;
  p.stat_atomic = ex.stat_atomic;
  DG_CHECK_ARG(!ex.aff_scale || (ex.aff_shift && !ex.mask && !img_mode && !ex.stat_part),
               "conv gemm: the folded affine epilogue excludes mask / image / statistics epilogues");
  p.aff_scale = ex.aff_scale;     // split-K tiles go to the workspace raw: the finish kernel applies the affine
  p.aff_shift = ex.aff_shift;
  p.aff_act = ex.aff_act;
  p.aff_slope = ex.aff_slope;
  // split-K for big GEMMs that would leave most SMs idle (M = B*Hs*Ws small, K = taps*Ck large): partial tiles are
  // accumulated in an fp32 workspace and converted afterwards; such launches cannot fuse the BatchNorm statistics
  p.splits = 1;
  p.kps = p.k_iters;
  p.ws = nullptr;
  static int debug_mode = -1;
  if (debug_mode < 0) {
    const char* e = getenv("DG_GEMM_DEBUG");
    debug_mode = e ? atoi(e) : 0;
  }
  p.debug = debug_mode;
  static int nacc_mode = -1;
  if (nacc_mode < 0) {
    const char* e = getenv("DG_GEMM_NACC");
    nacc_mode = e ? atoi(e) : 1;
  }
  p.nacc = 1;
  while (!img_mode && p.nacc * 2 <= nacc_mode && p.nacc * 2 * bn <= kAccStride) p.nacc *= 2;
  const double gflop = 2.0 * B * Hs * Ws * (double)Cs * Cb * 16 * 1e-9;
  const size_t ws_need = (size_t)B * p.Ho * p.Wo * N * sizeof(float);
  if (!use_swap && !img_mode && !ex.mask && ex.splitk_ws && ws_need <= ex.splitk_ws_bytes && gflop > sk_gflop &&
      p.num_tiles * 2 <= num_sms() && p.k_iters >= 2 * sk_mink) {
    int splits = num_sms() / p.num_tiles;
    if (splits > p.k_iters / sk_mink) splits = p.k_iters / sk_mink;
    if (splits > 1) {
      p.kps = dg_ceil_div(p.k_iters, splits);
      p.splits = dg_ceil_div(p.k_iters, p.kps);
      p.ws = ex.splitk_ws;
    }
  }
  // CTA pairs (cta_group::2) whenever two consecutive M tiles of one parity plane exist and half an N tile is still
  // a whole number of 64-row TMA boxes
  static int pair_mode = -1;
  if (pair_mode < 0) {
    const char* e = getenv("DG_GEMM_PAIR");
    pair_mode = e ? atoi(e) : 1;
  }
  const int tiles_per_plane = p.tiles_w * p.tiles_h * p.tiles_b;
  const int want_pair = ex.force_pair >= 0 ? ex.force_pair : pair_mode;
  const int ncta = (!use_swap && want_pair && !img_mode && bn >= 64 && tiles_per_plane % 2 == 0) ? 2 : 1;
  int work = (p.num_tiles / ncta) * p.splits;
  if (use_swap) work = p.num_tiles / 2;
  const int max_units = num_sms() / ncta;
  const int grid = (work < max_units ? work : max_units) * ncta;
  if (ex.grid_out) {
    *ex.grid_out = p.ws ? 0 : grid;   // 0: statistics cannot be fused for this shape (split-K)
    return DG_OK;
  }
  float* splitk_acc = nullptr;
  if (p.ws) {
    DG_CHECK_ARG(ex.stat_part == nullptr || ex.stat_atomic, "conv gemm: fused statistics requested for a split-K shape");
    splitk_acc = ex.stat_part;      // accumulator mode: the finish kernel sums the fp32 values it converts
    p.stat_part = nullptr;
    p.aff_scale = p.aff_shift = nullptr;
    cudaMemsetAsync(p.ws, 0, ws_need, stream);
  }
  const int stage_bytes = use_swap ? 2 * kATileBytes + N * 128 : kATileBytes + (bn / ncta) * 128;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  p.num_stages = stages;
  const int smem_bytes = stages * stage_bytes + 1024 + 256 + (p.stat_part && !use_swap ? 2 * N * (int)sizeof(float) : 0) +
                         (use_swap ? 4 * 32 * kSwapTrStride * 4 + 768 : 0);
  DG_CHECK_ARG(smem_bytes <= 227 * 1024, "conv gemm: N=%d too wide for fused statistics", N);

  CUtensorMap tmA, tmB;
  int rc;
  if (mode == 0) {
    rc = make_parity_map(&tmA, a, B, 2 * Hs, 2 * Ws, Cb, p.Wt, p.Ht, p.Bt);
    if (rc) return rc;
    rc = make_weight_map(&tmB, wpacked, Cs, 16 * Cb, use_swap ? N : bn / ncta);
  } else {
    rc = make_nhwc_map(&tmA, a, B, Hs, Ws, Cs, p.Wt, p.Ht, p.Bt);
    if (rc) return rc;
    rc = make_weight_map(&tmB, wpacked, Cb, (p.img4 ? 9 : 16) * Cs, use_swap ? N : bn / ncta);
  }
  if (rc) return rc;
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_gemm_swap_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_gemm_swap_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_gemm_swap_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      dg_set_error("conv gemm: cannot raise dynamic smem: %s", cudaGetErrorString(e));
      return DG_ERR_CUDA;
    }
    attr_set = true;
  }
  if (use_swap)
  {
    if (p.aff_scale)
      dg_launch(conv_gemm_swap_kernel<true, false>, dg_cfg(grid, kThreads, smem_bytes, stream), tmA, tmB, p);
    else if (p.stat_part)
      dg_launch(conv_gemm_swap_kernel<false, true>, dg_cfg(grid, kThreads, smem_bytes, stream), tmA, tmB, p);
    else
      dg_launch(conv_gemm_swap_kernel<false, false>, dg_cfg(grid, kThreads, smem_bytes, stream), tmA, tmB, p);
  }
  else if (ncta == 2)
    dg_launch(conv_gemm_kernel<2>, dg_cfg(grid, kThreads, smem_bytes, stream, 2), tmA, tmB, p);
  else
    dg_launch(conv_gemm_kernel<1>, dg_cfg(grid, kThreads, smem_bytes, stream), tmA, tmB, p);
  DG_CHECK_LAUNCH("conv_gemm_kernel");
  if (p.ws) {
    const long long n8 = (long long)(ws_need / sizeof(float)) / 8;
    long long blocks = (n8 + 255) / 256;
    if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
    if (splitk_acc) {
      const long long rows = (long long)B * p.Ho * p.Wo;
      const int rows_iter = 256 / (N / 8);
      long long sb = (rows + rows_iter * 4 - 1) / (rows_iter * 4);     // ~4 rows per thread
      if (sb > 2LL * num_sms()) sb = 2LL * num_sms();
      if (sb < 1) sb = 1;
      dg_launch(splitk_finish_stats_kernel, dg_cfg((int)sb, 256, 0, stream), (const float*)p.ws, p.out, rows, N, splitk_acc);
    } else {
      dg_launch(splitk_finish_kernel, dg_cfg((int)blocks, 256, 0, stream), (const float*)p.ws, p.out, n8, N, ex.aff_scale,
                ex.aff_shift, ex.aff_act, ex.aff_slope);
    }
    DG_CHECK_LAUNCH("splitk_finish_kernel");
  }
  return DG_OK;
}

void wgrad_plan(int B, int Hs, int Ws, int Cs, int Cb, WgradParams* p) {
  p->B = B;
  p->Hs = Hs;
  p->Ws = Ws;
  tile_shape(kWgKC, Hs, Ws, &p->Wt, &p->Ht, &p->Bt);
  p->chunks_w = Ws / p->Wt;
  p->chunks_h = Hs / p->Ht;
  p->chunks_b = dg_ceil_div(B, p->Bt);
  p->total_chunks = p->chunks_w * p->chunks_h * p->chunks_b;
  p->Cs = Cs;
  p->Cb = Cb;
  p->m_tiles = Cs / 128;
  p->n_tiles = Cb / 64;
  const int work = p->m_tiles * p->n_tiles * 2;
  // one CTA per SM (200 KB of smem): aim for a single full wave -- work*splits <= #SMs -- instead of 2+ partial waves,
  // which also halves the split-K partial traffic
  int splits = work >= num_sms() ? 1 : num_sms() / work;
  const int max_splits = p->total_chunks / 16;  // at least 16 K-chunks (512 pixels) per split: keeps the
  if (splits > max_splits) splits = max_splits;  // workspace traffic below the operand traffic on small layers
  if (splits < 1) splits = 1;
  p->chunks_per_split = dg_ceil_div(p->total_chunks, splits);
  p->splits = dg_ceil_div(p->total_chunks, p->chunks_per_split);
}

}  // namespace

static ConvGemmExtras extras_from(const dg_conv_opts* o) {
  ConvGemmExtras ex;
  if (o) {
    ex.splitk_ws = reinterpret_cast<float*>(o->splitk_ws);
    ex.splitk_ws_bytes = o->splitk_ws_bytes;
    ex.force_bn = o->block_n;
    ex.force_pair = o->pair;
    ex.stat_atomic = o->stat_accumulate;
    ex.aff_scale = o->affine_scale;
    ex.aff_shift = o->affine_scale ? o->affine_shift : nullptr;
    ex.aff_act = o->affine_act;
    ex.aff_slope = o->affine_slope;
  }
  return ex;
}

extern "C" {

int dg_conv_opts_check(const dg_conv_opts* o) {
  if (!o) return DG_OK;
  DG_CHECK_ARG(o->block_n == 0 || o->block_n == 1 || o->block_n == 64 || o->block_n == 128 || o->block_n == 256,
               "conv opts: block_n=%d", o->block_n);
  DG_CHECK_ARG(o->pair >= -1 && o->pair <= 1 && o->wgrad_pair >= -1 && o->wgrad_pair <= 1, "conv opts: pair=%d wgrad_pair=%d",
               o->pair, o->wgrad_pair);
  DG_CHECK_ARG(((uintptr_t)o->splitk_ws & 15) == 0, "conv opts: split-K workspace must be 16-byte aligned");
  DG_CHECK_ARG(o->stat_accumulate == 0 || o->stat_accumulate == 1, "conv opts: stat_accumulate=%d", o->stat_accumulate);
  DG_CHECK_ARG(!o->affine_scale || (o->affine_shift && o->affine_act >= 0 && o->affine_act <= 2 &&
                                    (((uintptr_t)o->affine_scale | (uintptr_t)o->affine_shift) & 15) == 0),
               "conv opts: affine epilogue needs 16-byte aligned scale and shift vectors and an activation code 0..2");
  return DG_OK;
}

int dg_conv4x4s2_fprop(const void* x, const void* wd, void* z, int B, int H, int W, int Cb, int Cs,
                       const dg_conv_opts* opts, cudaStream_t stream) {
  DG_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "fprop: H=%d W=%d must be even", H, W);
  if (int rc = dg_conv_opts_check(opts)) return rc;
  return launch_conv_gemm(0, x, wd, z, B, H / 2, W / 2, Cs, Cb, stream, extras_from(opts));
}

int dg_conv4x4s2_dgrad(const void* dz, const void* wu, void* dx, int B, int Hs, int Ws, int Cs, int Cb,
                       const dg_conv_opts* opts, cudaStream_t stream) {
  if (int rc = dg_conv_opts_check(opts)) return rc;
  return launch_conv_gemm(1, dz, wu, dx, B, Hs, Ws, Cs, Cb, stream, extras_from(opts));
}

// Forward convolutions with the BatchNorm statistics of their output fused in the epilogue.
// mode 0 = Conv2d fprop (x big -> z small), mode 1 = ConvTranspose2d fprop (x small -> z big).
// stat_part: float[2 * rows * N] with rows = dg_conv_stats_rows(...), N = output channels; feed dg_bn_stats_finalize.
int dg_conv_stats_rows(int mode, int B, int Hs, int Ws, int Cs, int Cb, const dg_conv_opts* opts) {
  int grid = 0;
  if (dg_conv_opts_check(opts)) return 0;
  ConvGemmExtras ex = extras_from(opts);
  if (ex.splitk_ws_bytes && !ex.splitk_ws) ex.splitk_ws = reinterpret_cast<float*>(16);   // plan query without a buffer
  ex.grid_out = &grid;
  if (launch_conv_gemm(mode, (const void*)16, (const void*)16, (void*)16, B, Hs, Ws, Cs, Cb, 0, ex) != DG_OK) return 0;
  if (grid == 0 && ex.stat_atomic) return 1;   // split-K shape: in accumulator mode the finish kernel produces the sums
  return grid;
}
int dg_conv4x4s2_fprop_stats(const void* x, const void* wd, void* z, float* stat_part, int B, int H, int W, int Cb,
                             int Cs, const dg_conv_opts* opts, cudaStream_t stream) {
  DG_CHECK_ARG(H % 2 == 0 && W % 2 == 0 && stat_part, "fprop_stats: bad args");
  if (int rc = dg_conv_opts_check(opts)) return rc;
  ConvGemmExtras ex = extras_from(opts);
  ex.stat_part = stat_part;
  return launch_conv_gemm(0, x, wd, z, B, H / 2, W / 2, Cs, Cb, stream, ex);
}
int dg_convT4x4s2_fprop_stats(const void* x_small, const void* wu, void* y_big, float* stat_part, int B, int Hs, int Ws,
                              int Cs, int Cb, const dg_conv_opts* opts, cudaStream_t stream) {
  DG_CHECK_ARG(stat_part != nullptr, "convT_fprop_stats: bad args");
  if (int rc = dg_conv_opts_check(opts)) return rc;
  ConvGemmExtras ex = extras_from(opts);
  ex.stat_part = stat_part;
  return launch_conv_gemm(1, x_small, wu, y_big, B, Hs, Ws, Cs, Cb, stream, ex);
}

// dgrad whose consumer is a BN-less LeakyReLU layer: dx = dgrad * (mask > 0 ? 1 : slope), mask = that layer's output
int dg_conv4x4s2_dgrad_masked(const void* dz, const void* wu, void* dx, const void* mask, float slope, int B, int Hs,
                              int Ws, int Cs, int Cb, const dg_conv_opts* opts, cudaStream_t stream) {
  if (int rc = dg_conv_opts_check(opts)) return rc;
  ConvGemmExtras ex = extras_from(opts);
  ex.mask = mask;
  ex.mask_slope = slope;
  return launch_conv_gemm(1, dz, wu, dx, B, Hs, Ws, Cs, Cb, stream, ex);
}

// 64 -> 3 channel "up" layer on the tensor cores (N padded to 16): img[B,3,S,S] fp32 NCHW (+)= [sigmoid](convT(x64, wu3))
// x64 = bf16 [B,S/2,S/2,64], wu3 = bf16 [16][16 taps][64] from dg_c3_pack_weights.
int dg_c3_up_tc(const void* x64, const void* wu3, float* img, int B, int S, int sigmoid, int accumulate,
                cudaStream_t stream) {
  DG_CHECK_ARG(img != nullptr && S % 2 == 0, "c3_up_tc: bad args");
  ConvGemmExtras ex;
  ex.img = img;
  ex.img_sigmoid = sigmoid;
  ex.img_accumulate = accumulate;
  return launch_conv_gemm(1, x64, wu3, nullptr, B, S / 2, S / 2, 64, 16, stream, ex);
}

size_t dg_conv4x4s2_wgrad_workspace(int B, int Hs, int Ws, int Cs, int Cb) {
  if (B <= 0 || !is_pow2(Hs) || !is_pow2(Ws) || Cs % 128 || Cb % 64) return 0;
  WgradParams p;
  wgrad_plan(B, Hs, Ws, Cs, Cb, &p);
  return (size_t)p.splits * p.m_tiles * p.n_tiles * 2 * 128 * 512 * sizeof(float);
}

int dg_conv4x4s2_wgrad(const void* small, const void* big, float* dw, float beta, int B, int Hs, int Ws, int Cs,
                       int Cb, void* ws, size_t ws_bytes, const dg_conv_opts* opts, cudaStream_t stream) {
  if (int rc = dg_conv_opts_check(opts)) return rc;
  DG_CHECK_ARG(B > 0 && is_pow2(Hs) && is_pow2(Ws), "wgrad: B=%d Hs=%d Ws=%d must be positive / powers of two", B, Hs,
               Ws);
  DG_CHECK_ARG(Cs % 128 == 0 && Cb % 64 == 0, "wgrad: Cs=%d must be a multiple of 128 and Cb=%d of 64", Cs, Cb);
  DG_CHECK_ARG(Hs * Ws * 1LL >= 1 && (Hs * Ws >= kWgKC || (kWgKC % (Hs * Ws)) == 0), "wgrad: bad spatial size");
  WgradParams p;
  wgrad_plan(B, Hs, Ws, Cs, Cb, &p);
  p.direct = p.splits == 1;
  p.dw = dw;
  p.beta = beta;
  const size_t need = p.direct ? 0 : (size_t)p.splits * p.m_tiles * p.n_tiles * 2 * 128 * 512 * sizeof(float);
  DG_CHECK_ARG(p.direct || (ws != nullptr && ws_bytes >= need), "wgrad: workspace too small (%zu < %zu)", ws_bytes, need);
  DG_CHECK_ARG(((uintptr_t)small & 15) == 0 && ((uintptr_t)big & 15) == 0 && ((uintptr_t)dw & 15) == 0 &&
                   ((uintptr_t)ws & 15) == 0,
               "wgrad: pointers must be 16-byte aligned");
  p.ws = reinterpret_cast<float*>(ws);
  static int share_mode = -1;
  if (share_mode < 0) {
    const char* e = getenv("DG_WGRAD_SHARE");
    share_mode = e ? atoi(e) : 0;
  }
  p.share = (share_mode > 0 && p.Wt == 32 && p.Ht == 1 && p.Bt == 1) ? 2 : 0;
  static int debug_mode = -1;
  if (debug_mode < 0) {
    const char* e = getenv("DG_WGRAD_DEBUG");
    debug_mode = e ? atoi(e) : 0;
  }
  p.debug = debug_mode;
  CUtensorMap tmS, tmBig, tmBig33;
  int rc = make_nhwc_map(&tmS, small, B, Hs, Ws, Cs, p.Wt, p.Ht, p.Bt);
  if (rc) return rc;
  rc = make_parity_map(&tmBig, big, B, 2 * Hs, 2 * Ws, Cb, p.Wt, p.Ht, p.Bt);
  if (rc) return rc;
  rc = make_parity_map(&tmBig33, big, B, 2 * Hs, 2 * Ws, Cb, p.share ? 33 : p.Wt, p.Ht, p.Bt);
  if (rc) return rc;
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      dg_set_error("wgrad: cannot raise dynamic smem: %s", cudaGetErrorString(e));
      return DG_ERR_CUDA;
    }
    attr_set = true;
  }
  const int smem_bytes = kWgStages * kWgStageBytes + 1024 + 256;
  dim3 grid(p.m_tiles * p.n_tiles * 2, p.splits);
  static int cluster_mode = -1;
  if (cluster_mode < 0) {
    const char* e = getenv("DG_WGRAD_CLUSTER");
    cluster_mode = e ? atoi(e) : 1;
  }
  p.cluster = (cluster_mode && !p.share && p.m_tiles % 2 == 0) ? 1 : 0;
  static int pair_mode = -1;   // cta_group::2 variant, see wgrad_gemm_pair_kernel
  if (pair_mode < 0) {
    const char* e = getenv("DG_WGRAD_PAIR");
    pair_mode = e ? atoi(e) : 1;
  }
  const int want_wpair = (opts && opts->wgrad_pair >= 0) ? opts->wgrad_pair : pair_mode;
  if (want_wpair && p.cluster && !p.debug) {
    static bool pair_attr_set_dev[kMaxDevices] = {};
    bool& pair_attr_set = pair_attr_set_dev[current_device()];
    if (!pair_attr_set) {
      cudaError_t e = cudaFuncSetAttribute(wgrad_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) {
        dg_set_error("wgrad: cannot raise dynamic smem: %s", cudaGetErrorString(e));
        return DG_ERR_CUDA;
      }
      pair_attr_set = true;
    }
    dg_launch(wgrad_gemm_pair_kernel, dg_cfg(grid, kThreads, kWpStages * kWpStageBytes + 1024 + 256, stream, 2), tmS, tmBig, p);
  } else {
    dg_launch(wgrad_gemm_kernel, dg_cfg(grid, kThreads, smem_bytes, stream, p.cluster ? 2 : 1), tmS, tmBig, tmBig33, p);
  }
  DG_CHECK_LAUNCH("wgrad_gemm_kernel");
  if (p.direct) return DG_OK;
  const long long total = (long long)Cs * Cb * 2;
  dg_launch(wgrad_reduce_kernel, dg_cfg(dg_ceil_div(total, 256), 256, 0, stream), p.ws, dw, beta, Cs, Cb, p.m_tiles, p.n_tiles,
                                                                   p.splits);
  DG_CHECK_LAUNCH("wgrad_reduce_kernel");
  return DG_OK;
}

// ConvTranspose2d(ci,co,4,2,1) is the same three GEMMs with the roles swapped (model.py:118-138).
int dg_convT4x4s2_fprop(const void* x_small, const void* wu, void* y_big, int B, int Hs, int Ws, int Cs, int Cb,
                        const dg_conv_opts* opts, cudaStream_t stream) {
  return dg_conv4x4s2_dgrad(x_small, wu, y_big, B, Hs, Ws, Cs, Cb, opts, stream);
}
int dg_convT4x4s2_dgrad(const void* dy_big, const void* wd, void* dx_small, int B, int H, int W, int Cb, int Cs,
                        const dg_conv_opts* opts, cudaStream_t stream) {
  return dg_conv4x4s2_fprop(dy_big, wd, dx_small, B, H, W, Cb, Cs, opts, stream);
}
int dg_convT4x4s2_wgrad(const void* x_small, const void* dy_big, float* dw, float beta, int B, int Hs, int Ws, int Cs,
                        int Cb, void* ws, size_t ws_bytes, const dg_conv_opts* opts, cudaStream_t stream) {
  return dg_conv4x4s2_wgrad(x_small, dy_big, dw, beta, B, Hs, Ws, Cs, Cb, ws, ws_bytes, opts, stream);
}

}  // extern "C"
