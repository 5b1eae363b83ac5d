// Implicit-GEMM convolution for sm_100a.
//
//   GEMM view:  D[pixels, cout] = sum over (tap, cin-chunk) A[pixels, chunk] * B[cout, chunk]^T
//   A operand:  NHWC activations fetched by 4-D tiled TMA boxes (chunk x TW x TH x 1). A filter tap is a shifted
//               box; out-of-image coordinates are zero-filled by the TMA unit, which *is* the im2col padding.
//               Stride-2 convolutions use the tensor map's element strides.
//   B operand:  bf16 weights [cout][tap][cin] through a 3-D map (chunk x 1 x BN).
//   MMA:        tcgen05.mma cta_group::1 kind::f16, M = 128 pixels (TH x TW patch), N = BN <= 256, K = 16 per
//               instruction, fp32 accumulators double-buffered in TMEM (2 x 256 columns).
//   Epilogue:   4 warps, thread == pixel: tcgen05.ld -> (+bias) -> round to bf16 -> 16-byte stores; per-channel
//               sum / sum-of-squares of the *stored* values for training-mode BatchNorm (network/mynn.py:18-24)
//               reduced with a shuffle butterfly and kept per-warp in shared memory (deterministic order).
//   Schedule:   persistent CTAs (<= 1 per SM), static round-robin tile order, warp-specialised:
//               warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue.
#include "ptx.cuh"
#include "tma_host.h"
#include "../../include/b200seg.h"
#include "conv_common.h"

namespace b200seg {

struct ConvKParams {
  int N, Ho, Wo, Cout;
  int ksize, stride, pad;
  int cchunks, KC, BN, n_tiles;
  int TH, TW, tiles_h, tiles_w, total_tiles;
  int y_ld, out_fp32, has_bias, emit_stats, cout_pad;
  int layout_type, sbo;
  int a_bytes, b_bytes, stage_bytes, nstages;
};

constexpr int kThreads = 256;
constexpr int kMaxStages = 8;
constexpr int kAccCols = 256;   // TMEM columns per accumulator stage

__device__ __forceinline__ void store16_bf16(void* dst, const float (&v)[16]) {
  uint4 a, b;
  a.x = pack_bf16x2(v[0], v[1]);   a.y = pack_bf16x2(v[2], v[3]);
  a.z = pack_bf16x2(v[4], v[5]);   a.w = pack_bf16x2(v[6], v[7]);
  b.x = pack_bf16x2(v[8], v[9]);   b.y = pack_bf16x2(v[10], v[11]);
  b.z = pack_bf16x2(v[12], v[13]); b.w = pack_bf16x2(v[14], v[15]);
  uint4* p = reinterpret_cast<uint4*>(dst);
  p[0] = a;
  p[1] = b;
}

// Sum v[0..15] over the 32 lanes of the warp: afterwards v[0] holds the total for channel (lane & 15).
__device__ __forceinline__ void butterfly16(float (&v)[16], uint32_t lane) {
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
}

__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const ConvKParams p, void* __restrict__ y, const float* __restrict__ bias,
                  float* __restrict__ stats_partials) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A|B] (1024-aligned) | barriers | tmem ptr | stats[4][2][cout_pad]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.nstages * p.stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tfull_bar = bars + 2 * kMaxStages;
  uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  float* s_stats = reinterpret_cast<float*>(tmem_ptr_smem + 4);   // [4 warps][2][cout_pad]

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  if (p.emit_stats) {
    for (int i = threadIdx.x; i < 4 * 2 * p.cout_pad; i += kThreads) s_stats[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int kblocks = p.ksize * p.ksize * p.cchunks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
        const int tw_i = m_tile % p.tiles_w;
        const int th_i = (m_tile / p.tiles_w) % p.tiles_h;
        const int img = m_tile / (p.tiles_w * p.tiles_h);
        const int h0 = th_i * p.TH * p.stride - p.pad;
        const int w0 = tw_i * p.TW * p.stride - p.pad;
        const int n0 = n_tile * p.BN;
        for (int kh = 0; kh < p.ksize; ++kh)
          for (int kw = 0; kw < p.ksize; ++kw)
            for (int cc = 0; cc < p.cchunks; ++cc) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = stage_base + (size_t)stage * p.stage_bytes;
              uint8_t* sb = sa + p.a_bytes;
              mbar_arrive_expect_tx(&full_bar[stage], p.a_bytes + p.b_bytes);
              tma_load_4d(&tmA, &full_bar[stage], sa, cc * p.KC, w0 + kw, h0 + kh, img);
              tma_load_3d(&tmB, &full_bar[stage], sb, cc * p.KC, kh * p.ksize + kw, n0);
              if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_bf16(128, p.BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      if (lane == 0) mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
      __syncwarp();
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * kAccCols;
      for (int kb = 0; kb < kblocks; ++kb) {
        if (lane == 0) mbar_wait(&full_bar[stage], phase);
        __syncwarp();
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(stage_base + (size_t)stage * p.stage_bytes);
          const uint32_t sb = sa + p.a_bytes;
          const uint64_t adesc = make_smem_desc(sa, 16, p.sbo, p.layout_type);
          const uint64_t bdesc = make_smem_desc(sb, 16, p.sbo, p.layout_type);
          const int ksteps = p.KC >> 4;
          for (int k = 0; k < ksteps; ++k) {
            // advancing K by 16 bf16 = 32 bytes inside the swizzle atom: +2 in the 16-byte-unit address field
            umma_f16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == kblocks - 1) umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++stage == p.nstages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const uint32_t q = warp - 4;               // TMEM sub-partition == warp % 4
    const int m = q * 32 + lane;               // accumulator row == pixel within the tile
    const int th = m / p.TW, tw = m - th * p.TW;
    float* my_stats = s_stats + (size_t)q * 2 * p.cout_pad;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
      const int tw_i = m_tile % p.tiles_w;
      const int th_i = (m_tile / p.tiles_w) % p.tiles_h;
      const int img = m_tile / (p.tiles_w * p.tiles_h);
      const int ho = th_i * p.TH + th, wo = tw_i * p.TW + tw;
      const bool valid = (ho < p.Ho) && (wo < p.Wo);
      const int n0 = n_tile * p.BN;
      const size_t pix = ((size_t)img * p.Ho + ho) * p.Wo + wo;

      mbar_wait(&tfull_bar[as], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((q * 32u) << 16) + as * kAccCols;
      const int nchunks = p.BN >> 4;
      for (int ch = 0; ch < nchunks; ++ch) {
        const int c0 = n0 + ch * 16;
        if (c0 >= p.Cout) break;               // warp-uniform
        uint32_t r[16];
        tmem_ld16(taddr + ch * 16, r);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        if (p.has_bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += (c0 + j < p.Cout) ? __ldg(bias + c0 + j) : 0.f;
        }
        if (p.out_fp32) {
          if (valid) {
            float* dst = reinterpret_cast<float*>(y) + pix * p.y_ld + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.Cout) dst[j] = v[j];
          }
        } else {
          if (valid) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(y) + pix * p.y_ld + c0;
            if (c0 + 16 <= p.Cout) {
              store16_bf16(dst, v);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < p.Cout) dst[j] = __float2bfloat16_rn(v[j]);
            }
          }
          if (p.emit_stats) {
            float s1[16], s2[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float rv = valid ? bf16_round(v[j]) : 0.f;
              s1[j] = rv;
              s2[j] = rv * rv;
            }
            butterfly16(s1, lane);
            butterfly16(s2, lane);
            if (lane < 16) {
              my_stats[c0 + lane] += s1[0];
              my_stats[p.cout_pad + c0 + lane] += s2[0];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.emit_stats) {
    float* out = stats_partials + (size_t)blockIdx.x * 2 * p.cout_pad;
    for (int i = threadIdx.x; i < 2 * p.cout_pad; i += kThreads) {
      out[i] = (s_stats[i] + s_stats[2 * p.cout_pad + i]) + (s_stats[4 * p.cout_pad + i] + s_stats[6 * p.cout_pad + i]);
    }
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------- host side
int conv_plan(const b200seg_conv_desc* d, ConvPlan* pl) {
  if (!d || d->n <= 0 || d->h <= 0 || d->w <= 0) return B200SEG_E_BADARG;
  if (!((d->ksize == 1 && d->pad == 0) || (d->ksize == 3 && d->pad == 1))) return B200SEG_E_BADARG;
  if (d->stride != 1 && d->stride != 2) return B200SEG_E_BADARG;
  if (d->cin % 8 || d->x_ld % 8 || d->x_ld < d->cin) return B200SEG_E_BADARG;
  if (d->cout <= 0 || d->y_ld < d->cout) return B200SEG_E_BADARG;
  if (!d->out_fp32 && (d->y_ld % 8)) return B200SEG_E_BADARG;
  pl->Ho = (d->h + 2 * d->pad - d->ksize) / d->stride + 1;
  pl->Wo = (d->w + 2 * d->pad - d->ksize) / d->stride + 1;
  // channel chunk / swizzle width: the widest of {64, 32, 16} that tiles cin without waste (zero-fill covers 720).
  int KC = 64;
  if (d->cin % 64 != 0) {
    if (d->cin % 32 == 0 && d->cin < 256) KC = 32;
    else if (d->cin % 16 == 0 && d->cin < 128) KC = 16;
    else KC = 64;   // remainder chunk is zero-filled by TMA on both operands
  }
  if (d->reserved == 16 || d->reserved == 32 || d->reserved == 64) KC = d->reserved;   // test hook: force the chunk width
  pl->KC = KC;
  pl->cchunks = (d->cin + KC - 1) / KC;
  // N tile: whole cout when it fits one accumulator stage, else the smallest even split into <= 256-wide multiples of 16
  int cout16 = (d->cout + 15) / 16 * 16;
  int n_tiles = (cout16 + 255) / 256;
  int BN = ((cout16 / 16 + n_tiles - 1) / n_tiles) * 16;
  if (BN < 16) BN = 16;
  pl->BN = BN;
  pl->n_tiles = (cout16 + BN - 1) / BN;
  pl->cout_pad = pl->n_tiles * BN;
  // spatial patch of 128 output pixels
  int TW = 16, TH = 8;
  if (pl->Wo <= 8) { TW = 8; TH = 16; }
  pl->TW = TW; pl->TH = TH;
  pl->tiles_w = (pl->Wo + TW - 1) / TW;
  pl->tiles_h = (pl->Ho + TH - 1) / TH;
  pl->total_tiles = d->n * pl->tiles_h * pl->tiles_w * pl->n_tiles;
  pl->grid = pl->total_tiles < B200SEG_MAX_CTAS ? pl->total_tiles : B200SEG_MAX_CTAS;
  pl->a_bytes = 128 * KC * 2;
  pl->b_bytes = BN * KC * 2;
  pl->stage_bytes = (pl->a_bytes + pl->b_bytes + 1023) / 1024 * 1024;
  size_t fixed = 1024 /*align slack*/ + (2 * kMaxStages + 4) * 8 + 16 + (size_t)4 * 2 * pl->cout_pad * 4;
  int nst = (int)((227 * 1024 - fixed) / pl->stage_bytes);
  if (nst > kMaxStages) nst = kMaxStages;
  if (nst < 2) return B200SEG_E_BADARG;
  pl->nstages = nst;
  pl->smem_bytes = fixed + (size_t)nst * pl->stage_bytes;
  if (pl->smem_bytes < 120 * 1024) pl->smem_bytes = 120 * 1024;   // keep one CTA per SM: each allocates all 512 TMEM columns
  return 0;
}

}  // namespace b200seg

using namespace b200seg;

extern "C" size_t b200seg_conv2d_stats_elems(const b200seg_conv_desc* d) {
  ConvPlan pl;
  if (conv_plan(d, &pl) != 0) return 0;
  return (size_t)B200SEG_MAX_CTAS * 2 * pl.cout_pad;
}

extern "C" int b200seg_conv2d_fwd(const b200seg_conv_desc* d, const void* x, const void* w_ohwi, const float* bias,
                                  void* y, float* stats_partials, int32_t* stats_grid, void* stream) {
  ConvPlan pl;
  int rc = conv_plan(d, &pl);
  if (rc) return rc;
  if (!x || !w_ohwi || !y) return B200SEG_E_BADARG;
  if (d->has_bias && !bias) return B200SEG_E_BADARG;
  if (d->emit_stats && (!stats_partials || d->out_fp32)) return B200SEG_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w_ohwi) & 15) ||
      (reinterpret_cast<uintptr_t>(y) & 15))
    return B200SEG_E_BADARG;

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)d->cin, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->n};
    uint64_t strides[3] = {(uint64_t)d->x_ld * 2, (uint64_t)d->w * d->x_ld * 2, (uint64_t)d->h * d->w * d->x_ld * 2};
    uint32_t box[4] = {(uint32_t)pl.KC, (uint32_t)(pl.TW * d->stride), (uint32_t)(pl.TH * d->stride), 1};
    uint32_t es[4] = {1, (uint32_t)d->stride, (uint32_t)d->stride, 1};
    rc = encode_bf16(&tmA, x, 4, dims, strides, box, es, swizzle_for_bytes(pl.KC * 2));
    if (rc) return rc;
  }
  {
    const int taps = d->ksize * d->ksize;
    uint64_t dims[3] = {(uint64_t)d->cin, (uint64_t)taps, (uint64_t)d->cout};
    uint64_t strides[2] = {(uint64_t)d->cin * 2, (uint64_t)taps * d->cin * 2};
    uint32_t box[3] = {(uint32_t)pl.KC, 1, (uint32_t)pl.BN};
    rc = encode_bf16(&tmB, w_ohwi, 3, dims, strides, box, nullptr, swizzle_for_bytes(pl.KC * 2));
    if (rc) return rc;
  }
  ConvKParams p;
  p.N = d->n; p.Ho = pl.Ho; p.Wo = pl.Wo; p.Cout = d->cout;
  p.ksize = d->ksize; p.stride = d->stride; p.pad = d->pad;
  p.cchunks = pl.cchunks; p.KC = pl.KC; p.BN = pl.BN; p.n_tiles = pl.n_tiles;
  p.TH = pl.TH; p.TW = pl.TW; p.tiles_h = pl.tiles_h; p.tiles_w = pl.tiles_w; p.total_tiles = pl.total_tiles;
  p.y_ld = d->y_ld; p.out_fp32 = d->out_fp32; p.has_bias = d->has_bias; p.emit_stats = d->emit_stats;
  p.cout_pad = pl.cout_pad;
  p.layout_type = pl.KC == 64 ? 2 : (pl.KC == 32 ? 4 : 6);
  p.sbo = 8 * pl.KC * 2;
  p.a_bytes = pl.a_bytes; p.b_bytes = pl.b_bytes; p.stage_bytes = pl.stage_bytes; p.nstages = pl.nstages;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  conv_igemm_kernel<<<pl.grid, kThreads, pl.smem_bytes, (cudaStream_t)stream>>>(tmA, tmB, p, y, bias, stats_partials);
  if (stats_grid) *stats_grid = pl.grid;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
