// Weight-gradient convolution as a tcgen05 GEMM with both operands MN-major (pixels are the reduction dimension):
//
//   dW[co][ci][kh][kw] += sum_{n,ho,wo} dy[n,ho,wo,co] * x[n, ho*s+kh-p, wo*s+kw-p, ci]
//
//   A = dy  [pixels][cout]  -> M = 128 output channels (two 64-channel SWIZZLE_128B blocks of 128 pixel-rows)
//   B = x   [pixels][cin]   -> N <= 256 input channels (64-channel blocks), one shifted TMA box per filter tap
//   D[tap]  = fp32 accumulator in TMEM, taps of a group side by side (taps_in_group * N <= 512 columns)
//
// Work unit = (M tile, N tile, tap group, pixel split). Units are distributed round-robin over a persistent grid;
// each unit reduces its share of the pixels in TMEM and adds the result into the fp32 OIHW gradient with red.global
// (parameter gradients accumulate across the two scale passes and, later, across micro-batches).
// Replaces cuDNN convolution_backward (weight part) behind every nn.Conv2d listed in SURVEY.md §2b K1-K5.
#include "ptx.cuh"
#include "tma_host.h"
#include "launch.h"
#include "../../include/b200seg.h"

namespace b200seg {

struct WgradParams {
  int N, Ho, Wo, Cout, Cin;
  int ksize, stride, pad, taps;
  int TH, TW, tiles_h, tiles_w, pix_tiles;
  int m_tiles, n_tiles, tap_groups, taps_per_group, splits, total_units;
  int nblocksB_max;
  int halo;              // 3x3 stride-1: one x halo tile per pixel tile, taps are row-shifted descriptors
  int ntile_w;           // N tile width in channels (256, or 128 in halo mode)
  int b_block_bytes;     // bytes of one 64-channel B block (16384 dense, 24576 halo)
  int unit_n;            // accumulator row pitch of a unit slab (= min(256, Cin))
  long long unit_stride; // floats per unit slab
  int a_slot_bytes, b_slot_bytes, b_slots;
};

constexpr int kWThreads = 256;
constexpr int kABlock = 128 * 128;   // 128 pixel rows x 64 channels bf16
constexpr int kMaxBSlots = 6;

__global__ void __launch_bounds__(kWThreads, 1)
wgrad_igemm_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                   const WgradParams p, float* __restrict__ ws) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_base = smem;                               // 2 slots
  uint8_t* b_base = smem + 2 * p.a_slot_bytes;          // b_slots
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + (size_t)p.b_slots * p.b_slot_bytes);
  uint64_t* a_full = bars;              // [2]
  uint64_t* a_empty = bars + 2;         // [2]
  uint64_t* b_full = bars + 4;          // [kMaxBSlots]
  uint64_t* b_empty = bars + 4 + kMaxBSlots;
  uint64_t* acc_full = bars + 4 + 2 * kMaxBSlots;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmDy); tma_prefetch_desc(&tmX); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.b_slots; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();   // the prologue above overlapped the previous kernel; global memory is touched only below

  // unit decode helpers (identical in every role)
  auto decode = [&](int unit, int& m_tile, int& n_tile, int& tg, int& split) {
    split = unit % p.splits;
    int item = unit / p.splits;
    tg = item % p.tap_groups; item /= p.tap_groups;
    n_tile = item % p.n_tiles;
    m_tile = item / p.n_tiles;
  };

  if (warp == 0) {
    if (elect_one()) {   // one elected lane runs the whole producer loop (TMA operands stay in uniform registers)
      int a_slot = 0, b_slot = 0;
      uint32_t a_phase = 0, b_phase = 0;
      for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
        int m_tile, n_tile, tg, split;
        decode(unit, m_tile, n_tile, tg, split);
        const int n0 = n_tile * p.ntile_w;
        const int nblk = min(p.ntile_w / 64, (p.Cin - n0 + 63) / 64);
        const int tap0 = tg * p.taps_per_group, tap1 = min(p.taps, tap0 + p.taps_per_group);
        const bool a_two = (p.Cout - m_tile * 128) > 64;      // second 64-channel dy block holds real channels
        for (int t = split; t < p.pix_tiles; t += p.splits) {
          const int tw_i = t % p.tiles_w;
          const int th_i = (t / p.tiles_w) % p.tiles_h;
          const int img = t / (p.tiles_w * p.tiles_h);
          mbar_wait(&a_empty[a_slot], a_phase ^ 1);
          uint8_t* sa = a_base + (size_t)a_slot * p.a_slot_bytes;
          mbar_arrive_expect_tx(&a_full[a_slot], (a_two ? 2 : 1) * kABlock);
          tma_load_4d(&tmDy, &a_full[a_slot], sa, m_tile * 128, tw_i * p.TW, th_i * p.TH, img);
          if (a_two)
            tma_load_4d(&tmDy, &a_full[a_slot], sa + kABlock, m_tile * 128 + 64, tw_i * p.TW, th_i * p.TH, img);
          if (++a_slot == 2) { a_slot = 0; a_phase ^= 1; }
          if (p.halo) {
            mbar_wait(&b_empty[b_slot], b_phase ^ 1);
            uint8_t* sb = b_base + (size_t)b_slot * p.b_slot_bytes;
            mbar_arrive_expect_tx(&b_full[b_slot], nblk * 18 * 10 * 128);
            for (int b = 0; b < nblk; ++b)
              tma_load_4d(&tmX, &b_full[b_slot], sb + (size_t)b * p.b_block_bytes, n0 + b * 64, tw_i * p.TW - 1,
                          th_i * p.TH - 1, img);
            if (++b_slot == p.b_slots) { b_slot = 0; b_phase ^= 1; }
            continue;
          }
          for (int tap = tap0; tap < tap1; ++tap) {
            const int kh = tap / p.ksize, kw = tap - kh * p.ksize;
            mbar_wait(&b_empty[b_slot], b_phase ^ 1);
            uint8_t* sb = b_base + (size_t)b_slot * p.b_slot_bytes;
            mbar_arrive_expect_tx(&b_full[b_slot], nblk * kABlock);
            for (int b = 0; b < nblk; ++b)
              tma_load_4d(&tmX, &b_full[b_slot], sb + (size_t)b * kABlock, n0 + b * 64,
                          tw_i * p.TW * p.stride + kw - p.pad, th_i * p.TH * p.stride + kh - p.pad, img);
            if (++b_slot == p.b_slots) { b_slot = 0; b_phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {   // one elected thread: barrier waits + unrolled tcgen05.mma issue for the whole CTA
      // 16 pixels per MMA = two 8-pixel tile rows: dy rows are dense (2048 B per K step); halo x rows sit 10 apart
      const uint64_t a_tmpl = make_smem_desc(0, kABlock, 1024, 2);
      const uint64_t bh_tmpl = make_smem_desc(0, p.b_block_bytes, 1280, 2);
      const uint64_t bd_tmpl = make_smem_desc(0, kABlock, 1024, 2);
      const uint32_t a0 = smem_u32(a_base) >> 4, b0 = smem_u32(b_base) >> 4;
      const uint32_t aslot16 = (uint32_t)p.a_slot_bytes >> 4, bslot16 = (uint32_t)p.b_slot_bytes >> 4;
      int a_slot = 0, b_slot = 0;
      uint32_t a_phase = 0, b_phase = 0;
      int uit = 0;
      for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x, ++uit) {
        int m_tile, n_tile, tg, split;
        decode(unit, m_tile, n_tile, tg, split);
        const int n0 = n_tile * p.ntile_w;
        const int Nn = min(p.ntile_w, p.Cin - n0);         // multiple of 16
        const uint32_t idesc = make_idesc_bf16(128, Nn, 1, 1);
        const int tap0 = tg * p.taps_per_group, tap1 = min(p.taps, tap0 + p.taps_per_group);
        mbar_wait(acc_empty, (uit & 1) ^ 1);
        tc_fence_after();
        uint32_t acc = 0;
        for (int t = split; t < p.pix_tiles; t += p.splits) {
          mbar_wait(&a_full[a_slot], a_phase);
          tc_fence_after();
          const uint64_t ad = a_tmpl + (uint64_t)(a0 + (uint32_t)a_slot * aslot16);
          if (p.halo) {
            mbar_wait(&b_full[b_slot], b_phase);
            tc_fence_after();
            const uint64_t bd = bh_tmpl + (uint64_t)(b0 + (uint32_t)b_slot * bslot16);
            uint32_t d_tmem = tmem_base;
            for (int tap = tap0; tap < tap1; ++tap, d_tmem += Nn) {
              const int kh = tap / 3, kw = tap - kh * 3;
              const uint64_t bt = bd + (uint64_t)((kh * 10 + kw) * 8);
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma_f16(d_tmem, ad + (uint64_t)(k * 128), bt + (uint64_t)(k * 160), idesc, k ? 1u : acc);
            }
            umma_commit(&b_empty[b_slot]);
            umma_commit(&a_empty[a_slot]);
            if (++b_slot == p.b_slots) { b_slot = 0; b_phase ^= 1; }
            if (++a_slot == 2) { a_slot = 0; a_phase ^= 1; }
            acc = 1;
            continue;
          }
          uint32_t d_tmem = tmem_base;
          for (int tap = tap0; tap < tap1; ++tap, d_tmem += Nn) {
            mbar_wait(&b_full[b_slot], b_phase);
            tc_fence_after();
            const uint64_t bd = bd_tmpl + (uint64_t)(b0 + (uint32_t)b_slot * bslot16);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_f16(d_tmem, ad + (uint64_t)(k * 128), bd + (uint64_t)(k * 128), idesc, k ? 1u : acc);
            umma_commit(&b_empty[b_slot]);
            if (++b_slot == p.b_slots) { b_slot = 0; b_phase ^= 1; }
          }
          umma_commit(&a_empty[a_slot]);
          if (++a_slot == 2) { a_slot = 0; a_phase ^= 1; }
          acc = 1;
        }
        umma_commit(acc_full);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const uint32_t q = warp - 4;
    int uit = 0;
    for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x, ++uit) {
      int m_tile, n_tile, tg, split;
      decode(unit, m_tile, n_tile, tg, split);
      const int n0 = n_tile * p.ntile_w;
      const int Nn = min(p.ntile_w, p.Cin - n0);
      const int tap0 = tg * p.taps_per_group, tap1 = min(p.taps, tap0 + p.taps_per_group);
      const bool has_tiles = split < p.pix_tiles;
      mbar_wait(acc_full, uit & 1);
      tc_fence_after();
      const int co = m_tile * 128 + q * 32 + lane;
      if (has_tiles) {
        for (int tap = tap0; tap < tap1; ++tap) {
          for (int ch = 0; ch < Nn / 16; ++ch) {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((q * 32u) << 16) + (tap - tap0) * Nn + ch * 16, r);
            tmem_ld_wait();
            if (co < p.Cout) {
              // per-unit slab [tap_local][128 rows][unit_n] fp32: plain 16-byte stores, summed by wgrad_reduce_kernel
              float4* dst = reinterpret_cast<float4*>(ws + (size_t)unit * p.unit_stride +
                                                      ((size_t)(tap - tap0) * 128 + q * 32 + lane) * p.unit_n + ch * 16);
              dst[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
              dst[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
              dst[2] = make_float4(__uint_as_float(r[8]), __uint_as_float(r[9]), __uint_as_float(r[10]), __uint_as_float(r[11]));
              dst[3] = make_float4(__uint_as_float(r[12]), __uint_as_float(r[13]), __uint_as_float(r[14]), __uint_as_float(r[15]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}


// dw[co][ci][tap] += sum over the pixel splits of the unit slabs (fixed order -> deterministic).
// Block = 32 consecutive ci (coalesced slab reads) x 8 split-slices; one (co, tap) row per blockIdx.y iteration.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const WgradParams p, const float* __restrict__ ws, float* __restrict__ dw) {
  __shared__ float sh[8][32];
  pdl_sync();
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int ci_blocks = (p.Cin + 31) / 32;
  const long long rows = (long long)p.Cout * p.taps * ci_blocks;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const int cb = (int)(r % ci_blocks);
    const int tap = (int)((r / ci_blocks) % p.taps);
    const int co = (int)(r / ((long long)ci_blocks * p.taps));
    const int ci = cb * 32 + lane;
    float acc = 0.f;
    if (ci < p.Cin) {
      const int m_tile = co >> 7, row = co & 127;
      const int n_tile = ci / p.ntile_w, col = ci - n_tile * p.ntile_w;
      const int tg = tap / p.taps_per_group, tl = tap - tg * p.taps_per_group;
      const int item = (m_tile * p.n_tiles + n_tile) * p.tap_groups + tg;
      const float* src = ws + (size_t)item * p.splits * p.unit_stride + ((size_t)tl * 128 + row) * p.unit_n + col;
      float a0 = 0.f, a1 = 0.f;
      int s_ = slice;
      for (; s_ + 8 < p.splits; s_ += 16) {
        a0 += src[(size_t)s_ * p.unit_stride];
        a1 += src[(size_t)(s_ + 8) * p.unit_stride];
      }
      if (s_ < p.splits) a0 += src[(size_t)s_ * p.unit_stride];
      acc = a0 + a1;
    }
    sh[slice][lane] = acc;
    __syncthreads();
    if (slice == 0 && ci < p.Cin) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += sh[k][lane];
      dw[((size_t)co * p.Cin + ci) * p.taps + tap] += t;
    }
    __syncthreads();
  }
}

}  // namespace b200seg

using namespace b200seg;

static int wgrad_plan(const b200seg_conv_desc* d, WgradParams& p) {
  if (!d) return B200SEG_E_BADARG;
  if (!((d->ksize == 1 && d->pad == 0) || (d->ksize == 3 && d->pad == 1))) return B200SEG_E_BADARG;
  if (d->stride != 1 && d->stride != 2) return B200SEG_E_BADARG;
  if (d->cin % 16 || d->x_ld % 8) return B200SEG_E_BADARG;
  p.N = d->n; p.Cout = d->cout; p.Cin = d->cin;
  p.ksize = d->ksize; p.stride = d->stride; p.pad = d->pad; p.taps = d->ksize * d->ksize;
  p.Ho = (d->h + 2 * d->pad - d->ksize) / d->stride + 1;
  p.Wo = (d->w + 2 * d->pad - d->ksize) / d->stride + 1;
  p.halo = (d->ksize == 3 && d->stride == 1 && d->reserved == 0) ? 1 : 0;
  p.ntile_w = p.halo ? 128 : 256;
  p.b_block_bytes = p.halo ? 24576 : kABlock;
  p.TW = 16; p.TH = 8;
  if (p.Wo <= 8 || p.halo) { p.TW = 8; p.TH = 16; }
  p.tiles_w = (p.Wo + p.TW - 1) / p.TW;
  p.tiles_h = (p.Ho + p.TH - 1) / p.TH;
  p.pix_tiles = d->n * p.tiles_h * p.tiles_w;
  p.m_tiles = (d->cout + 127) / 128;
  p.n_tiles = (d->cin + p.ntile_w - 1) / p.ntile_w;
  const int Nmax = d->cin < p.ntile_w ? d->cin : p.ntile_w;
  int tpg = 512 / Nmax;
  if (tpg > p.taps) tpg = p.taps;
  p.tap_groups = (p.taps + tpg - 1) / tpg;
  p.taps_per_group = (p.taps + p.tap_groups - 1) / p.tap_groups;   // balanced groups
  const int items = p.m_tiles * p.n_tiles * p.tap_groups;
  int splits = (B200SEG_MAX_CTAS + items - 1) / items;              // ~1 unit per SM
  if (splits > p.pix_tiles) splits = p.pix_tiles;
  if (splits < 1) splits = 1;
  p.splits = splits;
  p.total_units = items * splits;
  p.nblocksB_max = (Nmax + 63) / 64;
  p.unit_n = Nmax;
  p.unit_stride = (long long)p.taps_per_group * 128 * Nmax;
  p.a_slot_bytes = 2 * kABlock;
  p.b_slot_bytes = p.nblocksB_max * p.b_block_bytes;
  return 0;
}

extern "C" size_t b200seg_conv2d_wgrad_ws_bytes(const b200seg_conv_desc* d) {
  WgradParams p;
  if (wgrad_plan(d, p) != 0) return 0;
  return (size_t)p.total_units * p.unit_stride * sizeof(float);
}

extern "C" int b200seg_conv2d_wgrad(const b200seg_conv_desc* d, const void* x, const void* dy, int32_t dy_ld,
                                    float* dw_oihw, void* workspace, size_t ws_bytes, void* stream) {
  if (!d || !x || !dy || !dw_oihw || !workspace) return B200SEG_E_BADARG;
  if (dy_ld % 8 || dy_ld < 8) return B200SEG_E_BADARG;
  WgradParams p;
  int rc0 = wgrad_plan(d, p);
  if (rc0) return rc0;
  if (ws_bytes < (size_t)p.total_units * p.unit_stride * sizeof(float)) return B200SEG_E_BADARG;
  if (reinterpret_cast<uintptr_t>(workspace) & 15) return B200SEG_E_BADARG;
  const size_t fixed = 1024 + 2 * (size_t)p.a_slot_bytes + (4 + 2 * kMaxBSlots + 2) * 8 + 16;
  int bs = (int)((227 * 1024 - fixed) / p.b_slot_bytes);
  if (bs > kMaxBSlots) bs = kMaxBSlots;
  if (bs < 2) return B200SEG_E_BADARG;
  p.b_slots = bs;
  size_t smem_bytes = fixed + (size_t)bs * p.b_slot_bytes;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;

  CUtensorMap tmDy, tmX;
  {
    uint64_t dims[4] = {(uint64_t)((d->cout + 7) / 8 * 8), (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)d->n};
    uint64_t strides[3] = {(uint64_t)dy_ld * 2, (uint64_t)p.Wo * dy_ld * 2, (uint64_t)p.Ho * p.Wo * dy_ld * 2};
    uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    int rc = encode_bf16(&tmDy, dy, 4, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)d->cin, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->n};
    uint64_t strides[3] = {(uint64_t)d->x_ld * 2, (uint64_t)d->w * d->x_ld * 2, (uint64_t)d->h * d->w * d->x_ld * 2};
    uint32_t box[4] = {64, (uint32_t)(p.TW * d->stride), (uint32_t)(p.TH * d->stride), 1};
    if (p.halo) { box[1] = 10; box[2] = 18; }
    uint32_t es[4] = {1, (uint32_t)d->stride, (uint32_t)d->stride, 1};
    int rc = encode_bf16(&tmX, x, 4, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int grid = p.total_units < B200SEG_MAX_CTAS ? p.total_units : B200SEG_MAX_CTAS;
  cudaError_t e = launch_k(wgrad_igemm_kernel, dim3(grid), dim3(kWThreads), smem_bytes, (cudaStream_t)stream, tmDy, tmX, p,
                           (float*)workspace);
  if (e != cudaSuccess) return (int)e;
  long long rb = (long long)p.Cout * p.taps * ((p.Cin + 31) / 32);
  if (rb > 148 * 8) rb = 148 * 8;
  e = launch_k(wgrad_reduce_kernel, dim3((int)rb), dim3(256), 0, (cudaStream_t)stream, p, (const float*)workspace, dw_oihw);
  return e == cudaSuccess ? 0 : (int)e;
}
