// Hierarchical multi-scale attention blend + per-pixel cross-entropy, forward and backward, without ever
// materialising a full-resolution [N,19,H,W] fp32 logit tensor.
//
// Reference semantics (network/ocrnet.py:170-183,264-319; twin network/mscale.py:182-220; loss/utils.py:133-134):
//   per scale pass  : cls/aux/attn at 1/4 of the pass input, bilinearly upsampled x4 (Upsample, fp32, align_corners=False)
//   two-scale blend : p_lo = attn * cls_lo  at the lo-pass input size ("mid" grid), then scale_as (x2) to full size;
//                     joint = up2(p_lo) + (1 - up2(attn)) * cls_hi      (same for aux)
//   loss            : OCR_ALPHA * CE(joint_aux) + CE(joint_cls) [+ w * CE(up2(cls_lo_x4)) + w * CE(cls_hi_x4)]
// The two cascaded upsamples are kept separate (x4 then x2), exactly like the reference (SURVEY §7 hard part 5).
//
// Kernels (C = 19 classes, logits fp32 [pixels][20], class-gradient outputs bf16 [pixels][32] zero padded):
//   mid_fwd      lo quarter maps -> mid buffer M[m][40] = {attn4*cls4 (19), attn4*aux4 (19), attn4, 0} (+ cls4 for the
//                supervised term)
//   loss_fwd     per full-res pixel: gather 4 mid taps + 4 hi taps, blend, log-softmax / NLL for each head, write the
//                per-pixel gradients w.r.t. the hi path (Ghi), the lo path (Glo) and block-partial loss sums
//   hi_bwd       adjoint of the x4 upsample: Ghi -> d cls_hi, d aux_hi
//   mid_bwd      adjoint of the x2 upsample + product rule at the mid grid -> D[m][40]
//   lo_bwd       adjoint of the x4 upsample on the lo pass -> d cls_lo, d aux_lo, d attn_logit (through the sigmoid)
#include "ptx.cuh"
#include "launch.h"
#include "../../include/b200seg.h"
#include "vec.cuh"

namespace b200seg {

constexpr int NC = 19;      // classes
constexpr int LD = 20;      // logits pitch (floats)
constexpr int MW = 40;      // mid / gradient buffer width

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

struct Taps {
  int i00, i01, i10, i11;   // linear pixel indices within the image
  float w00, w01, w10, w11;
};
__device__ __forceinline__ Taps make_taps(int Y, int X, int H, int W, int h, int w) {
  int y0, y1, x0, x1;
  float ly, lx;
  bilinear_src(Y, (float)h / (float)H, h, y0, y1, ly);
  bilinear_src(X, (float)w / (float)W, w, x0, x1, lx);
  Taps t;
  t.i00 = y0 * w + x0; t.i01 = y0 * w + x1; t.i10 = y1 * w + x0; t.i11 = y1 * w + x1;
  const float hy = 1.f - ly, hx = 1.f - lx;
  t.w00 = hy * hx; t.w01 = hy * lx; t.w10 = ly * hx; t.w11 = ly * lx;
  return t;
}
// PyTorch evaluates h0*(w0*a + w1*b) + h1*(w0*c + w1*d); keep that association.
__device__ __forceinline__ float tap_eval(const Taps& t, float a, float b, float c, float d, float hy, float ly, float hx,
                                          float lx) {
  return hy * (hx * a + lx * b) + ly * (hx * c + lx * d);
}

struct TapW {   // separable weights kept for the PyTorch association order
  int y0, y1, x0, x1;
  float hy, ly, hx, lx;
};
__device__ __forceinline__ TapW make_tapw(int Y, int X, int H, int W, int h, int w) {
  TapW t;
  bilinear_src(Y, (float)h / (float)H, h, t.y0, t.y1, t.ly);
  bilinear_src(X, (float)w / (float)W, w, t.x0, t.x1, t.lx);
  t.hy = 1.f - t.ly; t.hx = 1.f - t.lx;
  return t;
}
__device__ __forceinline__ float interp(const TapW& t, const float* __restrict__ base, int w, int ld, int c) {
  const float a = base[((size_t)t.y0 * w + t.x0) * ld + c], b = base[((size_t)t.y0 * w + t.x1) * ld + c];
  const float cc = base[((size_t)t.y1 * w + t.x0) * ld + c], d = base[((size_t)t.y1 * w + t.x1) * ld + c];
  return t.hy * (t.hx * a + t.lx * b) + t.ly * (t.hx * cc + t.lx * d);
}

// adjoint helpers: fine index range that can touch coarse index y, and the weight of coarse y for fine Y
__device__ __forceinline__ void adj_range(int y, int h, int H, int& lo, int& hi) {
  const float r = (float)H / (float)h;
  lo = (int)floorf(r * (y - 1)) - 1;
  hi = (int)ceilf(r * (y + 2)) + 1;
  if (y == 0 || lo < 0) lo = 0;
  if (y == h - 1 || hi > H) hi = H;
}
__device__ __forceinline__ float adj_weight(int Y, int y, int h, int H) {
  int y0, y1;
  float l;
  bilinear_src(Y, (float)h / (float)H, h, y0, y1, l);
  return (y0 == y ? 1.f - l : 0.f) + (y1 == y ? l : 0.f);
}

struct MsGeom {
  int N, H, W;        // full resolution (labels)
  int Hq, Wq;         // hi-pass quarter maps
  int Hm, Wm;         // mid grid (= lo-pass input size); 0 when there is no lo pass
  int Hl, Wl;         // lo-pass quarter maps
  int nheads;         // 1 (cls only) or 2 (cls, aux)
  float w_head0, w_head1, sup_wt;
  int ignore_index;
};

// ------------------------------------------------------------------------------------------------ mid_fwd
__global__ void __launch_bounds__(256)
mid_fwd_kernel(const MsGeom g, const float* __restrict__ lo_cls, const float* __restrict__ lo_aux,
               const float* __restrict__ lo_attn, float* __restrict__ M, float* __restrict__ M2) {
  pdl_sync();
  const long long total = (long long)g.N * g.Hm * g.Wm;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(idx % g.Wm), Y = (int)((idx / g.Wm) % g.Hm), n = (int)(idx / ((long long)g.Wm * g.Hm));
    const TapW t = make_tapw(Y, X, g.Hm, g.Wm, g.Hl, g.Wl);
    const size_t img = (size_t)n * g.Hl * g.Wl;
    const float* at = lo_attn + img;
    const float a00 = sigmoidf_(at[t.y0 * g.Wl + t.x0]), a01 = sigmoidf_(at[t.y0 * g.Wl + t.x1]);
    const float a10 = sigmoidf_(at[t.y1 * g.Wl + t.x0]), a11 = sigmoidf_(at[t.y1 * g.Wl + t.x1]);
    const float A4 = t.hy * (t.hx * a00 + t.lx * a01) + t.ly * (t.hx * a10 + t.lx * a11);
    float* out = M + idx * MW;
    const float* cb = lo_cls + img * LD;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float v = interp(t, cb, g.Wl, LD, c);
      out[c] = A4 * v;
      if (M2) M2[idx * LD + c] = v;
    }
    if (g.nheads > 1) {
      const float* ab = lo_aux + img * LD;
#pragma unroll
      for (int c = 0; c < NC; ++c) out[NC + c] = A4 * interp(t, ab, g.Wl, LD, c);
    } else {
#pragma unroll
      for (int c = 0; c < NC; ++c) out[NC + c] = 0.f;
    }
    out[2 * NC] = A4;
    out[2 * NC + 1] = 0.f;
  }
}

// log-softmax / NLL of one 19-vector; returns nll (0 if invalid) and overwrites J with (softmax - onehot) * coef
__device__ __forceinline__ float ce_grad(float (&J)[NC], int label, bool valid, float coef) {
  float m = J[0];
#pragma unroll
  for (int c = 1; c < NC; ++c) m = fmaxf(m, J[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) s += expf(J[c] - m);
  const float lse = m + logf(s);
  float nll = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float pr = expf(J[c] - lse);
    if (c == label) nll = lse - J[c];
    J[c] = valid ? (pr - (c == label ? 1.f : 0.f)) * coef : 0.f;
  }
  return valid ? nll : 0.f;
}

__device__ __forceinline__ void store_row40(__nv_bfloat16* dst, const float (&a)[NC], const float (&b)[NC], float e38,
                                            float e39) {
  float v[MW];
#pragma unroll
  for (int c = 0; c < NC; ++c) { v[c] = a[c]; v[NC + c] = b[c]; }
  v[38] = e38; v[39] = e39;
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = v[q * 8 + j];
    store8(dst + q * 8, o);
  }
}

// ------------------------------------------------------------------------------------------------ loss_fwd
// partial[block][4] = sums of nll over valid pixels for {cls head, aux head, supervised lo, supervised hi}
__global__ void __launch_bounds__(128)
loss_fwd_kernel(const MsGeom g, const long long* __restrict__ labels, const float* __restrict__ inv_count,
                const float* __restrict__ hi_cls, const float* __restrict__ hi_aux, const float* __restrict__ M,
                const float* __restrict__ M2, __nv_bfloat16* __restrict__ Ghi, __nv_bfloat16* __restrict__ Glo,
                __nv_bfloat16* __restrict__ Gsup, float* __restrict__ partial) {
  pdl_sync();
  const long long total = (long long)g.N * g.H * g.W;
  const bool has_lo = g.Hm > 0;
  const float icnt = *inv_count;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(idx % g.W), Y = (int)((idx / g.W) % g.H), n = (int)(idx / ((long long)g.W * g.H));
    const long long lab64 = labels[idx];
    const bool valid = lab64 != (long long)g.ignore_index;
    const int label = valid ? (int)lab64 : -1;
    const TapW th = make_tapw(Y, X, g.H, g.W, g.Hq, g.Wq);
    const size_t imq = (size_t)n * g.Hq * g.Wq;
    float hc[NC], ha[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) hc[c] = interp(th, hi_cls + imq * LD, g.Wq, LD, c);
    if (g.nheads > 1) {
#pragma unroll
      for (int c = 0; c < NC; ++c) ha[c] = interp(th, hi_aux + imq * LD, g.Wq, LD, c);
    } else {
#pragma unroll
      for (int c = 0; c < NC; ++c) ha[c] = 0.f;
    }
    float oma = 1.f;   // 1 - up2(attn)
    float Jc[NC], Ja[NC];
    TapW tm;
    size_t imm = 0;
    if (has_lo) {
      tm = make_tapw(Y, X, g.H, g.W, g.Hm, g.Wm);
      imm = (size_t)n * g.Hm * g.Wm;
      const float Aup = interp(tm, M + imm * MW, g.Wm, MW, 2 * NC);
      oma = 1.f - Aup;
#pragma unroll
      for (int c = 0; c < NC; ++c) Jc[c] = interp(tm, M + imm * MW, g.Wm, MW, c) + oma * hc[c];
      if (g.nheads > 1) {
#pragma unroll
        for (int c = 0; c < NC; ++c) Ja[c] = interp(tm, M + imm * MW, g.Wm, MW, NC + c) + oma * ha[c];
      }
    } else {
#pragma unroll
      for (int c = 0; c < NC; ++c) { Jc[c] = hc[c]; Ja[c] = ha[c]; }
    }
    acc[0] += ce_grad(Jc, label, valid, g.w_head0 * icnt);
    if (g.nheads > 1) acc[1] += ce_grad(Ja, label, valid, g.w_head1 * icnt);
    else {
#pragma unroll
      for (int c = 0; c < NC; ++c) Ja[c] = 0.f;
    }
    // lo-path gradients: dP = gJ, dA = -sum hi*gJ
    if (has_lo) {
      float gA = 0.f;
#pragma unroll
      for (int c = 0; c < NC; ++c) gA -= hc[c] * Jc[c] + ha[c] * Ja[c];
      store_row40(Glo + idx * MW, Jc, Ja, gA, 0.f);
    }
    // hi-path gradients: (1 - Aup) * gJ (+ supervised term on the hi prediction alone)
    float gh[NC], gha[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { gh[c] = oma * Jc[c]; gha[c] = oma * Ja[c]; }
    if (g.sup_wt != 0.f) {
      float S[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) S[c] = hc[c];
      acc[3] += ce_grad(S, label, valid, g.sup_wt * icnt);
#pragma unroll
      for (int c = 0; c < NC; ++c) gh[c] += S[c];
      if (has_lo) {
#pragma unroll
        for (int c = 0; c < NC; ++c) S[c] = interp(tm, M2 + imm * LD, g.Wm, LD, c);
        acc[2] += ce_grad(S, label, valid, g.sup_wt * icnt);
        float z[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) z[c] = 0.f;
        store_row40(Gsup + idx * MW, S, z, 0.f, 0.f);
      }
    }
    store_row40(Ghi + idx * MW, gh, gha, 0.f, oma);
  }
  __shared__ float s_acc[4][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float a = acc[k];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if (lane == 0) s_acc[warp][k] = a;
  }
  __syncthreads();
  if (threadIdx.x < 4)
    partial[(size_t)blockIdx.x * 4 + threadIdx.x] =
        (s_acc[0][threadIdx.x] + s_acc[1][threadIdx.x]) + (s_acc[2][threadIdx.x] + s_acc[3][threadIdx.x]);
}

// loss = icnt * (w0*s0 + w1*s1 + sup*(s2 + s3)); also returns the four mean NLLs
__global__ void loss_finalize_kernel(const float* __restrict__ partial, int nblocks, const float* __restrict__ inv_count,
                                     float w0, float w1, float sup, float* __restrict__ out) {
  pdl_sync();
  __shared__ double s[4][32];
  double a[4] = {0, 0, 0, 0};
  for (int b = threadIdx.x; b < nblocks; b += 32)
    for (int k = 0; k < 4; ++k) a[k] += (double)partial[(size_t)b * 4 + k];
  for (int k = 0; k < 4; ++k) s[k][threadIdx.x] = a[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[4] = {0, 0, 0, 0};
    for (int k = 0; k < 4; ++k)
      for (int i = 0; i < 32; ++i) t[k] += s[k][i];
    const double ic = (double)*inv_count;
    out[0] = (float)(ic * (w0 * t[0] + w1 * t[1] + sup * (t[2] + t[3])));
    for (int k = 0; k < 4; ++k) out[1 + k] = (float)(ic * t[k]);
  }
}

__global__ void count_valid_kernel(const long long* __restrict__ labels, long long total, int ignore_index,
                                   unsigned long long* __restrict__ counter) {
  pdl_sync();
  unsigned int c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    c += labels[i] != (long long)ignore_index;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(counter, (unsigned long long)c);   // integer: order independent
}
__global__ void inv_count_kernel(const unsigned long long* __restrict__ counter, float* __restrict__ inv_count) {
  pdl_sync();
  *inv_count = 1.f / (float)(*counter);    // mean over non-ignored pixels; all-ignored -> inf/NaN like the reference
}

// ------------------------------------------------------------------------------------------------ hi_bwd
// d hi[q][c] = sum_X w(X,q) * Ghi[X][c]     (one thread per quarter pixel and head)
__global__ void __launch_bounds__(128)
hi_bwd_kernel(const MsGeom g, const __nv_bfloat16* __restrict__ Ghi, __nv_bfloat16* __restrict__ d_cls,
              __nv_bfloat16* __restrict__ d_aux) {
  pdl_sync();
  const long long total = (long long)g.N * g.Hq * g.Wq * g.nheads;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int head = (int)(idx % g.nheads);
    const long long q = idx / g.nheads;
    const int x = (int)(q % g.Wq), y = (int)((q / g.Wq) % g.Hq), n = (int)(q / ((long long)g.Wq * g.Hq));
    float acc[24];
#pragma unroll
    for (int c = 0; c < 24; ++c) acc[c] = 0.f;
    int Ylo, Yhi, Xlo, Xhi;
    adj_range(y, g.Hq, g.H, Ylo, Yhi);
    adj_range(x, g.Wq, g.W, Xlo, Xhi);
    for (int Y = Ylo; Y < Yhi; ++Y) {
      const float wy = adj_weight(Y, y, g.Hq, g.H);
      if (wy == 0.f) continue;
      for (int X = Xlo; X < Xhi; ++X) {
        const float wx = adj_weight(X, x, g.Wq, g.W);
        if (wx == 0.f) continue;
        const __nv_bfloat16* src = Ghi + (((size_t)n * g.H + Y) * g.W + X) * MW;
        const float wgt = wy * wx;
        // head 0 occupies [0,19), head 1 [19,38): read the three 8-wide groups covering the head
        const int base = head == 0 ? 0 : 16;
        float v[24];
        float t8[8];
        load8(src + base, t8);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = t8[j];
        load8(src + base + 8, t8);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 + j] = t8[j];
        load8(src + base + 16, t8);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[16 + j] = t8[j];
#pragma unroll
        for (int c = 0; c < 24; ++c) acc[c] += wgt * v[c];
      }
    }
    // head 0: channels acc[0..18]; head 1: acc[3..21]
    float o[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) o[c] = 0.f;
    if (head == 0) {
#pragma unroll
      for (int c = 0; c < NC; ++c) o[c] = acc[c];
    } else {
#pragma unroll
      for (int c = 0; c < NC; ++c) o[c] = acc[c + 3];
    }
    __nv_bfloat16* dst = (head == 0 ? d_cls : d_aux) + q * 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float t8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t8[j] = o[k * 8 + j];
      store8(dst + k * 8, t8);
    }
  }
}

// ------------------------------------------------------------------------------------------------ mid_bwd
// D[m][c] = A4[m] * dP[m][c] (+ supervised d cls4), D[m][38] = sum_c C4[m][c]*dP[m][c] + dA[m]
__global__ void __launch_bounds__(128)
mid_bwd_kernel(const MsGeom g, const __nv_bfloat16* __restrict__ Glo, const __nv_bfloat16* __restrict__ Gsup,
               const float* __restrict__ lo_cls, const float* __restrict__ lo_aux, const float* __restrict__ M,
               float* __restrict__ D) {
  pdl_sync();
  const long long total = (long long)g.N * g.Hm * g.Wm;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % g.Wm), y = (int)((idx / g.Wm) % g.Hm), n = (int)(idx / ((long long)g.Wm * g.Hm));
    float dP[MW];
    float dS[24];
#pragma unroll
    for (int c = 0; c < MW; ++c) dP[c] = 0.f;
#pragma unroll
    for (int c = 0; c < 24; ++c) dS[c] = 0.f;
    int Ylo, Yhi, Xlo, Xhi;
    adj_range(y, g.Hm, g.H, Ylo, Yhi);
    adj_range(x, g.Wm, g.W, Xlo, Xhi);
    for (int Y = Ylo; Y < Yhi; ++Y) {
      const float wy = adj_weight(Y, y, g.Hm, g.H);
      if (wy == 0.f) continue;
      for (int X = Xlo; X < Xhi; ++X) {
        const float wx = adj_weight(X, x, g.Wm, g.W);
        if (wx == 0.f) continue;
        const size_t fp = ((size_t)n * g.H + Y) * g.W + X;
        const float wgt = wy * wx;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          float t8[8];
          load8(Glo + fp * MW + k * 8, t8);
#pragma unroll
          for (int j = 0; j < 8; ++j) dP[k * 8 + j] += wgt * t8[j];
        }
        if (Gsup) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            float t8[8];
            load8(Gsup + fp * MW + k * 8, t8);
#pragma unroll
            for (int j = 0; j < 8; ++j) dS[k * 8 + j] += wgt * t8[j];
          }
        }
      }
    }
    const float A4 = M[idx * MW + 2 * NC];
    const TapW t = make_tapw(y, x, g.Hm, g.Wm, g.Hl, g.Wl);
    const size_t img = (size_t)n * g.Hl * g.Wl;
    float dA = dP[2 * NC];
    float* out = D + idx * MW;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float C4 = interp(t, lo_cls + img * LD, g.Wl, LD, c);
      dA += C4 * dP[c];
      out[c] = A4 * dP[c] + dS[c];
    }
    if (g.nheads > 1) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float C4 = interp(t, lo_aux + img * LD, g.Wl, LD, c);
        dA += C4 * dP[NC + c];
        out[NC + c] = A4 * dP[NC + c];
      }
    } else {
#pragma unroll
      for (int c = 0; c < NC; ++c) out[NC + c] = 0.f;
    }
    out[2 * NC] = dA;
    out[2 * NC + 1] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ lo_bwd
__global__ void __launch_bounds__(128)
lo_bwd_kernel(const MsGeom g, const float* __restrict__ D, const float* __restrict__ lo_attn,
              __nv_bfloat16* __restrict__ d_cls, __nv_bfloat16* __restrict__ d_aux, __nv_bfloat16* __restrict__ d_attn) {
  pdl_sync();
  const long long total = (long long)g.N * g.Hl * g.Wl;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % g.Wl), y = (int)((idx / g.Wl) % g.Hl), n = (int)(idx / ((long long)g.Wl * g.Hl));
    float acc[MW];
#pragma unroll
    for (int c = 0; c < MW; ++c) acc[c] = 0.f;
    int Ylo, Yhi, Xlo, Xhi;
    adj_range(y, g.Hl, g.Hm, Ylo, Yhi);
    adj_range(x, g.Wl, g.Wm, Xlo, Xhi);
    for (int Y = Ylo; Y < Yhi; ++Y) {
      const float wy = adj_weight(Y, y, g.Hl, g.Hm);
      if (wy == 0.f) continue;
      for (int X = Xlo; X < Xhi; ++X) {
        const float wx = adj_weight(X, x, g.Wl, g.Wm);
        if (wx == 0.f) continue;
        const float4* src = reinterpret_cast<const float4*>(D + (((size_t)n * g.Hm + Y) * g.Wm + X) * MW);
        const float wgt = wy * wx;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          const float4 v = src[k];
          acc[k * 4] += wgt * v.x; acc[k * 4 + 1] += wgt * v.y; acc[k * 4 + 2] += wgt * v.z; acc[k * 4 + 3] += wgt * v.w;
        }
      }
    }
    float o[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) o[c] = c < NC ? acc[c] : 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float t8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t8[j] = o[k * 8 + j];
      store8(d_cls + idx * 32 + k * 8, t8);
    }
    if (g.nheads > 1) {
#pragma unroll
      for (int c = 0; c < 32; ++c) o[c] = c < NC ? acc[NC + c] : 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float t8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t8[j] = o[k * 8 + j];
        store8(d_aux + idx * 32 + k * 8, t8);
      }
    }
    const float s = sigmoidf_(lo_attn[idx]);
    float t8[8] = {acc[2 * NC] * s * (1.f - s), 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    store8(d_attn + idx * 8, t8);
  }
}

static inline int blocks_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

static MsGeom to_geom(const b200seg_mscale_desc* d) {
  MsGeom g;
  g.N = d->n; g.H = d->h; g.W = d->w; g.Hq = d->hq; g.Wq = d->wq; g.Hm = d->hm; g.Wm = d->wm; g.Hl = d->hl; g.Wl = d->wl;
  g.nheads = d->nheads; g.w_head0 = d->w_head0; g.w_head1 = d->w_head1; g.sup_wt = d->sup_wt;
  g.ignore_index = d->ignore_index;
  return g;
}

}  // namespace b200seg

using namespace b200seg;

#define RET_LAUNCH()                        \
  do {                                      \
    cudaError_t e_ = cudaGetLastError();    \
    return e_ == cudaSuccess ? 0 : (int)e_; \
  } while (0)

extern "C" int32_t b200seg_mscale_loss_blocks(const b200seg_mscale_desc* d) {
  return blocks_for((long long)d->n * d->h * d->w, 128);
}

extern "C" int b200seg_count_valid(const int64_t* labels, int64_t total, int32_t ignore_index, uint64_t* counter_ws,
                                   float* inv_count, void* stream) {
  if (!labels || !counter_ws || !inv_count) return B200SEG_E_BADARG;
  cudaError_t e = cudaMemsetAsync(counter_ws, 0, sizeof(uint64_t), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  launch_k(count_valid_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const long long*)labels, total,
                                                                              ignore_index,
                                                                              (unsigned long long*)counter_ws);
  launch_k(inv_count_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, (const unsigned long long*)counter_ws, inv_count);
  RET_LAUNCH();
}

extern "C" int b200seg_mscale_mid_fwd(const b200seg_mscale_desc* d, const float* lo_cls, const float* lo_aux,
                                      const float* lo_attn_logit, float* mid, float* mid_sup, void* stream) {
  if (!d || !lo_cls || !lo_attn_logit || !mid || d->hm <= 0 || (d->nheads > 1 && !lo_aux)) return B200SEG_E_BADARG;
  launch_k(mid_fwd_kernel, dim3(blocks_for((long long)d->n * d->hm * d->wm, 256)), dim3(256), 0, (cudaStream_t)stream, to_geom(d), lo_cls, lo_aux, lo_attn_logit, mid, mid_sup);
  RET_LAUNCH();
}

extern "C" int b200seg_mscale_loss_fwd(const b200seg_mscale_desc* d, const int64_t* labels, const float* inv_count,
                                       const float* hi_cls, const float* hi_aux, const float* mid, const float* mid_sup,
                                       void* g_hi, void* g_lo, void* g_sup, float* partial_ws, float* loss_out,
                                       void* stream) {
  if (!d || !labels || !inv_count || !hi_cls || !g_hi || !partial_ws || !loss_out) return B200SEG_E_BADARG;
  if (d->nheads > 1 && !hi_aux) return B200SEG_E_BADARG;
  if (d->hm > 0 && (!mid || !g_lo)) return B200SEG_E_BADARG;
  if (d->sup_wt != 0.f && d->hm > 0 && (!mid_sup || !g_sup)) return B200SEG_E_BADARG;
  const int nb = b200seg_mscale_loss_blocks(d);
  launch_k(loss_fwd_kernel, dim3(nb), dim3(128), 0, (cudaStream_t)stream, to_geom(d), (const long long*)labels, inv_count, hi_cls, hi_aux,
                                                        mid, mid_sup, (__nv_bfloat16*)g_hi, (__nv_bfloat16*)g_lo,
                                                        (__nv_bfloat16*)g_sup, partial_ws);
  launch_k(loss_finalize_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, partial_ws, nb, inv_count, d->w_head0,
                                                           d->nheads > 1 ? d->w_head1 : 0.f, d->sup_wt, loss_out);
  RET_LAUNCH();
}

extern "C" int b200seg_mscale_hi_bwd(const b200seg_mscale_desc* d, const void* g_hi, void* d_cls, void* d_aux,
                                     void* stream) {
  if (!d || !g_hi || !d_cls || (d->nheads > 1 && !d_aux)) return B200SEG_E_BADARG;
  launch_k(hi_bwd_kernel, dim3(blocks_for((long long)d->n * d->hq * d->wq * d->nheads, 128)), dim3(128), 0, (cudaStream_t)stream, to_geom(d), (const __nv_bfloat16*)g_hi, (__nv_bfloat16*)d_cls, (__nv_bfloat16*)d_aux);
  RET_LAUNCH();
}

extern "C" int b200seg_mscale_lo_bwd(const b200seg_mscale_desc* d, const void* g_lo, const void* g_sup,
                                     const float* lo_cls, const float* lo_aux, const float* lo_attn_logit,
                                     const float* mid, float* dmid_ws, void* d_cls, void* d_aux, void* d_attn,
                                     void* stream) {
  if (!d || !g_lo || !lo_cls || !lo_attn_logit || !mid || !dmid_ws || !d_cls || !d_attn || d->hm <= 0)
    return B200SEG_E_BADARG;
  const MsGeom g = to_geom(d);
  launch_k(mid_bwd_kernel, dim3(blocks_for((long long)d->n * d->hm * d->wm, 128)), dim3(128), 0, (cudaStream_t)stream, g, (const __nv_bfloat16*)g_lo, d->sup_wt != 0.f ? (const __nv_bfloat16*)g_sup : nullptr, lo_cls, lo_aux, mid,
      dmid_ws);
  launch_k(lo_bwd_kernel, dim3(blocks_for((long long)d->n * d->hl * d->wl, 128)), dim3(128), 0, (cudaStream_t)stream, g, dmid_ws, lo_attn_logit, (__nv_bfloat16*)d_cls, (__nv_bfloat16*)d_aux, (__nv_bfloat16*)d_attn);
  RET_LAUNCH();
}
