/*
 * b200seg.h — C ABI of libb200seg.so: the sm_100a kernels behind the HRNet-OCR-MScale hot path of
 * NVIDIA/semantic-segmentation (reference @ /root/reference).
 *
 * Conventions (SURVEY.md §8b "What the C-ABI replacement must export"):
 *   - plain pointers and sizes only; every buffer (including workspaces) is owned by the caller;
 *     the library never allocates, frees or retains a pointer past return;
 *   - every entry point takes the CUDA stream it must launch on (cudaStream_t passed as void*), never
 *     synchronises the device and never touches the default stream, so calls are CUDA-graph capturable;
 *   - return value: 0 = ok, negative = B200SEG_E_* argument/driver-entry error, positive = cudaError_t;
 *   - activations are NHWC ("channels-last") bf16 with an explicit pixel pitch `ld` (elements between
 *     consecutive pixels, >= channels, multiple of 8) so channel slices of a wider buffer are addressable
 *     (concat == channel offset); statistics, logits, losses are fp32; labels int64.
 *
 * Each entry point cites the reference call site (file:line under /root/reference) whose library call it replaces.
 */
#ifndef B200SEG_H_
#define B200SEG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SEG_E_BADARG (-1)      /* shape / alignment not supported by the kernel */
#define B200SEG_E_NODRIVER (-100)  /* cuTensorMapEncodeTiled entry point unavailable */
#define B200SEG_MAX_CTAS 148       /* persistent grids never exceed one CTA per SM */

/* ABI / build identification. */
int b200seg_abi_version(void);
const char* b200seg_build_info(void);

/* ------------------------------------------------------------------------------------------------
 * Convolution as implicit GEMM on tcgen05 (TMA-fed, TMEM accumulators).
 * Replaces cuDNN behind nn.Conv2d at network/hrnetv2.py:31-34,74-80,193-198,208-210,270-274,
 * network/ocrnet.py:54-57,66-76, network/ocr_utils.py:68-91,142-144, network/utils.py:348-362.
 * ------------------------------------------------------------------------------------------------ */
typedef struct b200seg_conv_desc {
  int32_t n, h, w;         /* INPUT batch / height / width */
  int32_t cin, cout;       /* cin multiple of 8 */
  int32_t ksize;           /* 1 or 3 (square) */
  int32_t stride;          /* 1 or 2 */
  int32_t pad;             /* 0 for 1x1, 1 for 3x3 */
  int32_t x_ld;            /* input pixel pitch (elements) */
  int32_t y_ld;            /* output pixel pitch (elements) */
  int32_t out_fp32;        /* 0: bf16 output, 1: fp32 output (logit heads) */
  int32_t has_bias;        /* bias[cout] fp32 added before rounding */
  int32_t emit_stats;      /* write per-CTA per-channel sum / sum-of-squares partials of the stored output */
  int32_t reserved;
} b200seg_conv_desc;

/* Number of fp32 elements the stats partial buffer must hold: B200SEG_MAX_CTAS * 2 * cout_padded. */
size_t b200seg_conv2d_stats_elems(const b200seg_conv_desc* d);

/* y[n,ho,wo,co] = sum_{kh,kw,ci} x[n, ho*s+kh-p, wo*s+kw-p, ci] * w[co,kh,kw,ci] (+ bias[co]).
 * x: bf16 NHWC (pitch x_ld); w: bf16 [cout][k*k][cin] ("OHWI", see b200seg_pack_weight); y: bf16|fp32 NHWC.
 * stats (optional): fp32 [grid][2][cout_pad]; *stats_grid receives the number of valid partial rows. */
int b200seg_conv2d_fwd(const b200seg_conv_desc* d, const void* x, const void* w_ohwi, const float* bias, void* y,
                       float* stats_partials, int32_t* stats_grid, void* stream);

/* Slow, obviously-correct CUDA-core direct convolution with identical numerics contract (fp32 accumulate, one rounding).
 * Used by the GPU test-suite as an on-device cross-check and for shapes the GEMM path does not take. */
int b200seg_conv2d_fwd_direct(const b200seg_conv_desc* d, const void* x, const void* w_ohwi, const float* bias,
                              void* y, void* stream);

/* Repack fp32 OIHW master weights (the nn.Parameter layout the reference checkpoints use) into the kernel layouts:
 *   w_ohwi  bf16 [O][kh*kw][I]            forward operand
 *   w_dgrad bf16 [I][kh*kw (flipped)][O]  data-gradient operand (may be NULL) */
int b200seg_pack_weight(const float* w_oihw, int32_t o, int32_t i, int32_t ksize, void* w_ohwi, void* w_dgrad,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SEG_H_ */
