// Hardware probe: one CTA, one tcgen05.mma chain with fully caller-controlled shared-memory descriptors.
// Used only by tests/probes to pin down descriptor semantics (swizzle phase / base offset, MN-major operands,
// stride-byte-offset) on real silicon before a kernel design depends on them. Not on the product path.
#include "ptx.cuh"
#include "tma_host.h"
#include "../../include/b200seg.h"
#include "probe.h"

namespace b200seg {

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const b200seg_probe_desc p, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + 96 * 1024;
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_ptr;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&tmem_ptr, 256); tmem_relinquish(); }
  // poison shared memory so stale data is recognisable
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x7fc07fc0u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t a_box_bytes = p.a.box_cols * p.a.box_rows * 2, b_box_bytes = p.b.box_cols * p.b.box_rows * 2;
    mbar_arrive_expect_tx(&bar_full, a_box_bytes * p.a.nboxes + b_box_bytes * p.b.nboxes);
    for (int i = 0; i < p.a.nboxes; ++i)
      tma_load_2d(&tmA, &bar_full, sA + (size_t)i * p.a.smem_stride, p.a.c0 + i * p.a.dcol, p.a.r0 + i * p.a.drow);
    for (int i = 0; i < p.b.nboxes; ++i)
      tma_load_2d(&tmB, &bar_full, sB + (size_t)i * p.b.smem_stride, p.b.c0 + i * p.b.dcol, p.b.r0 + i * p.b.drow);
    mbar_wait(&bar_full, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(p.M, p.N, p.a_major, p.b_major);
    for (int k = 0; k < p.ksteps; ++k) {
      const uint64_t ad = make_smem_desc(smem_u32(sA) + p.a_off + k * p.a_kstep, p.a_lbo, p.a_sbo, p.a_layout, p.a_base);
      const uint64_t bd = make_smem_desc(smem_u32(sB) + p.b_off + k * p.b_kstep, p.b_lbo, p.b_sbo, p.b_layout, p.b_base);
      umma_f16(tmem_base, ad, bd, idesc, k != 0);
    }
    umma_commit(&bar_mma);
  }
  __syncthreads();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int ch = 0; ch < p.N / 16; ++ch) {
    uint32_t r[16];
    tmem_ld16(tmem_base + ((warp * 32u) << 16) + ch * 16, r);
    tmem_ld_wait();
    float* dst = D + (size_t)(warp * 32 + lane) * p.N + ch * 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) dst[j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

static int make_map(CUtensorMap* m, const void* base, const b200seg_probe_operand& o) {
  uint64_t dims[2] = {(uint64_t)o.cols, (uint64_t)o.rows};
  uint64_t strides[1] = {(uint64_t)o.cols * 2};
  uint32_t box[2] = {(uint32_t)o.box_cols, (uint32_t)o.box_rows};
  return encode_bf16(m, base, 2, dims, strides, box, nullptr, swizzle_for_bytes(o.swizzle_bytes));
}

}  // namespace b200seg

using namespace b200seg;

extern "C" int b200seg_umma_probe(const b200seg_probe_desc* p, const void* A, const void* B, float* D, void* stream) {
  if (!p || !A || !B || !D) return B200SEG_E_BADARG;
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, A, p->a);
  if (rc) return rc;
  rc = make_map(&tmB, B, p->b);
  if (rc) return rc;
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return (int)e;
  umma_probe_kernel<<<1, 128, 162 * 1024, (cudaStream_t)stream>>>(tmA, tmB, *p, D);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
