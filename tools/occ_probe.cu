// Occupancy probe: does the runtime co-schedule two CTAs per SM for kernels that allocate tensor memory (tcgen05.alloc)?
// Prints cudaOccupancyMaxActiveBlocksPerMultiprocessor for twin kernels with / without a TMEM allocation, and measures
// real co-residency: 296 CTAs spin 200 us each and record their SM id and start time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/occ_probe tools/occ_probe.cu && tools/bin/occ_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool TMEM, int COLS>
__global__ void __launch_bounds__(384, 2) probe_kernel(unsigned long long* rec, int spin_us) {
  __shared__ uint32_t tmem_ptr;
  extern __shared__ uint8_t dyn[];
  unsigned long long t0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  uint32_t smid;
  asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
  if (TMEM) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
  }
  unsigned long long t1;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
  unsigned long long now = t1;
  while (now - t1 < (unsigned long long)spin_us * 1000ull) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
  if (threadIdx.x == 0) {
    rec[blockIdx.x * 4 + 0] = smid;
    rec[blockIdx.x * 4 + 1] = t0;
    rec[blockIdx.x * 4 + 2] = t1;
    rec[blockIdx.x * 4 + 3] = now;
    if (dyn[0] == 255 && spin_us < 0) rec[0] = 0;     // keep the dynamic shared memory alive
  }
  __syncthreads();
  if (TMEM && threadIdx.x < 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(COLS) : "memory");
  }
}

template <typename K>
static void run(const char* name, K kernel, size_t smem) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 115712);
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int nb = -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, 384, smem);
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, kernel);
  unsigned long long* rec;
  const int grid = 296;
  cudaMalloc(&rec, grid * 4 * sizeof(unsigned long long));
  cudaMemset(rec, 0, grid * 4 * sizeof(unsigned long long));
  kernel<<<grid, 384, smem>>>(rec, 200);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<unsigned long long> h(grid * 4);
  cudaMemcpy(h.data(), rec, h.size() * 8, cudaMemcpyDeviceToHost);
  unsigned long long tmin = ~0ull, tmax = 0;
  for (int i = 0; i < grid; ++i) { tmin = std::min(tmin, h[i * 4 + 1]); tmax = std::max(tmax, h[i * 4 + 3]); }
  // pairs of CTAs on the same SM whose spin intervals overlap
  int overlapping = 0;
  for (int i = 0; i < grid; ++i)
    for (int j = i + 1; j < grid; ++j)
      if (h[i * 4] == h[j * 4] && h[i * 4 + 2] < h[j * 4 + 3] && h[j * 4 + 2] < h[i * 4 + 3]) ++overlapping;
  printf("%-28s regs %3d dyn smem %6zu: occupancy API %d CTA/SM; 296 CTAs x 200 us took %.0f us; co-resident pairs %d (%s)\n",
         name, a.numRegs, smem, nb, (tmax - tmin) / 1e3, overlapping, cudaGetErrorString(e));
  cudaFree(rec);
}

int main() {
  run("plain", probe_kernel<false, 32>, 1024);
  run("plain 100KB", probe_kernel<false, 32>, 100 * 1024);
  run("tmem 32 cols", probe_kernel<true, 32>, 1024);
  run("tmem 128 cols", probe_kernel<true, 128>, 1024);
  run("tmem 256 cols", probe_kernel<true, 256>, 1024);
  run("tmem 256 cols 100KB", probe_kernel<true, 256>, 100 * 1024);
  run("tmem 512 cols", probe_kernel<true, 512>, 1024);
  return 0;
}
