#pragma once
#include <cstddef>
#include <cstdint>
struct b200seg_conv_desc;
namespace b200seg {
struct ConvPlan {
  int Ho, Wo;
  int KC, cchunks;
  int BN, n_tiles, cout_pad;
  int TH, TW, tiles_h, tiles_w, total_tiles, grid;
  int a_bytes, b_bytes, stage_bytes, nstages;
  int acc_stride, tmem_cols, occ;
  size_t smem_bytes;
};
int conv_plan(const b200seg_conv_desc* d, ConvPlan* pl);
// Shared memory per SM the tensor-core kernels leave unused (b200seg_set_smem_reserve): in SyncBN mode a few one-warp
// waiter CTAs (1 KB of reserved shared memory each) must fit next to ANY convolution CTA, see bn_kernels.cu.
int smem_reserve();
}  // namespace b200seg
