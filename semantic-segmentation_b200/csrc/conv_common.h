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
}  // namespace b200seg
