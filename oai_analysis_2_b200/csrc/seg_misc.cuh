// Parameter blocks of the stem / pool / head kernels (seg_misc.cu) shared with their C-ABI wrappers.
#pragma once
#include <cuda_runtime.h>

namespace oai {

struct StemParams {
  const float* vol;  // [VD, VH, VW]
  int VD, VH, VW;
  int td, th, tw;    // tile size (z,y,x)
  int ed, eh, ew;    // effective size = tile - 2*overlap
  int od, oh, ow;    // overlap
  int gh, gw;        // tile grid in y, x
  int tile0, ntiles; // first tile index of this batch, tiles in batch
  int c0;            // output channels (<= 64, multiple of 8)
  const float* w;    // [27][c0] folded weights (tap-major)
  const float* b;    // [c0]
  void* out;         // [ntiles, td, th, tw, c0] 16-bit ([.., 2*c0] = [hi | lo] planes when out_split)
  int fmt;
  int out_split;
};

struct HeadParams {
  const void* act;   // [ntiles, td, th, tw, C] 16-bit (dc1 output)
  int C, ncls;
  const float* w;    // [ncls][C]
  const float* b;    // [ncls]
  float* out;        // [ncls][VD][VH][VW]
  int VD, VH, VW;
  int td, th, tw, ed, eh, ew, od, oh, ow, gh, gw, tile0, ntiles;
  int cz, cy, cx;    // zeroed border shell (assemble's crop_size, already permuted to z,y,x)
  int out_mode;      // 0: sigmoid probability, 1: (p > 0.5) mask, 2: raw logit
  int fmt;
};

int stem_launch(const StemParams& p, cudaStream_t st);
int maxpool2_launch(const void* in, void* out, int N, int D, int H, int W, int C, int in_split, int out_split, int fmt,
                    cudaStream_t st);
int head_launch(const HeadParams& p, cudaStream_t st);

}  // namespace oai
