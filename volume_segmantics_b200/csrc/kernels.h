// Internal launcher prototypes shared by the engine and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vsb200.h"

namespace vsb {

// ---- packed key --------------------------------------------
//  [63:48] fp16 bits of p_max (RNE from fp32; order-isomorphic for p >= 0)
//  [47:44] 15 - d        (earliest direction wins ties, as numpy argmax does)
//  [43:36] label
//  [35:32] 0
//  [31: 0] fp32 bits of p_max (diagnostic only; below every tie-break field)
__host__ __device__ inline unsigned long long pack_key(uint16_t h, int d, uint32_t label,
                                                       uint32_t f32bits) {
  return ((unsigned long long)h << 48) | ((unsigned long long)(15 - d) << 44) |
         ((unsigned long long)(label & 0xffu) << 36) | (unsigned long long)f32bits;
}

struct SrcView {
  const void* ptr;  // NHWC bf16
  int C, H, W;      // stored dims (before the optional x2 nearest upsample)
  int up;
};

struct ConvArgs {
  SrcView src[VSB_MAX_SRC];
  int n_src;
  int NB, H, W;  // output dims
  int cin, cout, kh, kw, stride, pad, dil, groups, relu;
  const void* weights;   // bf16 OHWI [cout][kh][kw][cin/groups]
  const float* bias;     // f32 [cout]
  const void* residual;  // bf16 NHWC [NB,H,W,cout] or null
  void* out;             // bf16 or f32 NHWC [NB,H,W,cout]
  int out_f32;
};

struct HeadArgs {
  const float* logits;  // f32 [nb, Hl, Wl, C]   (Hl = Hp/factor); s2d: [nb, Hp/2, Wp/2, 4*C]
  int C, factor;        // factor > 1: bilinear align_corners=True upsampling
  int s2d;              // 1: space-to-depth logits, channel (2a+b)*C + k = class k of pixel (2i+a, 2j+b)
  int nb;
  vsb_direction g;
  int d;
  int64_t s0;
  unsigned long long* keys;  // or null in vote mode
  uint8_t* votes;            // [C][Z*Y*X] or null
  int64_t nvox;
};

// Number of byte values (0..255) whose fused-multiply-add normalisation differs from the reference
// formula in this build's 16-bit format (must be 0), or -1 on a CUDA error.
int slicer_norm_selfcheck(cudaStream_t st);
void launch_slicer(const uint8_t* vol, const vsb_direction& g, int64_t s0, int nb, uint16_t* out,
                   cudaStream_t st);
void launch_stem7x7(const uint16_t* in, int NB, int Hin, int Win, const void* w_bf16,
                    const float* bias, uint16_t* out, int relu, cudaStream_t st);
void launch_maxpool3x3s2(const uint16_t* in, int NB, int Hin, int Win, int C, uint16_t* out,
                         cudaStream_t st);
void launch_conv_simt(const ConvArgs& a, cudaStream_t st);
void launch_dwconv3x3(const ConvArgs& a, cudaStream_t st, bool tiled = true);  // a.weights = [9][C] repacked
size_t gap_scratch_bytes(int NB, int C);  // fp32 partial sums of the two-phase global average pool
void launch_gap(const uint16_t* in, int NB, int H, int W, int C, uint16_t* out, float* scratch, cudaStream_t st);
void launch_upsample(const uint16_t* in, int NB, int Hin, int Win, int C, int Hout, int Wout,
                     int mode, uint16_t* out, cudaStream_t st);
void launch_head(const HeadArgs& a, cudaStream_t st);
bool launch_head_s2d(const HeadArgs& a, cudaStream_t st);  // kernels_head.cu; false: class count not instantiated
void launch_merge_injected(const float* probs, const uint8_t* labels, const vsb_direction& g, int d,
                           unsigned long long* keys, cudaStream_t st);
void launch_unpack(const unsigned long long* keys, int64_t n, uint8_t* labels, uint16_t* probs,
                   cudaStream_t st);
void launch_reduce_unpack(const unsigned long long* const* keys, int n_ranks, int64_t v0, int64_t n, uint8_t* labels,
                          uint16_t* probs, cudaStream_t st);
// kernels_ingest.cu: typed slicer (datasets.py:129-135 for non-uint8 volumes), moments, clip with counts
bool slicer_typed_supported(int dtype);
void launch_slicer_typed(const void* vol, int dtype, const vsb_direction& g, int64_t s0, int nb, uint16_t* out,
                         cudaStream_t st);
int moments_partials();  // doubles written per moments launch (2 per block, fixed grid)
void launch_moments(const void* x, int dtype, int64_t n, int pass, double mean, double* partial, cudaStream_t st);
void launch_clip_count(const void* in, int dtype, int64_t n, double mean, double lower, double upper, uint8_t* out,
                       unsigned long long* counts, cudaStream_t st);
void launch_f32_to_act(const float* in, uint16_t* out, int64_t n, cudaStream_t st);
void launch_to_f32(const void* in, int is_f32, float* out, int64_t n, cudaStream_t st);

}  // namespace vsb
