// tcgen05 / TMEM implicit-GEMM convolution for sm_100a -- parameter block shared
// between the kernel (conv_tc.cu) and the host-side planner (engine.cu).
//
// GEMM view of a convolution on NHWC 16-bit activations:
//   D[M = output pixels, N = Cout] = sum over K-slabs  A_slab[M, KB] * W_slab[N, KB]^T
// A K-slab is (filter tap, source tensor, block of KB = 16/32/64 input channels).
// Its A operand is fetched by ONE 5-D TMA box load per "pixel class" straight
// from the activation tensor: the tap offset is a coordinate shift and the
// zero padding of the convolution is TMA's out-of-bounds zero fill -- there is
// no im2col buffer.  Every activation tensor is described by 5-D tensor maps
//   plain :  (c,       x,     1,  y,     n)
//   folded:  (px*C+c,  x/2,   py, y/2,   n)     (H, W even)
// The folded view makes stride-2 sampling a unit-stride box, which gives
// stride-2 convolutions and -- together with the 4-class "parity split" of
// the M tile -- the nearest-x2 up-sampled sources and the skip concat of the
// smp decoder blocks without materialising either.
//
// The K loop is described by "runs": one run = (tap, source); its slabs step
// through the source's channels.  All box coordinates that do not depend on
// the tile are precomputed on the host, per pixel class.
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace vsb {

struct TcRun {
  int32_t map;        // index into the tensor-map array (= source index)
  int32_t nblk;       // number of KB-channel slabs in this run
  int32_t w_off16;    // offset (16-byte units) of the first slab's [Npad][KB] weight image
  int32_t w_step16;   // distance between consecutive slabs' weight images
  int32_t cls[4][4];  // per pixel class {c, dx, p, dy}: box at (c + blk*KB, X0+dx, p, Y0+dy, N0)
};

struct ConvTcParams {
  const TmaDesc* maps;
  const TcRun* runs;
  int32_t num_runs;
  int32_t num_slabs;          // sum of nblk
  int32_t row_bytes;          // 2 * KB: 32 / 64 / 128, uniform per launch
  const uint8_t* wpacked;
  const float* bias;          // [n_tiles * BN] (zero padded)
  const uint16_t* residual;   // NHWC [NB,H,W,cout] or null
  void* out;                  // 16-bit / f32 NHWC [NB,H,W,cout]
  int32_t out_f32;
  int32_t relu;
  int32_t cout;               // channels actually stored
  int32_t BN, n_tiles;        // N tile and number of N tiles
  int32_t NB, H, W;           // output dims (NB images starting at image n_base)
  int32_t n_base;
  int32_t cin_off, cout_off;  // channel window (grouped conv: one 64-channel block per launch)
  int32_t ncls_log2;          // 0: one pixel class, 2: four parity classes
  int32_t bw_log2, bh_log2, nt_log2;  // box (per class), bw*bh*nt*ncls == 128
  int32_t tiles_x, tiles_y, tiles_n;  // box-grid tile counts
  int32_t num_stages;
  int32_t stage_bytes;        // A (128 rows) + B (BN rows) of row_bytes, 1024-aligned
  int32_t a_bytes;            // 128 * row_bytes
  // Shared-memory epilogue (out_map != null; 16-bit output, one pixel class, cout % 64 == 0, BN % 64 == 0):
  // a finished 64-channel group of the tile (bias, residual, ReLU, pack) is staged in a swizzled 16 KB buffer
  // and written by ONE cp.async.bulk.tensor store; the residual group arrives the same way.  The per-thread
  // 16-byte global accesses of the direct epilogue touch 32 different lines per warp instruction -- for the
  // wide 1x1 convolutions of the bottleneck encoders (K = 64..512, N = 256..2048, HBM-bound) that was the limit.
  const TmaDesc* out_map;     // box {64, bw, 1, bh, nt} over the output tensor
  const TmaDesc* res_map;     // same box over the residual tensor, or null
};

constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_MAX_STAGES = 12;
constexpr int TC_MAX_RUNS = 128;

size_t conv_tc_smem_bytes(const ConvTcParams& p);
size_t conv_tc_epilogue_bytes(bool with_residual);  // staging of the shared-memory epilogue
cudaError_t launch_conv_tc(const ConvTcParams& p, int num_sms, cudaStream_t st);
cudaError_t conv_tc_configure();  // cudaFuncSetAttribute(max dynamic smem)

}  // namespace vsb
