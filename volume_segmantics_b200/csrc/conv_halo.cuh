// "Halo" tcgen05 convolution: 3x3, stride 1, pad = dilation, one NHWC source with
// Cin % 64 == 0.  Parameter block shared by conv_halo.cu and engine.cu.
//
// An output tile is 8 (x) by 16 (y) pixels = the 128 rows of one UMMA.  Its input
// window with a `dil`-pixel halo ((8+2d) x (16+2d) pixels x 64 channels) is fetched
// ONCE per 64-channel slab by a single TMA box load; all nine filter taps then read
// it in place through shifted shared-memory descriptors:
//     A(tap ky,kx) = halo + ((ky*d)*(8+2d) + kx*d) * 128 bytes,  SBO = (8+2d)*128
// (8 x-adjacent pixels form one 8-row swizzle group, consecutive groups are
// consecutive image rows).  Compared with one box per tap (conv_tc.cu) this moves
// 6.4x fewer rows through TMA -- the measured limiter there (about 4.5 cycles per box
// row) -- and reads each input pixel from L2 1.4x instead of 9x.
// Weights are either resident in shared memory for the whole (persistent) CTA, when
// 9 * Cin * BN * 2 bytes fit, or streamed one (slab, tap) image at a time through
// their own mbarrier ring by a second producer warp.
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace vsb {

constexpr int HALO_KMASK_SLABS = 4;
// Division by a launch constant: q = umulhi(t, m) with m = ceil(2^32 / d), exact while
// t * d < 2^32 (checked by the launcher); m == 0 encodes d == 1.
struct FastDiv {
  uint32_t d, m;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  f.m = d <= 1 ? 0u : (uint32_t)(((1ull << 32) + d - 1) / d);
  return f;
}
struct ConvHaloParams {
  const TmaDesc* map;        // plain 5-D map (c, x, 1, y, n), box {64, 8+2d, 1, 16+2d, 1}
  const uint8_t* wpacked;    // [n_tile][slab][tap] images of [BN][64] pre-swizzled rows
  const float* bias;         // [n_tiles * BN]
  const uint16_t* residual;  // NHWC [NB,H,W,cout] or null
  void* out;
  int32_t out_f32, relu, cout;
  int32_t BN, n_tiles;
  int32_t NB, H, W;
  int32_t n_base;            // first image of this launch
  int32_t cin_off, cout_off; // channel window of this launch (grouped conv = one 64-channel block per launch)
  int32_t kc;                // channels per K-slab: 64 (default when 0) or 32 (Cin % 64 != 0)
  int32_t ncs;               // input channels of the launch / kc
  int32_t dil;
  int32_t tiles_x, tiles_y;
  int32_t a_stages, a_stage_bytes;
  int32_t b_stages;          // 0: all weights resident in shared memory
  int32_t b_bytes;           // BN * 128
  // Resident-weight launches only: bit tap*4+k of kmask[slab] = the K-step (tap, channels
  // 16k..16k+15) has non-zero weights; the others are not issued (space-to-depth convs).
  int32_t use_kmask;
  uint64_t kmask[HALO_KMASK_SLABS];
  // TMEM accumulator stages: 2, 4 or 8 stages of 512 / acc_stages columns (>= BN).  Narrow
  // layers have so little MMA work per tile that two stages leave the MMA warp waiting for
  // the epilogue's round trip; more stages hide it.
  int32_t acc_stages;
  // Shared-memory epilogue (out_map != null): the epilogue warps write the finished tile
  // (bias, residual, ReLU, 16-bit pack) into a swizzled staging buffer and ONE thread hands
  // it to TMA (cp.async.bulk.tensor store), which also clips tiles that overhang the image.
  // The residual tile arrives the same way (res_map, fetched by the A producer warp).
  // Per-thread 16-byte global stores at a pixel stride cost one L1 wavefront per lane --
  // measured as the limiter of every <= 128-channel layer (profiles/r01_ncu_l1conv*).
  const TmaDesc* out_map;   // box {64 ch (16 bit) | cout (f32), 8, 1, 16, 1}
  const TmaDesc* res_map;   // same box over the residual tensor, or null
  int32_t out_bufs, res_bufs;  // staging buffers (1 or 2 / 0, 1 or 2)
  int32_t res_inplace;         // 1: residual tiles land in the output staging buffers (out_bufs == res_bufs == 2), fetched by the store thread
  int32_t out_buf_bytes;       // bytes per staging buffer (multiple of 1024)
  // epi_groups == 2 (BN <= 64, out_bufs == 2): the eight epilogue warps form two groups of
  // four that take alternate tiles, each with its own staging buffer -- the per-tile latency
  // chain (accumulator wait, TMEM load, pack, fences, barriers; about 1 us, measured) is the
  // throughput limit of the small-K layers, and two chains run concurrently.
  int32_t epi_groups;
  int32_t mma_warps;  // 2: two MMA issuing warps on alternate tiles (resident weights, ncs == 1)
  // Streamed-weight launches only: mt == 2 makes a stage two horizontally adjacent 8 x 16 tiles (one
  // (16+2d)-wide halo box, accumulators in two column ranges), so every weight image streamed through the
  // B ring feeds twice the MMAs -- N >= 128 layers are bound by shared-memory traffic (operand reads +
  // incoming weights), not by math.
  int32_t mt;
  FastDiv div_n_tiles, div_tx, div_ty;  // set by the launcher
  // Timing experiments only (results are wrong when set): bit 0 = no halo TMA loads,
  // bit 1 = one MMA per slab, bit 2 = no output stores.
  int32_t dbg;
  // Optional cycle accounting of CTA 0 (vsb_set_flag("halo_prof", 1)): 32 x uint64, see
  // conv_halo.cu HaloProf; null in production.
  unsigned long long* prof;
};

// ---- generalised variant: the halo tile is assembled by four cp.async producer warps
// instead of TMA, so a K-slab may come from any of several concatenated sources and may
// be a nearest-x2 up-sampled view of a half-resolution tensor (smp DecoderBlock:
// upsample -> concat -> conv).  dil = 1.  Every slab of a launch holds KC = 16 / 32 / 64
// channels; pixels are stored with the compact pitch 2*KC bytes in the matching
// SW32 / SW64 / SW128 layout (the swizzle XOR uses absolute address bits, so the tap
// windows may start at any pixel; verified by tests/micro/halo_mma_test.cu).
// One stage covers MT horizontally adjacent 8x16 output tiles: their (8*MT+2) x 18 halo is
// loaded once, MT x 9 x KC/16 MMAs are issued per slab into MT accumulator column ranges,
// and each weight image is reused by MT tiles -- this amortises the mbarrier hand-offs
// that dominate narrow (Cout = 16 / 32) full-resolution layers.
struct HaloSrc {
  const uint16_t* ptr;  // NHWC
  int32_t C, Hs, Ws, up;
};
// Fused head (north-star kernel (c)): when the convolution is the segmentation head, the
// epilogue turns the C logits of a pixel straight into softmax -> first-max label ->
// fp16 p_max -> centre crop -> inverse rotation -> packed-key atomicMax, instead of
// storing logits for a separate kernel (vol_seg_2d_predictor.py:45-64, 90-98).
struct HeadFuse {
  int32_t on, C, d;
  int32_t Hc, Wc, crop_top, crop_left;  // cropped image size and torchvision crop offsets
  int64_t s0;                           // slice index of image 0 of the batch
  int64_t base, stride_s, stride_r, stride_c;
  unsigned long long* keys;
};
// ---- entry-list variant (conv_halo_el_kernel): space-to-depth lowering of the decoder's
// `nearest-x2 upsample -> concat(skip) -> conv3x3` (smp DecoderBlock.conv1).  The 2x2 output
// pixels of a block become 4*Cout GEMM columns at HALF the output resolution; the
// up-sampled source is then read at its own resolution (the 9 taps of the four sub-pixels
// collapse onto a 3x3 low-resolution neighbourhood with summed weights: 16 instead of 36
// (sub-pixel, tap) products), and the skip tensor is read through its "folded" tensor map
// (px*C+c, x/2, py, y/2, n), whose boxes are the four parity planes.  Most (K-slab, tap)
// weight images only feed some of the sub-pixels: every image carries the range of GEMM
// columns it touches, the MMA is issued with that N into that column range, and images
// that touch none are never loaded.  Compared with the parity-split kernels this issues
// the same products for the skip tensor with 2-4x wider N (the narrow-N layers are bound
// by the 128 B/clk shared-memory operand reads: A 4 KB + B N*32 B per MMA), and 2.25x fewer
// products for the up-sampled tensor.
struct HaloSlabRef {
  int32_t map;            // source index (tensor map p.map[map])
  int32_t c, p;           // box coordinates 0 (channel) and 2 (row parity of the folded view; 0 for plain maps)
  int32_t e_begin, e_end; // entries of this slab
};
// Consecutive entries of a slab whose images fit one weight-ring stage together form a GROUP: one bulk copy,
// one full/empty hand-off (the images of a group are consecutive in wpacked).
struct HaloEntry {
  uint32_t ab_off16;  // [15:0] tap offset inside the halo tile, [31:16] image offset inside the ring stage (16-byte units)
  uint32_t w_off;     // byte offset of the image ([n][64] pre-swizzled rows) from wpacked
  uint32_t ncol0_n;   // [11:0] first GEMM column, [27:16] columns (multiples of 16), [31:28] K-steps with non-zero weights
  uint32_t grp;       // [30:0] bytes of the group, on its first entry (else 0); bit 31: last entry of its group
};
constexpr int HALO_EL_MAX_SLABS = 112, HALO_EL_MAX_ENTRIES = 352, HALO_EL_MAX_NTILES = 4;
struct ConvHaloElParams {
  const TmaDesc* map;          // [n_src] halo maps, box {64, 8*mt+2, 1, 18, 1}
  const uint8_t* wpacked;
  const float* bias;           // [n_tiles * BN], index = space-to-depth channel
  void* out;                   // plain 16-bit NHWC [NB, 2H, 2W, cout]
  const TmaDesc* out_map;      // map of `out` (folded when s2d_out), box {64, 8, 1, 16, 1}: epilogue through shared memory + TMA store; null: per-thread stores
  const HaloSlabRef* slabs;
  const HaloEntry* entries;
  int32_t tile_begin[HALO_EL_MAX_NTILES + 1];  // slab refs of N tile t: [tile_begin[t], tile_begin[t+1])
  int32_t n_slabs, n_entries;
  int32_t relu;
  int32_t cout, cout_log2;     // channels of the plain output; space-to-depth channel = (a*2+b)*cout + c
  int32_t s2d_out;             // 1: GEMM columns are space-to-depth channels of an output of size 2H x 2W; 0: plain output H x W
  int32_t BN, n_tiles;
  int32_t NB, H, W;            // space-to-depth grid (half the output size)
  int32_t n_base;
  int32_t mt;                  // 8 x 16 tiles per stage (1 or 2)
  int32_t dbg;                 // timing experiments (results wrong): 1 no halo loads, 2 no weight loads, 4 no stores, 8 one MMA per entry
  int32_t tiles_x, tiles_y;
  int32_t a_stages, a_stage_bytes;
  int32_t b_stages, b_bytes;   // ring of BN * 128-byte stages
  FastDiv div_n_tiles, div_tx, div_ty;  // set by the launcher
};
size_t conv_halo_el_smem_bytes(const ConvHaloElParams& p);
cudaError_t launch_conv_halo_el(const ConvHaloElParams& p, int num_sms, cudaStream_t st);

constexpr int HALO2_MAX_SLABS = 64;
struct ConvHalo2Params {
  HaloSrc src[6];
  const TmaDesc* src_map[6];  // non-null: the source's slabs are fetched by TMA (box {kc, 8*MT+2, 1, 18, 1}); never for up-sampled sources
  int32_t n_src, nslabs;
  int32_t stem;  // 1: 7x7 stride-2 single-channel stem (im2col rows built by the loaders)
  int32_t kc;    // channels per slab (16 / 32 / 64), uniform
  int32_t mt;    // M tiles per stage (1 / 2 / 4)
  int8_t slab_src[HALO2_MAX_SLABS];
  int16_t slab_c0[HALO2_MAX_SLABS];
  const uint8_t* wpacked;           // [n_tile][slab][tap] images of [BN][128 B] rows
  const float* bias;
  const uint16_t* residual;
  void* out;
  int32_t out_f32, relu, cout;
  int32_t BN, n_tiles;
  int32_t NB, H, W;
  int32_t n_base;
  int32_t tiles_x, tiles_y;
  int32_t a_stages, a_stage_bytes;
  int32_t b_stages, b_bytes;
  HeadFuse head;
  // Stem only (MODE 1): fused 3x3 stride-2 pad-1 max-pool (torchvision ResNet.maxpool).  The
  // epilogue stages the ReLU'd tile in shared memory, writes it with coalesced stores and
  // max-reduces the part of every pooling window that lies inside the tile into the
  // zero-initialised pooled tensor [NB, H/2, W/2, cout] with 16-byte red.max (values >= 0).
  uint16_t* pool_out;
  const TmaDesc* out_map;  // stem + pool: the staged output tile is written by TMA (box {64, 8, 1, 16, 1}); null: coalesced stores
  int32_t mma_warps;  // 2: two MMA issuing warps on alternate tiles (MODE 0, resident weights, even a_stages >= 4)
  FastDiv div_n_tiles, div_tx, div_ty;  // set by the launcher
};
// Stem 7x7/2 + fused 3x3/2 max-pool, pool computed inside the CTA (conv_stem.cu).
struct ConvStemParams {
  const uint16_t* in;       // network input [NB, 2H, 2W] 16-bit (tensor 0)
  const uint8_t* wpacked;   // [64][128 B] SW128 rows: K index = ky * 8 + (kx + 1)
  const float* bias;        // [64]
  int32_t version;          // 2: im2col by loader warps (stem_pool_kernel); 3: raw window read in place (stem_pool_v3_kernel)
  // v2: conv output [NB, H, W, 64]: box {64, 14, 1, 16, 1}; pooled [NB, H/2, W/2, 64]: box {64, 7, 1, 8, 1}; SWIZZLE_128B
  // v3: conv output viewed as (c, x / 4, x % 4, y, n): box {64, 8, 1, 14, 1}; pooled: box {64, 15, 1, 7, 1}; SWIZZLE_128B
  const TmaDesc* out_map;
  const TmaDesc* pool_map;
  const TmaDesc* in_map;    // v3: network input as (x, y, n) 16-bit, box {64, 38, 1}, no swizzle, zero fill
  const TmaDesc* out_map7;  // v3: as out_map with box {64, 1, 7, 14, 1} (phase 0 of the leftmost tiles)
  int32_t NB, H, W;         // conv output size
  int32_t n_base;
  int32_t tiles_x, tiles_y; // blocks of 8 x 7 pooled pixels
  int32_t a_stages;         // v2: 2..4 im2col stages of 32 KB; v3: 2..3 window stages of 20 KB
  int32_t dbg;              // v3 bring-up only (results wrong): 1 no MMAs, 2 no TMA loads, 4 no TMA stores, 8 LBO / SBO swapped
  FastDiv div_tx, div_ty;   // set by the launcher
};
size_t conv_stem_smem_bytes(const ConvStemParams& p);
void conv_stem_tiles(int version, int Hc, int Wc, int* tiles_x, int* tiles_y);
cudaError_t conv_stem_configure();
cudaError_t launch_conv_stem(const ConvStemParams& p, int num_sms, cudaStream_t st);

constexpr int HALO2_LOAD_WARPS = 4;
constexpr int HALO2_MMA2_WARP = HALO2_LOAD_WARPS + 2 + 8;  // second MMA issuer (mma_warps == 2)
constexpr int HALO2_THREADS = 32 * (HALO2_LOAD_WARPS + 2 + 8 + 1);

size_t conv_halo2_smem_bytes(const ConvHalo2Params& p);
cudaError_t launch_conv_halo2(const ConvHalo2Params& p, int num_sms, cudaStream_t st);

constexpr int HALO_EPI_WARPS = 8;
constexpr int HALO_MMA2_WARP = 3 + HALO_EPI_WARPS;   // second MMA issuer (mma_warps == 2)
// 12 warps = 384 threads: ptxas sizes the register file for the block rounded up to 128 threads, so a 13th
// warp would cap the kernel at 128 registers (spills in the epilogue).  The TMA-store issuer of the
// shared-memory epilogue is therefore the weight-producer warp (2), idle once resident weights are loaded.
constexpr int HALO_THREADS = 32 * (4 + HALO_EPI_WARPS);
constexpr int HALO_MAX_A_STAGES = 8;
constexpr int HALO_MAX_B_STAGES = 8;

size_t conv_halo_smem_bytes(const ConvHaloParams& p);
cudaError_t launch_conv_halo(const ConvHaloParams& p, int num_sms, cudaStream_t st);
cudaError_t conv_halo_configure();

}  // namespace vsb
