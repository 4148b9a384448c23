// libvsb200 engine: plan loading, workspace / tensor-map construction, the
// per-direction prediction loop and the C ABI declared in include/vsb200.h.
//
// Host-side mirror of VolSeg2dPredictor (reference
// volume_segmantics/model/operations/vol_seg_2d_predictor.py:16-136); the Python
// shim above this library only folds BatchNorm, lowers the network to the op
// list and forwards calls.  There is no CPU fallback anywhere in this file.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/vsb200.h"
#include "common.cuh"
#include "conv_halo.cuh"
#include "conv_tc.cuh"
#include "kernels.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t err__ = (call);                                                           \
    if (err__ != cudaSuccess)                                                             \
      return fail(VSB_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,      \
                  cudaGetErrorString(err__));                                             \
  } while (0)

using vsb::TcRun;
using vsb::TmaDesc;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int pow2ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}
int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

enum ProfClass { PC_SLICER = 0, PC_CONV_TC, PC_CONV_SIMT, PC_STEM, PC_POOL, PC_HEAD, PC_OTHER, PC_N };

struct TensorBuf {
  int C = 0, ds = 0, dtype = 0;
  int H = 0, W = 0;
  size_t bytes = 0;
  void* ptr = nullptr;
  int first_def = -1, last_use = -1;
};

// Per-conv state that depends only on the plan (not on the spatial size).
struct ConvPlan {
  bool tc = false;
  bool ps = false;        // parity-split (some source is nearest-x2 up-sampled)
  bool s2 = false;        // stride 2
  int BN = 0, n_tiles = 0;
  std::vector<TcRun> runs;         // map index = source index
  int kb = 64;                     // channels per K-slab (uniform per conv)
  int num_slabs = 0;
  TcRun* d_runs = nullptr;
  uint8_t* d_wpacked = nullptr;
  float* d_bias_pad = nullptr;
  // halo kernel (3x3 stride 1, one source, Cin % 64 == 0)
  bool halo_ok = false;   // plan-level eligibility
  bool use_halo = false;  // decided per workspace (tile efficiency)
  uint8_t* d_whalo = nullptr;
  uint8_t* d_wdw = nullptr;   // depthwise CUDA-core kernel: weights [9][C]
  std::vector<int> blk_src, blk_cin_off;  // grouped halo launches: source and channel offset of every 64-channel block
  vsb::ConvHaloParams hparams{};
  bool depthwise = false;     // groups == cin == cout, 3x3 stride 1: vectorised depthwise kernel, weights [9][C]
  bool grouped_s2 = false;    // groups > 1, 3x3 stride 2: per-block launches of the per-tap kernel (folded maps)
  bool grouped_halo = false;  // groups > 1, 3x3 stride 1: one resident-weight halo launch per 64-channel block
  int n_blocks = 0;
  bool stem_tc = false;   // 7x7/2 single-channel stem on tensor cores (halo2 kernel, MODE 1)
  int pool_op = -1;       // stem: index of the MAXPOOL op fused into its epilogue (per workspace), or -1
  bool fused_away = false;  // MAXPOOL: computed by the preceding stem launch
  bool use_stem2 = false;   // stem + pool by the in-CTA pooling kernel (conv_stem.cu), per workspace
  vsb::ConvStemParams sparams{};
  bool halo2_ok = false, use_halo2 = false;  // cp.async-assembled halo (concat / up-sampled / narrow sources)
  vsb::ConvHalo2Params h2params{};
  // space-to-depth lowering of `upsample x2 + concat -> conv3x3` (op.mode == 2, conv_halo.cuh ConvHaloElParams)
  bool el_ok = false, use_el = false;
  int el_kind = 0;  // 1: space-to-depth output (grid = half the output size), 2: plain output (stride-2 lowering)
  uint8_t* d_wel = nullptr;
  float* d_bias_el = nullptr;
  vsb::HaloSlabRef* d_el_slabs = nullptr;
  vsb::HaloEntry* d_el_entries = nullptr;
  TmaDesc* d_el_maps = nullptr;  // [n_src] halo maps (per workspace)
  vsb::ConvHaloElParams elparams{};
  // spatial-size dependent
  TmaDesc* d_maps = nullptr;
  vsb::ConvTcParams params{};
};

}  // namespace

struct vsb_engine {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  int64_t launches = 0;
  EncodeTiledFn encode = nullptr;

  // plan
  std::vector<vsb_tensor_desc> tdesc;
  std::vector<vsb_op> ops;
  std::vector<ConvPlan> conv;  // parallel to ops
  uint8_t* d_weights = nullptr;
  size_t weight_bytes = 0;
  std::vector<uint8_t> h_weights;
  int num_classes = 0;
  bool has_plan = false;

  // volume
  const uint8_t* d_vol = nullptr;
  uint8_t* d_vol_owned = nullptr;
  size_t vol_owned_bytes = 0;
  int vol_dtype = 2;          // dtype code of the resident volume (2 = uint8; see vsb_set_volume_typed)
  void* d_raw = nullptr;      // raw (unclipped) volume of the pre-processing calls (vsb_raw_*)
  int raw_dtype = 0;
  int64_t raw_n = 0;
  int64_t vol_generation = 0;  // bumped whenever the resident volume changes (host-side residency token)
  int64_t Z = 0, Y = 0, X = 0;
  unsigned long long* d_keys = nullptr;
  unsigned long long* d_keys_owned = nullptr;
  uint8_t* d_votes = nullptr;
  size_t votes_bytes = 0;  // capacity of d_votes (voxels x classes of the plan it was allocated for)
  int vote_mode = 0;
  uint8_t* d_labels = nullptr;
  uint16_t* d_probs = nullptr;
  // multi-GPU peers (CUDA IPC mappings of the other ranks' key volumes)
  int n_ranks = 1, my_rank = 0;
  unsigned long long* peer_keys[8] = {nullptr};
  bool peers_ipc = false;  // peer_keys are CUDA-IPC mappings (closed with cudaIpcCloseMemHandle) vs same-process pointers

  // workspace
  int ws_Hp = 0, ws_Wp = 0, ws_nb = 0;
  std::vector<TensorBuf> tens;
  uint8_t* arena = nullptr;   // one allocation holding every activation tensor of the workspace
  size_t arena_bytes = 0;
  bool keep_all = false;
  bool ws_keep = false;
  float* d_gap_scratch = nullptr;  // fp32 partials of the two-phase global average pool
  size_t gap_scratch_bytes = 0;

  int batch_override = 0;
  int64_t row_batch_px = 0;      // vsb_set_flag("row_batch_mpx", n): padded pixels per launch sequence for row directions (0 = 128 Mi)
  int conv_impl = 0;
  bool no_halo = false;
  bool no_tma_epilogue = false;  // vsb_set_flag("tma_epilogue", 0): per-thread global stores in the halo epilogue
  int halo_a_stages_max = 8;     // vsb_set_flag("halo_a_stages", n)
  bool sync_each = false;        // vsb_set_flag("sync_each", 1): synchronise after every op and name the one that failed
  bool no_fuse_pool = false;     // vsb_set_flag("fuse_pool", 0): separate max-pool kernel after the stem
  bool no_tc_smem_epilogue = false;  // vsb_set_flag("tc_smem_epilogue", 0): per-thread global stores in the per-tap kernel
  bool no_dw_tiled = false;      // vsb_set_flag("dw_tiled", 0): one output per thread in the depthwise kernel
  int stem_dbg = 0;              // vsb_set_flag("stem_dbg", bits): bring-up experiments of the v3 stem (results wrong)
  int stem_version = 3;          // vsb_set_flag("stem", v): 3 = raw window read in place (no im2col), 2 = im2col by loader warps,
                                 // 1 = conv_halo2_kernel variant with red.global.max pooling
  int halo_ab_override = 0;      // vsb_set_flag("halo_ab", a*10+b): ring depths of streamed-weight launches (tuning aid)
  int halo_mt_max_bn = 128;      // vsb_set_flag("halo_mt_bn", n): largest BN that gets two tiles per stage
  bool no_halo_mt = false;       // vsb_set_flag("halo_mt", 1): one tile per stage in streamed-weight conv_halo launches
  bool no_halo2_tma = false;     // vsb_set_flag("halo2_tma", 0): cp.async loaders for every halo2 source
  bool halo2_mma2 = false;       // vsb_set_flag("halo2_mma2", 1): two MMA warps in the cp.async halo kernel as well
  bool no_mma2 = false;          // vsb_set_flag("mma_warps", 1): a single MMA issuing warp everywhere
  bool no_s2d_up = false;        // vsb_set_flag("s2d_up", 0): decoder conv1 layers on the parity-split kernels
  bool no_res_inplace = false;   // vsb_set_flag("res_inplace", 0): separate residual staging buffers in the halo kernel
  bool no_dw_tc = false;         // vsb_set_flag("dw_tc", 0): depthwise convolutions on the CUDA-core kernel
  bool no_el_conv = false;       // vsb_set_flag("el_conv", 0): 32 -> 32 and stride-2 3x3 convolutions on their round-1 kernels
  int el_a_stages = 2;           // vsb_set_flag("el_a_stages", n): halo ring depth of the entry-list kernel
  bool no_el_tma_epilogue = false;  // vsb_set_flag("el_tma_epilogue", 0): per-thread stores in the entry-list kernel
  bool no_epi_groups = false;    // vsb_set_flag("epi_groups", 0): one epilogue group even for BN <= 64
  unsigned long long* d_halo_prof = nullptr;  // vsb_set_flag("halo_prof", 1): per-launch cycle accounting to stderr
  int halo_dbg = 0;              // vsb_set_flag("halo_dbg", bits): timing experiments, see ConvHaloParams::dbg
  bool no_fuse_head = true;   // fused head is bit-identical but measured slower (epilogue-bound); opt-in via vsb_set_flag
  int fuse_op = -1;            // conv op whose epilogue performs the head (set per predict_range batch)
  vsb::HeadFuse fuse{};
  int sub_batch_mb = 0;  // L2 budget (MB) per tensor for depth-first sub-batches; 0 = off (measured slower: per-launch prologue dominates)
  bool profiling = false;
  std::vector<float> op_ms;
  std::vector<int64_t> op_launches;
  std::vector<int> ev_op;
  float prof_ms[PC_N] = {0};
  int64_t prof_launches[PC_N] = {0};
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  std::vector<int> ev_cls;
  size_t ev_used = 0;
};

namespace {

// ---------------------------------------------------------------------------
// Direction geometry (host arithmetic).  d = 3k + a.
//   np.rot90(V, k) in the (axis0, axis1) plane   vol_seg_2d_predictor.py:108
//   rotate_array_to_axis                          base_data_utils.py:132-138
// ---------------------------------------------------------------------------
int round_half_even_half(int64_t p) {
  // Python round(p / 2.0): banker's rounding of a half-integer
  if ((p & 1) == 0) return (int)(p / 2);
  const int64_t lo = p / 2;  // floor for p >= 0
  return (int)((lo & 1) ? lo + 1 : lo);
}

int direction_geometry(int64_t Z, int64_t Y, int64_t X, int d, vsb_direction* g) {
  if (d < 0 || d > 11 || Z <= 0 || Y <= 0 || X <= 0) return fail(VSB_ERR_INVALID, "bad direction/shape");
  const int k = d / 3, a = d % 3;
  const int64_t ni = (k & 1) ? Y : Z, nj = (k & 1) ? Z : Y;
  auto vox = [&](int64_t s, int64_t r, int64_t c) -> int64_t {
    int64_t i, j, x;
    if (a == 0) { i = s; j = r; x = c; }
    else if (a == 1) { i = r; j = s; x = c; }
    else { i = c; j = r; x = s; }
    int64_t z, y;
    switch (k) {
      case 0: z = i; y = j; break;
      case 1: z = j; y = Y - 1 - i; break;
      case 2: z = Z - 1 - i; y = Y - 1 - j; break;
      default: z = Z - 1 - j; y = i; break;
    }
    return (z * Y + y) * X + x;
  };
  if (a == 0) { g->S = ni; g->H = nj; g->W = X; }
  else if (a == 1) { g->S = nj; g->H = ni; g->W = X; }
  else { g->S = X; g->H = nj; g->W = ni; }
  g->Hp = (g->H + 31) / 32 * 32;
  g->Wp = (g->W + 31) / 32 * 32;
  const int64_t ph = g->Hp - g->H, pw = g->Wp - g->W;
  g->pad_top = ph / 2;
  g->pad_left = pw / 2;
  g->crop_top = round_half_even_half(ph);
  g->crop_left = round_half_even_half(pw);
  g->base = vox(0, 0, 0);
  g->stride_s = vox(1, 0, 0) - g->base;
  g->stride_r = vox(0, 1, 0) - g->base;
  g->stride_c = vox(0, 0, 1) - g->base;
  return VSB_OK;
}

// ---------------------------------------------------------------------------
// Plan-time preparation of a tcgen05 convolution: eligibility, N tiling, the
// K-slab table and the pre-swizzled weight image.
// ---------------------------------------------------------------------------
bool conv_tc_eligible(const vsb_engine* e, const vsb_op& op) {
  if (op.kind != VSB_OP_CONV) return false;
  if (op.groups != 1) return false;
  if (op.stride != 1 && op.stride != 2) return false;
  bool any_up = false;
  for (int s = 0; s < op.n_src; ++s) {
    const vsb_tensor_desc& t = e->tdesc[op.src[s]];
    if (t.channels % 16 != 0 || t.dtype != 0 || t.ds_log2 < 0) return false;
    any_up |= op.src_up[s] != 0;
  }
  if (any_up && op.stride != 1) return false;
  if (e->tdesc[op.out].ds_log2 < 0) return false;
  if ((op.cout + 15) / 16 * 16 > 2048) return false;
  return true;
}

// ---- entry-list halo kernel (conv_halo.cuh ConvHaloElParams): three lowerings share one table builder ----
// A lowered convolution is a 3x3 pad-1 stride-1 convolution on the "grid" (half the resolution of the folded
// sources) with weights w[N][3][3][K]; `slabs` says where every 64-channel K-slab comes from.
struct ElSlabSpec {
  int map, c, p;  // source index, box channel coordinate (folded view: px*C + c), row parity
  int k0[2];      // K index of channels 0 and 32 of the slab (a folded slab may straddle px when C = 32 mod 64)
};
// Builds slab refs, (slab, tap) entries with their non-zero GEMM column range and K-step mask, and the packed
// weight images.  kind: 1 = space-to-depth output (N = 4*cout, un-shuffled by the epilogue), 2 = plain output.
int build_el_tables(vsb_engine* e, int oi, const uint16_t* w, int N, int K, const std::vector<ElSlabSpec>& slabs,
                    int kind) {
  const vsb_op& op = e->ops[oi];
  ConvPlan& cp = e->conv[oi];
  const int n_tiles = (N + 255) / 256, BN = N / n_tiles;
  if (n_tiles > vsb::HALO_EL_MAX_NTILES || BN * n_tiles != N || BN % 16) return VSB_OK;
  const int mt = 2, HWt = 8 * mt + 2;
  std::vector<vsb::HaloSlabRef> refs;
  std::vector<vsb::HaloEntry> entries;
  std::vector<uint8_t> packed;
  int tile_begin[vsb::HALO_EL_MAX_NTILES + 1] = {0};
  for (int nt = 0; nt < n_tiles; ++nt) {
    tile_begin[nt] = (int)refs.size();
    bool first = true;
    for (const ElSlabSpec& sl : slabs) {
      vsb::HaloSlabRef ref{sl.map, sl.c, sl.p, (int)entries.size(), 0};
      int grp_first = -1, grp_rows = 0;  // open group of this slab
      // the centre tap of the first slab goes first: it must initialise all BN columns of the accumulator
      const int order[9] = {4, 0, 1, 2, 3, 5, 6, 7, 8};
      for (int ti = 0; ti < 9; ++ti) {
        const int tap = order[ti];
        int r0 = BN, r1 = 0;    // non-zero row range of the image
        uint32_t kmask = 0;     // K-steps (16 channels) with any non-zero weight
        for (int n = 0; n < BN; ++n) {
          const uint16_t* wr = w + ((int64_t)(nt * BN + n) * 9 + tap) * K;
          bool nz = false;
          for (int ks = 0; ks < 4; ++ks) {
            const uint16_t* q = wr + sl.k0[ks >> 1] + (ks & 1) * 16;
            bool z = true;
            for (int j = 0; j < 16 && z; ++j) z = (q[j] & 0x7fffu) == 0;
            if (!z) { kmask |= 1u << ks; nz = true; }
          }
          if (nz) { r0 = std::min(r0, n); r1 = std::max(r1, n + 1); }
        }
        if (first) { r0 = 0; r1 = BN; kmask |= 1u; }
        if (r1 <= r0) continue;
        r0 = r0 / 16 * 16;
        r1 = (r1 + 15) / 16 * 16;
        vsb::HaloEntry en{};
        if (grp_first < 0 || grp_rows + (r1 - r0) > BN) {  // close the group, open the next
          if (grp_first >= 0) entries.back().grp |= 0x80000000u;
          grp_first = (int)entries.size();
          grp_rows = 0;
        }
        en.ab_off16 = (uint32_t)(((tap / 3) * HWt + tap % 3) * 8) | ((uint32_t)(grp_rows * 8) << 16);  // 128-byte pixels / rows
        en.w_off = (uint32_t)packed.size();
        en.ncol0_n = (uint32_t)r0 | ((uint32_t)(r1 - r0) << 16) | (kmask << 28);
        en.grp = 0;
        grp_rows += r1 - r0;
        const size_t base = packed.size();
        packed.resize(base + (size_t)(r1 - r0) * 128, 0);
        for (int n = r0; n < r1; ++n)
          for (int ch = 0; ch < 8; ++ch) {  // 16-byte chunks of 8 channels, Swizzle<3,4,3>
            const int h = ch / 4, j = (ch % 4) * 8;
            memcpy(packed.data() + base + (size_t)(n - r0) * 128 + ((ch ^ (n & 7)) * 16),
                   &w[((int64_t)(nt * BN + n) * 9 + tap) * K + sl.k0[h] + j], 16);
          }
        entries.push_back(en);
        entries[grp_first].grp = (entries[grp_first].grp & 0x80000000u) | (uint32_t)(grp_rows * 128);
        first = false;
      }
      if (grp_first >= 0) entries.back().grp |= 0x80000000u;
      ref.e_end = (int)entries.size();
      if (ref.e_end > ref.e_begin) refs.push_back(ref);
    }
    tile_begin[nt + 1] = (int)refs.size();
    if (tile_begin[nt + 1] == tile_begin[nt]) return VSB_OK;
  }
  if ((int)refs.size() > vsb::HALO_EL_MAX_SLABS || (int)entries.size() > vsb::HALO_EL_MAX_ENTRIES) return VSB_OK;
  CK(cudaMalloc(&cp.d_wel, packed.size()));
  CK(cudaMemcpy(cp.d_wel, packed.data(), packed.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&cp.d_el_slabs, refs.size() * sizeof(vsb::HaloSlabRef)));
  CK(cudaMemcpy(cp.d_el_slabs, refs.data(), refs.size() * sizeof(vsb::HaloSlabRef), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&cp.d_el_entries, entries.size() * sizeof(vsb::HaloEntry)));
  CK(cudaMemcpy(cp.d_el_entries, entries.data(), entries.size() * sizeof(vsb::HaloEntry), cudaMemcpyHostToDevice));
  std::vector<float> bias(N, 0.f);
  if (op.b_off >= 0)
    for (int i = 0; i < N; ++i) memcpy(&bias[i], e->h_weights.data() + op.b_off + (size_t)(i % op.cout) * 4, 4);
  CK(cudaMalloc(&cp.d_el_maps, sizeof(TmaDesc) * (VSB_MAX_SRC + 1)));  // sources, output
  CK(cudaMalloc(&cp.d_bias_el, N * 4));
  CK(cudaMemcpy(cp.d_bias_el, bias.data(), N * 4, cudaMemcpyHostToDevice));
  vsb::ConvHaloElParams& h = cp.elparams;
  h = vsb::ConvHaloElParams{};
  h.wpacked = cp.d_wel;
  h.bias = cp.d_bias_el;
  h.slabs = cp.d_el_slabs;
  h.entries = cp.d_el_entries;
  for (int i = 0; i <= n_tiles; ++i) h.tile_begin[i] = tile_begin[i];
  for (int i = n_tiles + 1; i <= vsb::HALO_EL_MAX_NTILES; ++i) h.tile_begin[i] = tile_begin[n_tiles];
  h.n_slabs = (int)refs.size();
  h.n_entries = (int)entries.size();
  h.relu = op.relu;
  h.cout = op.cout;
  h.cout_log2 = ilog2(op.cout);
  h.s2d_out = kind == 1;
  h.BN = BN;
  h.n_tiles = n_tiles;
  h.mt = mt;
  cp.el_kind = kind;
  cp.el_ok = true;
  return VSB_OK;
}

// K-slabs of the space-to-depth ("folded") view of source s: (py, 64 channels of px*C + c).
static void el_folded_slabs(std::vector<ElSlabSpec>& slabs, int s, int C, int kbase, int Ctotal, int coff) {
  for (int py = 0; py < 2; ++py)
    for (int d0 = 0; d0 < 2 * C; d0 += 64) {
      auto kidx = [&](int d) { const int px = d / C, c = d % C; return kbase + (py * 2 + px) * Ctotal + coff + c; };
      slabs.push_back({s, d0, py, {kidx(d0), kidx(d0 + 32)}});
    }
}

// Lowering 1 (op.mode == 2, written by plan.py s2d_up_concat_weights): decoder `upsample x2 + concat -> conv3x3`
// as a convolution at half the output resolution.  The blob holds, at offset op.factor * 256, the weights
// [4*cout][3][3][Cup + 4*Cskip]: K = channels of the up-sampled source at its own resolution, then the
// (py, px, c) sub-pixels of the concatenated skip sources; output channel (a*2+b)*cout + c = pixel (2i+a, 2j+b).
// Lowering 2 (built here, exact): a plain 3x3 stride-1 convolution with 32 output channels as a convolution
// between the space-to-depth views of its input and output (N = 128 instead of 32).
// Lowering 3 (built here, exact): a 3x3 stride-2 pad-1 convolution as a stride-1 convolution over the
// space-to-depth view of its input -- the nine taps become nine (parity plane, offset) pairs of FOUR halo boxes.
int prepare_el_plan(vsb_engine* e, int oi) {
  const vsb_op& op = e->ops[oi];
  ConvPlan& cp = e->conv[oi];
  cp.el_ok = false;
  if (op.kind != VSB_OP_CONV || op.kh != 3 || op.kw != 3 || op.pad != 1 || op.dil != 1 || op.groups != 1 || op.n_src < 1)
    return VSB_OK;
  const vsb_tensor_desc& ot = e->tdesc[op.out];
  if (ot.dtype != 0 || ot.ds_log2 < 0 || op.res >= 0) return VSB_OK;
  const int cout = op.cout;
  std::vector<ElSlabSpec> slabs;
  if (op.mode == 2) {
    if (op.stride != 1 || op.n_src < 2 || !op.src_up[0])
      return fail(VSB_ERR_INVALID, "op %d: mode 2 needs conv3x3 pad 1 over [up-sampled source, skip sources...]", oi);
    if (cout != 32 && cout != 64 && cout != 128 && cout != 256) return VSB_OK;
    const int Cup = e->tdesc[op.src[0]].channels;
    if (Cup % 64 || e->tdesc[op.src[0]].dtype != 0 || e->tdesc[op.src[0]].ds_log2 != ot.ds_log2 + 1) return VSB_OK;
    int Cskip = 0;
    for (int s = 1; s < op.n_src; ++s) {
      const vsb_tensor_desc& t = e->tdesc[op.src[s]];
      if (op.src_up[s] || t.channels % 32 || t.dtype != 0 || t.ds_log2 != ot.ds_log2) return VSB_OK;
      Cskip += t.channels;
    }
    if (Cup + Cskip != op.cin) return fail(VSB_ERR_INVALID, "op %d: cin %d != sum of sources", oi, op.cin);
    const int K = Cup + 4 * Cskip, N = 4 * cout;
    const int64_t woff = (int64_t)op.factor * 256;
    if (op.factor <= 0 || (size_t)woff + (size_t)N * 9 * K * 2 > e->weight_bytes)
      return fail(VSB_ERR_INVALID, "op %d: space-to-depth weight range", oi);
    for (int c0 = 0; c0 < Cup; c0 += 64) slabs.push_back({0, c0, 0, {c0, c0 + 32}});
    for (int s = 1, coff = 0; s < op.n_src; ++s) {
      el_folded_slabs(slabs, s, e->tdesc[op.src[s]].channels, Cup, Cskip, coff);
      coff += e->tdesc[op.src[s]].channels;
    }
    return build_el_tables(e, oi, reinterpret_cast<const uint16_t*>(e->h_weights.data() + woff), N, K, slabs, 1);
  }
  if (op.mode != 0) return VSB_OK;
  int Cin = 0;
  for (int s = 0; s < op.n_src; ++s) {
    const vsb_tensor_desc& t = e->tdesc[op.src[s]];
    if (op.src_up[s] || t.channels % 32 || t.dtype != 0 || t.ds_log2 < 0) return VSB_OK;
    Cin += t.channels;
  }
  if (Cin != op.cin) return VSB_OK;
  const uint16_t* w = reinterpret_cast<const uint16_t*>(e->h_weights.data() + op.w_off);  // [cout][3][3][Cin]
  auto floordiv2 = [](int v) { return v >= 0 ? v / 2 : -((1 - v) / 2); };
  if (op.stride == 1 && cout == 32 && Cin == 32 && op.n_src == 1) {
    // lowering 2: W2[(a*2+b)*O + o][dy+1][dx+1][(a2*2+b2)*I + i] = w[o][ky][kx][i], a + ky - 1 = 2*dy + a2
    const int K = 4 * Cin, N = 4 * cout;
    std::vector<uint16_t> w2((size_t)N * 9 * K, 0);
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx) {
            const int r = a + ky - 1, c = b + kx - 1;
            const int dy = floordiv2(r), a2 = r - 2 * dy, dx = floordiv2(c), b2 = c - 2 * dx;
            for (int o = 0; o < cout; ++o)
              memcpy(&w2[(((size_t)((a * 2 + b) * cout + o) * 3 + dy + 1) * 3 + dx + 1) * K + (a2 * 2 + b2) * Cin],
                     &w[(((size_t)o * 3 + ky) * 3 + kx) * Cin], (size_t)Cin * 2);
          }
    el_folded_slabs(slabs, 0, Cin, 0, Cin, 0);
    return build_el_tables(e, oi, w2.data(), N, K, slabs, 1);
  }
  if (op.stride == 2 && (cout == 64 || cout == 128 || cout == 256 || cout == 512) && Cin % 64 == 0 && op.n_src == 1 &&
      e->tdesc[op.src[0]].ds_log2 + 1 == ot.ds_log2) {
    // lowering 3: W2[o][ty+1][tx+1][(qy*2+qx)*I + i] = w[o][ky][kx][i], ky - 1 = 2*ty + qy
    const int K = 4 * Cin, N = cout;
    std::vector<uint16_t> w2((size_t)N * 9 * K, 0);
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int ty = floordiv2(ky - 1), qy = ky - 1 - 2 * ty, tx = floordiv2(kx - 1), qx = kx - 1 - 2 * tx;
        for (int o = 0; o < cout; ++o)
          memcpy(&w2[(((size_t)o * 3 + ty + 1) * 3 + tx + 1) * K + (qy * 2 + qx) * Cin], &w[(((size_t)o * 3 + ky) * 3 + kx) * Cin],
                 (size_t)Cin * 2);
      }
    el_folded_slabs(slabs, 0, Cin, 0, Cin, 0);
    return build_el_tables(e, oi, w2.data(), N, K, slabs, 2);
  }
  return VSB_OK;
}

int prepare_conv_plan(vsb_engine* e, int oi) {
  const vsb_op& op = e->ops[oi];
  ConvPlan& cp = e->conv[oi];
  cp.tc = conv_tc_eligible(e, op);
  cp.stem_tc = false;
  if (op.kind == VSB_OP_CONV && op.cin == 1 && op.kh == 7 && op.kw == 7 && op.stride == 2 && op.pad == 3 &&
      op.dil == 1 && op.groups == 1 && op.n_src == 1 && op.res < 0 && op.cout % 16 == 0 && op.cout <= 256 &&
      e->tdesc[op.out].dtype == 0) {
    cp.BN = op.cout;
    cp.n_tiles = 1;
    std::vector<uint8_t> img((size_t)cp.BN * 128, 0);
    // K index = ky * 8 + j with j = kx + 1 (j = 0 is input column 2*ox-4, outside the window: weight 0)
    for (int n = 0; n < op.cout; ++n)
      for (int ky = 0; ky < 7; ++ky)
        for (int kx = 0; kx < 7; ++kx)
          memcpy(img.data() + (size_t)n * 128 + ((ky ^ (n & 7)) * 16) + (kx + 1) * 2,
                 e->h_weights.data() + op.w_off + ((int64_t)n * 49 + ky * 7 + kx) * 2, 2);
    CK(cudaMalloc(&cp.d_whalo, img.size()));
    CK(cudaMemcpy(cp.d_whalo, img.data(), img.size(), cudaMemcpyHostToDevice));
    std::vector<float> bias(cp.BN, 0.f);
    if (op.b_off >= 0) memcpy(bias.data(), e->h_weights.data() + op.b_off, (size_t)op.cout * 4);
    CK(cudaMalloc(&cp.d_bias_pad, cp.BN * 4));
    CK(cudaMemcpy(cp.d_bias_pad, bias.data(), cp.BN * 4, cudaMemcpyHostToDevice));
    cp.stem_tc = true;
    return VSB_OK;
  }
  cp.depthwise = false;
  if (op.kind == VSB_OP_CONV && op.groups == op.cin && op.cin == op.cout && op.kh == 3 && op.kw == 3 &&
      op.stride == 1 && op.pad == op.dil && op.res < 0 && e->tdesc[op.out].dtype == 0) {
    bool ok = true;
    for (int s = 0; s < op.n_src; ++s)
      ok &= e->tdesc[op.src[s]].channels % 8 == 0 && !op.src_up[s] && e->tdesc[op.src[s]].dtype == 0;
    if (ok) {
      std::vector<uint16_t> w((size_t)9 * op.cout);
      for (int c = 0; c < op.cout; ++c)
        for (int tap = 0; tap < 9; ++tap)
          memcpy(&w[(size_t)tap * op.cout + c], e->h_weights.data() + op.w_off + ((int64_t)c * 9 + tap) * 2, 2);
      CK(cudaMalloc(&cp.d_wdw, w.size() * 2));
      CK(cudaMemcpy(cp.d_wdw, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
      cp.depthwise = true;
      // Round 2: a depthwise convolution is the extreme grouped convolution (one channel per group) and runs on
      // the tensor-core grouped paths below with block-diagonal (here: diagonal) 64 x 64 weight images -- 63/64 of
      // the products are zeros, and it is still 2.5-4x faster than the CUDA-core kernel, which is bound by its
      // 16-bit <-> fp32 conversions (1.1 TB/s).  The CUDA-core kernel stays as the small-image / cross-check path.
    }
  }
  cp.grouped_halo = false;
  // 64-channel block b reads channels [blk_cin_off[b], +64) of source blk_src[b]: a depthwise convolution over a
  // channel concat (DeepLabV3+ decoder: 256 up-sampled ASPP channels + 48 high-resolution channels) never
  // materialises the concat as long as every source but the last holds a multiple of 64 channels
  cp.blk_src.clear();
  cp.blk_cin_off.clear();
  // (at most three sources: slots VSB_MAX_SRC - 3 .. - 1 of the map array hold the epilogue maps of the halo launches)
  bool blocks_ok = op.kind == VSB_OP_CONV && op.n_src >= 1 && op.n_src <= 3 && (op.n_src == 1 || cp.depthwise);
  for (int s = 0; s < op.n_src && blocks_ok; ++s) {
    const vsb_tensor_desc& t = e->tdesc[op.src[s]];
    blocks_ok = !op.src_up[s] && t.dtype == 0 && t.channels % 8 == 0 && (s == op.n_src - 1 || t.channels % 64 == 0);
    for (int c0 = 0; c0 < t.channels; c0 += 64) {
      cp.blk_src.push_back(s);
      cp.blk_cin_off.push_back(c0);
    }
  }
  if (blocks_ok && op.groups > 1 && op.kh == 3 && op.kw == 3 &&
      op.stride == 1 && op.pad == op.dil && (op.dil == 1 || op.dil == 2) && op.cin == op.cout && op.cin % 8 == 0 &&
      (op.cin % 64 == 0 || op.res < 0) && 64 % (op.cin / op.groups) == 0 &&
      e->tdesc[op.out].dtype == 0 && e->tdesc[op.out].ds_log2 >= 0) {
    // Block-diagonal packing: the 64 output channels of block b only see the 64 input
    // channels of block b (groups never straddle a block); cross-group weights are zero.
    // A last block of fewer than 64 channels reads / writes past the channel count: TMA zero-fills the box,
    // the epilogue skips channels >= cout.
    const int cg = op.cin / op.groups;
    cp.n_blocks = (op.cin + 63) / 64;
    cp.BN = 64;
    cp.n_tiles = 1;
    const size_t img = 64 * 128;
    std::vector<uint8_t> hp((size_t)cp.n_blocks * 9 * img, 0);
    for (int b = 0; b < cp.n_blocks; ++b)
      for (int tap = 0; tap < 9; ++tap) {
        uint8_t* dst = hp.data() + ((size_t)b * 9 + tap) * img;
        for (int n = 0; n < 64; ++n) {
          const int o = b * 64 + n;
          if (o >= op.cout) break;
          const int g0 = (o / cg) * cg - b * 64;  // first input channel (block-local) of o's group
          for (int j = 0; j < cg; ++j) {
            const int k = g0 + j;  // block-local input channel
            uint32_t off = (uint32_t)(n * 128 + k * 2);
            off ^= ((off >> 7) & 7u) << 4;
            memcpy(dst + off, e->h_weights.data() + op.w_off + ((((int64_t)o * 3 + tap / 3) * 3 + tap % 3) * cg + j) * 2, 2);
          }
        }
      }
    CK(cudaMalloc(&cp.d_whalo, hp.size()));
    CK(cudaMemcpy(cp.d_whalo, hp.data(), hp.size(), cudaMemcpyHostToDevice));
    std::vector<float> bias((size_t)cp.n_blocks * 64, 0.f);
    if (op.b_off >= 0) memcpy(bias.data(), e->h_weights.data() + op.b_off, (size_t)op.cout * 4);
    CK(cudaMalloc(&cp.d_bias_pad, bias.size() * 4));
    CK(cudaMemcpy(cp.d_bias_pad, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&cp.d_maps, sizeof(TmaDesc) * 2 * VSB_MAX_SRC));  // [0,6): per-tap / halo / epilogue maps, [6,12): halo2 source maps
    cp.grouped_halo = true;
    return VSB_OK;
  }
  cp.grouped_s2 = false;
  // stride 2 (ResNeXt).  (Stride 1 with the ASPP dilations 12 / 24 / 36 also runs here -- set VSB_DW_TC_DILATED -- but one
  // box per tap reads the input nine times from L2 and the 2048-channel depthwise layers need 32 launches each:
  // 4.5 ms against 3.5-3.8 ms of the CUDA-core kernel per 32 slices of 2048^2, so it is off.)
  if (op.kind == VSB_OP_CONV && op.groups > 1 && op.n_src == 1 && !op.src_up[0] && op.kh == 3 && op.kw == 3 &&
      ((op.stride == 2 && op.pad == 1 && op.dil == 1) ||
       (op.stride == 1 && op.pad == op.dil && op.dil > 2 && getenv("VSB_DW_TC_DILATED") != nullptr)) &&
      op.cin == op.cout && op.cin % 64 == 0 &&
      64 % (op.cin / op.groups) == 0 && e->tdesc[op.src[0]].dtype == 0 && e->tdesc[op.out].dtype == 0 &&
      e->tdesc[op.out].ds_log2 >= 0) {
    const int cg = op.cin / op.groups, C = op.cin;
    cp.n_blocks = C / 64;
    cp.BN = 64;
    cp.n_tiles = 1;
    cp.kb = 64;
    cp.s2 = op.stride == 2;
    cp.ps = false;
    cp.runs.clear();
    cp.num_slabs = 9;
    const size_t img = 64 * 128;
    for (int tap = 0; tap < 9; ++tap) {
      TcRun run{};
      run.map = 0;
      run.nblk = 1;
      run.w_off16 = (int32_t)(tap * img / 16);
      run.w_step16 = 0;
      const int ty = tap / 3 - 1, tx = tap % 3 - 1;
      if (cp.s2) {
        run.cls[0][0] = (tx & 1) * C;  // folded view: c' = parity_x * C + c
        run.cls[0][1] = tx >> 1;
        run.cls[0][2] = ty & 1;
        run.cls[0][3] = ty >> 1;
      } else {
        run.cls[0][0] = 0;
        run.cls[0][1] = tx * op.dil;
        run.cls[0][2] = 0;
        run.cls[0][3] = ty * op.dil;
      }
      cp.runs.push_back(run);
    }
    std::vector<uint8_t> hp((size_t)cp.n_blocks * 9 * img, 0);
    for (int b = 0; b < cp.n_blocks; ++b)
      for (int tap = 0; tap < 9; ++tap) {
        uint8_t* dst = hp.data() + ((size_t)b * 9 + tap) * img;
        for (int n = 0; n < 64; ++n) {
          const int o = b * 64 + n;
          const int g0 = (o / cg) * cg - b * 64;
          for (int j = 0; j < cg; ++j) {
            uint32_t off = (uint32_t)(n * 128 + (g0 + j) * 2);
            off ^= ((off >> 7) & 7u) << 4;
            memcpy(dst + off, e->h_weights.data() + op.w_off + ((((int64_t)o * 3 + tap / 3) * 3 + tap % 3) * cg + j) * 2, 2);
          }
        }
      }
    CK(cudaMalloc(&cp.d_wpacked, hp.size()));
    CK(cudaMemcpy(cp.d_wpacked, hp.data(), hp.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&cp.d_runs, cp.runs.size() * sizeof(TcRun)));
    CK(cudaMemcpy(cp.d_runs, cp.runs.data(), cp.runs.size() * sizeof(TcRun), cudaMemcpyHostToDevice));
    std::vector<float> bias(op.cout, 0.f);
    if (op.b_off >= 0) memcpy(bias.data(), e->h_weights.data() + op.b_off, (size_t)op.cout * 4);
    CK(cudaMalloc(&cp.d_bias_pad, op.cout * 4));
    CK(cudaMemcpy(cp.d_bias_pad, bias.data(), op.cout * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&cp.d_maps, sizeof(TmaDesc) * 2 * VSB_MAX_SRC));  // [0,6): per-tap / halo / epilogue maps, [6,12): halo2 source maps
    cp.grouped_s2 = true;
    cp.tc = true;  // shares the workspace-time tile / tensor-map set-up of the per-tap kernel
    return VSB_OK;
  }
  if (!cp.tc || cp.depthwise) return VSB_OK;
  cp.ps = false;
  for (int s = 0; s < op.n_src; ++s) cp.ps |= op.src_up[s] != 0;
  cp.s2 = op.stride == 2;
  const int n_pad16 = (op.cout + 15) / 16 * 16;
  cp.n_tiles = (n_pad16 + 255) / 256;
  cp.BN = ((n_pad16 + cp.n_tiles - 1) / cp.n_tiles + 15) / 16 * 16;
  const int n_total = cp.BN * cp.n_tiles;

  // uniform slab width: the largest of 64/32/16 channels dividing every source
  cp.kb = 64;
  std::vector<int> src_c0;  // channel offset of each source in the concat
  int coff = 0;
  for (int s = 0; s < op.n_src; ++s) {
    const int C = e->tdesc[op.src[s]].channels;
    // A source of e.g. 304 channels (DeepLabV3+ decoder) is read in 64-channel slabs whose last one runs past the
    // channel count (TMA zero-fills the box, the weight image is zero there): 5 slabs instead of 19 of 16 channels.
    int kb = C % 64 == 0 ? 64 : (C % 32 == 0 ? 32 : 16);
    if (kb < 64 && C > 128 && C % 8 == 0) kb = 64;
    cp.kb = std::min(cp.kb, kb);
    src_c0.push_back(coff);
    coff += C;
  }
  if (coff != op.cin) return fail(VSB_ERR_INVALID, "op %d: cin %d != sum of sources %d", oi, op.cin, coff);
  const int KB = cp.kb, rb = KB * 2;
  const int ncls = cp.ps ? 4 : 1;

  // run table (tap x source) + packed weights: per slab a [n_total][KB] image,
  // rows pre-swizzled exactly as TMA would have written them
  cp.runs.clear();
  cp.num_slabs = 0;
  size_t wbytes = 0;
  const size_t slab_img = (size_t)n_total * rb;
  for (int ky = 0; ky < op.kh; ++ky)
    for (int kx = 0; kx < op.kw; ++kx)
      for (int s = 0; s < op.n_src; ++s) {
        const int C = e->tdesc[op.src[s]].channels;
        TcRun run{};
        run.map = s;
        run.nblk = (C + KB - 1) / KB;
        run.w_off16 = (int32_t)(wbytes / 16);
        run.w_step16 = (int32_t)(slab_img / 16);
        const int dy = ky * op.dil - op.pad, dx = kx * op.dil - op.pad;
        bool halve, folded;
        if (cp.ps) { halve = true; folded = !op.src_up[s]; }
        else if (cp.s2) { halve = true; folded = true; }
        else { halve = false; folded = false; }
        for (int q = 0; q < ncls; ++q) {
          const int ty = (q >> 1) + dy, tx = (q & 1) + dx;
          int cy, cx, pary = 0, parx = 0;
          if (halve) { cy = ty >> 1; cx = tx >> 1; pary = ty & 1; parx = tx & 1; }
          else { cy = ty; cx = tx; }
          run.cls[q][0] = folded ? parx * C : 0;
          run.cls[q][1] = cx;
          run.cls[q][2] = folded ? pary : 0;
          run.cls[q][3] = cy;
        }
        wbytes += slab_img * run.nblk;
        cp.num_slabs += run.nblk;
        cp.runs.push_back(run);
      }
  if ((int)cp.runs.size() > vsb::TC_MAX_RUNS) { cp.tc = false; return VSB_OK; }
  std::vector<uint8_t> packed(wbytes, 0);
  const int cin_g = op.cin;  // groups == 1
  size_t ri = 0;
  for (int ky = 0; ky < op.kh; ++ky)
    for (int kx = 0; kx < op.kw; ++kx)
      for (int s = 0; s < op.n_src; ++s, ++ri) {
        const TcRun& run = cp.runs[ri];
        for (int cb = 0; cb < run.nblk; ++cb) {
          uint8_t* img = packed.data() + (size_t)run.w_off16 * 16 + (size_t)cb * slab_img;
          for (int n = 0; n < op.cout; ++n) {
            const int64_t wrow = (((int64_t)n * op.kh + ky) * op.kw + kx) * cin_g + src_c0[s] + cb * KB;
            const int Cs = e->tdesc[op.src[s]].channels;
            for (int ch = 0; ch < rb / 16; ++ch) {  // 16-byte chunks of 8 channels
              if (cb * KB + ch * 8 >= Cs) break;  // past the source's channels (partial last slab): zeros
              int sw;  // Swizzle<B,4,3> on the byte address within the slab image
              if (rb == 128) sw = ch ^ (n & 7);
              else if (rb == 64) sw = ch ^ ((n >> 1) & 3);
              else sw = ch ^ ((n >> 2) & 1);
              memcpy(img + (size_t)n * rb + sw * 16, e->h_weights.data() + op.w_off + (wrow + ch * 8) * 2, 16);
            }
          }
        }
      }
  // 32-channel-slab variant only for BN >= 64: 32 -> 32 layers run faster on the MT-tiled halo2 kernel
  cp.halo_ok = op.n_src == 1 && !cp.ps && !cp.s2 && op.kh == 3 && op.kw == 3 && op.pad == op.dil &&
               (op.dil == 1 || op.dil == 2) &&
               (e->tdesc[op.src[0]].channels % 64 == 0 || (e->tdesc[op.src[0]].channels % 32 == 0 && cp.BN >= 64));
  if (cp.halo_ok) {
    // K-slabs of 64 channels (128-byte rows, SW128) or, when Cin is only a multiple of 32,
    // of 32 channels (64-byte rows, SW64)
    const int kc = op.cin % 64 == 0 ? 64 : 32, P = 2 * kc, ksteps = kc / 16;
    const uint32_t swz_mask = P == 128 ? 7u : 3u;
    const int ncs = op.cin / kc;
    const size_t img = (size_t)cp.BN * P;
    std::vector<uint8_t> hp((size_t)cp.n_tiles * ncs * 9 * img, 0);
    auto woff = [&](int n, int ch) {  // byte offset of 16-byte chunk `ch` of row n in a swizzled image
      uint32_t off = (uint32_t)(n * P + ch * 16);
      return off ^ (((off >> 7) & swz_mask) << 4);
    };
    for (int nt = 0; nt < cp.n_tiles; ++nt)
      for (int cs = 0; cs < ncs; ++cs)
        for (int tap = 0; tap < 9; ++tap) {
          uint8_t* dst = hp.data() + (((size_t)nt * ncs + cs) * 9 + tap) * img;
          for (int n = 0; n < cp.BN; ++n) {
            const int o = nt * cp.BN + n;
            if (o >= op.cout) break;
            const int64_t wrow = (((int64_t)o * 3 + tap / 3) * 3 + tap % 3) * op.cin + cs * kc;
            for (int ch = 0; ch < kc / 8; ++ch)
              memcpy(dst + woff(n, ch), e->h_weights.data() + op.w_off + (wrow + ch * 8) * 2, 16);
          }
        }
    CK(cudaMalloc(&cp.d_whalo, hp.size()));
    CK(cudaMemcpy(cp.d_whalo, hp.data(), hp.size(), cudaMemcpyHostToDevice));
    cp.hparams.kc = kc;
    // K-steps (tap, 16-channel block) whose weights are all zero need no MMA (structured
    // sparsity of the space-to-depth convolutions, plan.py): bit tap*ksteps+k of kmask[slab].
    cp.hparams.use_kmask = 0;
    if (cp.n_tiles == 1 && ncs <= vsb::HALO_KMASK_SLABS) {
      bool any_zero = false;
      for (int cs = 0; cs < ncs; ++cs) {
        uint64_t m = 0;
        for (int tap = 0; tap < 9; ++tap)
          for (int k = 0; k < ksteps; ++k) {
            bool nz = false;
            const uint8_t* im = hp.data() + ((size_t)cs * 9 + tap) * img;
            for (int n = 0; n < cp.BN && !nz; ++n)
              for (int ch = 2 * k; ch < 2 * k + 2 && !nz; ++ch) {
                const uint64_t* q = reinterpret_cast<const uint64_t*>(im + woff(n, ch));
                nz = (q[0] | q[1]) != 0;
              }
            if (nz) m |= 1ull << (tap * ksteps + k);
            else any_zero = true;
          }
        cp.hparams.kmask[cs] = m;
      }
      cp.hparams.kmask[0] |= 1ull;  // the first MMA of a tile initialises the accumulator
      cp.hparams.use_kmask = any_zero ? 1 : 0;
    }
  }
  cp.halo2_ok = false;
  if (!cp.halo_ok && !cp.s2 && op.kh == 3 && op.kw == 3 && op.pad == 1 && op.dil == 1) {
    vsb::ConvHalo2Params& h = cp.h2params;
    h = vsb::ConvHalo2Params{};
    int kc = 64;
    for (int s = 0; s < op.n_src; ++s) {
      const int C = e->tdesc[op.src[s]].channels;
      kc = std::min(kc, C % 64 == 0 ? 64 : (C % 32 == 0 ? 32 : 16));
    }
    int ns = 0;
    bool fits = true;
    for (int s = 0; s < op.n_src && fits; ++s) {
      const int C = e->tdesc[op.src[s]].channels;
      for (int c0 = 0; c0 < C; c0 += kc) {
        if (ns >= vsb::HALO2_MAX_SLABS) { fits = false; break; }
        h.slab_src[ns] = (int8_t)s;
        h.slab_c0[ns] = (int16_t)c0;
        ++ns;
      }
    }
    if (fits) {
      h.nslabs = ns;
      h.n_src = op.n_src;
      h.kc = kc;
      const int P = 2 * kc;
      const uint32_t mask = P == 128 ? 7u : (P == 64 ? 3u : 1u);
      const size_t img = align_up((size_t)cp.BN * P, 1024);
      h.b_bytes = (int32_t)img;
      std::vector<uint8_t> hp((size_t)cp.n_tiles * ns * 9 * img, 0);
      for (int nt = 0; nt < cp.n_tiles; ++nt)
        for (int sl = 0; sl < ns; ++sl)
          for (int tap = 0; tap < 9; ++tap) {
            uint8_t* dst = hp.data() + (((size_t)nt * ns + sl) * 9 + tap) * img;
            const int cbase = src_c0[h.slab_src[sl]] + h.slab_c0[sl];
            for (int n = 0; n < cp.BN; ++n) {
              const int o = nt * cp.BN + n;
              if (o >= op.cout) break;
              const int64_t wrow = (((int64_t)o * 3 + tap / 3) * 3 + tap % 3) * op.cin + cbase;
              for (int ch = 0; ch < kc / 8; ++ch) {
                uint32_t off = (uint32_t)(n * P + ch * 16);
                off ^= ((off >> 7) & mask) << 4;  // Swizzle<B,4,3> relative to the 1024-aligned image
                memcpy(dst + off, e->h_weights.data() + op.w_off + (wrow + ch * 8) * 2, 16);
              }
            }
          }
      CK(cudaMalloc(&cp.d_whalo, hp.size()));
      CK(cudaMemcpy(cp.d_whalo, hp.data(), hp.size(), cudaMemcpyHostToDevice));
      cp.halo2_ok = true;
    }
  }
  CK(cudaMalloc(&cp.d_wpacked, wbytes));
  CK(cudaMemcpy(cp.d_wpacked, packed.data(), wbytes, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&cp.d_runs, cp.runs.size() * sizeof(TcRun)));
  CK(cudaMemcpy(cp.d_runs, cp.runs.data(), cp.runs.size() * sizeof(TcRun), cudaMemcpyHostToDevice));
  std::vector<float> bias(n_total, 0.f);
  if (op.b_off >= 0) memcpy(bias.data(), e->h_weights.data() + op.b_off, (size_t)op.cout * 4);
  CK(cudaMalloc(&cp.d_bias_pad, n_total * 4));
  CK(cudaMemcpy(cp.d_bias_pad, bias.data(), n_total * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&cp.d_maps, sizeof(TmaDesc) * 2 * VSB_MAX_SRC));  // [0,6): per-tap / halo / epilogue maps, [6,12): halo2 source maps
  return VSB_OK;
}

// Forget the current workspace layout (the arena itself is kept and only ever grows; it is
// released by release_arena() when the engine is destroyed or a new plan is loaded).
void free_workspace(vsb_engine* e) {
  e->tens.clear();
  e->ws_Hp = e->ws_Wp = e->ws_nb = 0;
}

// f32_plain: the tensor holds f32 and the box is stored densely (no swizzle) -- the logits
// tile of the shared-memory epilogue; otherwise 16-bit elements, swizzle by KB.
int make_tensor_map(vsb_engine* e, TmaDesc* out_host, const TensorBuf& t, int nb, bool folded, int KB,
                    int bw, int bh, int nt, bool f32_plain = false) {
  CUtensorMap m;
  const cuuint64_t C = t.C, W = t.W, H = t.H, eb = f32_plain ? 4 : 2;
  cuuint64_t dims[5], strides[4];
  if (!folded) {
    dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = nb;
    strides[0] = C * eb; strides[1] = W * C * eb; strides[2] = W * C * eb; strides[3] = H * W * C * eb;
  } else {
    dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = nb;
    strides[0] = 2 * C * eb; strides[1] = W * C * eb; strides[2] = 2 * W * C * eb; strides[3] = H * W * C * eb;
  }
  cuuint32_t box[5] = {(cuuint32_t)KB, (cuuint32_t)bw, 1u, (cuuint32_t)bh, (cuuint32_t)nt};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = f32_plain ? CU_TENSOR_MAP_SWIZZLE_NONE
                                          : (KB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                      : (KB == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B));
  const CUtensorMapDataType dt = f32_plain ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                           : (VSB_ACT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  const CUresult r = e->encode(&m, dt, 5, t.ptr, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(VSB_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): C=%d W=%d H=%d nb=%d folded=%d KB=%d box=%dx%dx%d", (int)r,
                t.C, t.W, t.H, nb, (int)folded, KB, bw, bh, nt);
  static_assert(sizeof(CUtensorMap) == sizeof(TmaDesc), "tensor map size");
  memcpy(out_host, &m, sizeof(m));
  return VSB_OK;
}

// Tensor map with caller-chosen dimensions (16-bit elements): the stem's raw-window loads and its
// phase-wise stores (conv_stem.cu, v3).  strides[i] = byte stride of dimension i + 1.
int make_custom_map(vsb_engine* e, TmaDesc* out_host, void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                    const cuuint32_t* box, bool swizzle128) {
  CUtensorMap m;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapDataType dt = VSB_ACT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUresult r = e->encode(&m, dt, (cuuint32_t)rank, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VSB_ERR_CUDA, "cuTensorMapEncodeTiled (custom, rank %d) failed (%d)", rank, (int)r);
  memcpy(out_host, &m, sizeof(m));
  return VSB_OK;
}

// Shared-memory epilogue + pipeline depth of one conv_halo launch.  `h` holds BN, n_tiles,
// ncs, b_bytes, a_stage_bytes; decides out/res staging, accumulator stages, resident vs
// streamed weights and the A / B ring depths inside the 227 KB of one SM.
int configure_halo_pipeline(vsb_engine* e, ConvPlan& cp, vsb::ConvHaloParams& h, const vsb_op& op, const TensorBuf& ot,
                            int nb) {
  h.acc_stages = h.BN <= 64 ? 8 : (h.BN <= 128 ? 4 : 2);
  h.out_map = h.res_map = nullptr;
  h.out_bufs = h.res_bufs = 0;
  h.out_buf_bytes = 0;
  h.res_inplace = 0;
  const bool has_res = op.res >= 0;
  bool tma_epi = !e->no_tma_epilogue && h.n_tiles == 1;
  if (ot.dtype == 0)
    // (grouped launches: a last block of fewer than 64 channels is clipped by the TMA store)
    tma_epi = tma_epi && ((h.BN % 64 == 0 && h.BN <= 128 && (op.cout % 64 == 0 || (cp.grouped_halo && op.cout % 8 == 0))) ||
                          (h.BN == 32 && op.cout == 32));
  else tma_epi = tma_epi && !has_res && op.cout % 4 == 0 && op.cout <= 32 && op.cout <= h.BN;
  if (tma_epi) {
    h.out_buf_bytes = ot.dtype == 0 ? (h.BN >= 64 ? (h.BN / 64) * 16384 : 8192)
                                    : (int)align_up((size_t)128 * op.cout * 4, 1024);
    h.out_bufs = h.out_buf_bytes <= 16384 ? 2 : 1;
    h.res_bufs = has_res ? h.out_bufs : 0;
    h.res_inplace = (has_res && h.out_bufs == 2 && ot.dtype == 0 && !e->no_res_inplace) ? 1 : 0;
    // Only where the weights stay resident next to the staging buffers: streamed-weight
    // launches (BN >= 128) need the shared memory for their B ring (measured: 5 instead of 8
    // weight stages cost more than the scattered stores).
    const size_t staging = (size_t)(h.out_bufs + (h.res_inplace ? 0 : h.res_bufs)) * h.out_buf_bytes;
    const size_t all = 227 * 1024 - 1024 - 1024 - 2048 * 4;
    if ((size_t)h.ncs * 9 * h.b_bytes + 3 * (size_t)h.a_stage_bytes + staging > all) {
      tma_epi = false;
      h.out_bufs = h.res_bufs = 0;
      h.out_buf_bytes = 0;
    }
  }
  if (!tma_epi) h.res_inplace = 0;
  h.epi_groups = (tma_epi && h.out_bufs == 2 && h.BN <= 64 && !e->no_epi_groups) ? 2 : 1;
  h.mma_warps = 1;
  const size_t usable =
      227 * 1024 - 1024 - 1024 - 2048 * 4 - (size_t)(h.out_bufs + (h.res_inplace ? 0 : h.res_bufs)) * h.out_buf_bytes;
  const size_t res_bytes = (size_t)h.ncs * 9 * h.b_bytes;
  if (h.n_tiles == 1 && res_bytes + 2 * (size_t)h.a_stage_bytes <= usable) {
    h.b_stages = 0;
    h.a_stages = (int)std::min<size_t>(e->halo_a_stages_max, (usable - res_bytes) / h.a_stage_bytes);
  } else {
    h.a_stages = 2;
    if (usable < 2 * (size_t)h.a_stage_bytes + 2 * (size_t)h.b_bytes) return 1;  // does not fit
    h.b_stages = (int)std::min<size_t>(vsb::HALO_MAX_B_STAGES, (usable - 2 * (size_t)h.a_stage_bytes) / h.b_bytes);
    if (h.b_stages >= 6 && usable >= 3 * (size_t)h.a_stage_bytes + 4 * (size_t)h.b_bytes) {
      h.a_stages = 3;
      h.b_stages = (int)std::min<size_t>(vsb::HALO_MAX_B_STAGES, (usable - 3 * (size_t)h.a_stage_bytes) / h.b_bytes);
    }
  }
  if (h.b_stages != 0 && e->halo_ab_override > 0) {  // tuning aid: a_stages * 10 + b_stages for streamed launches
    const int a = e->halo_ab_override / 10, b = e->halo_ab_override % 10;
    if (a >= 2 && b >= 2 && (size_t)a * h.a_stage_bytes + (size_t)b * h.b_bytes <= usable) {
      h.a_stages = a;
      h.b_stages = b;
    }
  }
  // two tiles per stage for streamed-weight launches (see ConvHaloParams::mt)
  h.mt = 1;
  // (BN = 256 would leave a single 512-column accumulator stage: the epilogue no longer overlaps the next
  // tile's MMAs -- measured slower for layer3, so only BN <= 128)
  if (h.b_stages != 0 && !tma_epi && !e->no_halo_mt && ot.W % 16 == 0 && h.BN <= e->halo_mt_max_bn) {
    const int HW2 = 16 + 2 * op.dil, HH2 = 16 + 2 * op.dil;
    const size_t a2 = align_up((size_t)HW2 * HH2 * 2 * h.kc, 1024);
    if (usable >= 2 * a2 + 3 * (size_t)h.b_bytes) {
      h.mt = 2;
      h.a_stage_bytes = (int)a2;
      h.a_stages = 2;
      h.b_stages = (int)std::min<size_t>(vsb::HALO_MAX_B_STAGES, (usable - 2 * a2) / h.b_bytes);
      h.acc_stages = h.BN <= 64 ? 4 : (h.BN <= 128 ? 2 : 1);
    }
  }
  // the store issuer of the shared-memory epilogue lives in the resident branch of the weight-producer warp
  if (tma_epi && h.b_stages != 0) return fail(VSB_ERR_UNSUPPORTED, "internal: shared-memory epilogue with streamed weights");
  // two MMA warps need the tile sequence to be the A-ring slab sequence (one slab per tile)
  if (h.b_stages == 0 && h.ncs == 1 && h.a_stages >= 2 && !e->no_mma2) {
    // A ring stage must always meet the same MMA warp: with an odd depth the two warps would
    // alternate on a stage, and the one running ahead could mistake the previous phase of that
    // stage's mbarrier for its own (parity aliasing) -- seen as a launch failure at depth 7.
    // Depth 3 keeps one warp and all three stages (measured faster than 2 + 2).
    if (h.a_stages != 3) {
      h.mma_warps = 2;
      h.a_stages &= ~1;
    }
  }
  if (tma_epi) {
    TmaDesc m;
    int rc;
    const int obox = h.BN >= 64 ? 64 : 32;  // channels per staging row (SW128 / SW64)
    if (ot.dtype == 0) rc = make_tensor_map(e, &m, ot, nb, false, obox, 8, 16, 1);
    else rc = make_tensor_map(e, &m, ot, nb, false, op.cout, 8, 16, 1, true);
    if (rc) return rc;
    CK(cudaMemcpy(cp.d_maps + (VSB_MAX_SRC - 2), &m, sizeof(m), cudaMemcpyHostToDevice));
    h.out_map = cp.d_maps + (VSB_MAX_SRC - 2);
    if (has_res) {
      rc = make_tensor_map(e, &m, e->tens[op.res], nb, false, obox, 8, 16, 1);
      if (rc) return rc;
      CK(cudaMemcpy(cp.d_maps + (VSB_MAX_SRC - 3), &m, sizeof(m), cudaMemcpyHostToDevice));
      h.res_map = cp.d_maps + (VSB_MAX_SRC - 3);
    }
  }
  return VSB_OK;
}

int build_workspace(vsb_engine* e, int Hp, int Wp, int nb) {
  if (e->ws_Hp == Hp && e->ws_Wp == Wp && e->ws_nb >= nb && e->ws_keep == e->keep_all) return VSB_OK;
  CK(cudaStreamSynchronize(e->stream));
  free_workspace(e);
  const int nt_ = (int)e->tdesc.size();
  e->tens.assign(nt_, TensorBuf());
  for (int t = 0; t < nt_; ++t) {
    TensorBuf& b = e->tens[t];
    b.C = e->tdesc[t].channels;
    b.ds = e->tdesc[t].ds_log2;
    b.dtype = e->tdesc[t].dtype;
    b.H = b.ds < 0 ? 1 : Hp >> b.ds;
    b.W = b.ds < 0 ? 1 : Wp >> b.ds;
    b.bytes = align_up((size_t)nb * b.H * b.W * b.C * (b.dtype ? 4 : 2), 1024);
  }
  // liveness
  e->tens[0].first_def = -1;
  for (int i = 0; i < (int)e->ops.size(); ++i) {
    const vsb_op& op = e->ops[i];
    if (op.out >= 0 && e->tens[op.out].first_def < 0 && op.out != 0) {
      e->tens[op.out].first_def = i;
      // a max-pool right after its producer may be computed inside the producer's launch (stem
      // fusion): its output must exist -- and alias nothing the producer reads -- one op earlier
      if (op.kind == VSB_OP_MAXPOOL && i > 0 && e->ops[i - 1].out == op.src[0]) e->tens[op.out].first_def = i - 1;
    }
    for (int s = 0; s < op.n_src; ++s) e->tens[op.src[s]].last_use = i;
    if (op.res >= 0) e->tens[op.res].last_use = i;
  }
  // Allocation with reuse of dead buffers (best fit) inside ONE arena: the layout is planned on
  // offsets first, then the arena only ever grows.  A direction change of an anisotropic volume
  // (new Hp x Wp) therefore costs no cudaFree / cudaMalloc churn (measured: 0.05 - 1.1 s per
  // direction with per-tensor allocations).
  struct Block { size_t off; size_t bytes; };
  std::vector<Block> freeb;
  size_t arena_end = 0;
  std::vector<size_t> offs(nt_, (size_t)-1), blk_bytes(nt_, 0);
  auto alloc = [&](size_t bytes, size_t* off, size_t* got) {
    int best = -1;
    if (!e->keep_all)
      for (int i = 0; i < (int)freeb.size(); ++i)
        if (freeb[i].bytes >= bytes && (best < 0 || freeb[i].bytes < freeb[best].bytes)) best = i;
    if (best >= 0 && freeb[best].bytes <= bytes * 2) {
      *off = freeb[best].off;
      *got = freeb[best].bytes;
      freeb.erase(freeb.begin() + best);
      return;
    }
    *off = arena_end;
    *got = bytes;
    arena_end += align_up(bytes, 1024);
  };
  alloc(e->tens[0].bytes, &offs[0], &blk_bytes[0]);
  for (int i = 0; i < (int)e->ops.size(); ++i) {
    for (int t = 1; t < nt_; ++t) {
      TensorBuf& b = e->tens[t];
      if (b.first_def != i || offs[t] != (size_t)-1) continue;
      alloc(b.bytes, &offs[t], &blk_bytes[t]);
    }
    // free tensors whose last use is this op
    for (int t = 0; t < nt_; ++t)
      if (offs[t] != (size_t)-1 && e->tens[t].last_use == i && e->tdesc[t].dtype == 0) freeb.push_back({offs[t], blk_bytes[t]});
  }
  if (arena_end > e->arena_bytes) {
    if (e->arena) cudaFree(e->arena);
    e->arena = nullptr;
    e->arena_bytes = 0;
    CK(cudaMalloc(&e->arena, arena_end));
    e->arena_bytes = arena_end;
  }
  for (int t = 0; t < nt_; ++t)
    if (offs[t] != (size_t)-1) e->tens[t].ptr = e->arena + offs[t];
  e->ws_Hp = Hp; e->ws_Wp = Wp; e->ws_nb = nb; e->ws_keep = e->keep_all;

  // tensor maps + tile geometry of the tcgen05 convolutions
  for (int i = 0; i < (int)e->ops.size(); ++i) {
    ConvPlan& cp = e->conv[i];
    if (cp.stem_tc) {
      const vsb_op& op = e->ops[i];
      const TensorBuf& ot = e->tens[op.out];
      const TensorBuf& st = e->tens[op.src[0]];
      vsb::ConvHalo2Params& h = cp.h2params;
      h = vsb::ConvHalo2Params{};
      h.stem = 1;
      h.kc = 64;
      h.mt = 1;
      h.n_src = 1;
      h.nslabs = 1;
      h.src[0].ptr = (const uint16_t*)st.ptr;
      h.src[0].C = 1;
      h.src[0].Hs = st.H;
      h.src[0].Ws = st.W;
      h.wpacked = cp.d_whalo;
      h.bias = cp.d_bias_pad;
      h.out = ot.ptr;
      h.out_f32 = 0;
      h.relu = op.relu;
      h.cout = op.cout;
      h.BN = cp.BN;
      h.n_tiles = 1;
      h.NB = nb;
      h.H = ot.H;
      h.W = ot.W;
      h.tiles_x = (ot.W + 7) / 8;
      h.tiles_y = (ot.H + 15) / 16;
      h.a_stage_bytes = 128 * 128;
      h.a_stages = 4;
      h.b_stages = 0;
      h.b_bytes = cp.BN * 128;
      // fuse the following 3x3/2 max-pool (torchvision ResNet.maxpool) into the epilogue
      cp.pool_op = -1;
      h.pool_out = nullptr;
      if (i + 1 < (int)e->ops.size()) {
        const vsb_op& nx = e->ops[i + 1];
        e->conv[i + 1].fused_away = false;
        if (!e->no_fuse_pool && nx.kind == VSB_OP_MAXPOOL && nx.src[0] == op.out && op.relu && cp.BN == 64 &&
            op.cout == 64 && ot.H % 16 == 0 && ot.W % 8 == 0 && e->tens[nx.out].H == ot.H / 2 &&
            e->tens[nx.out].W == ot.W / 2) {
          cp.pool_op = i + 1;
          h.pool_out = (uint16_t*)e->tens[nx.out].ptr;
          e->conv[i + 1].fused_away = true;
          cp.use_stem2 = false;
          if (e->stem_version >= 2 && !e->no_tma_epilogue && (ot.W & 3) == 0) {
            // in-CTA pooling kernels (conv_stem.cu): TMA-store maps of the conv output and the pooled output
            if (!cp.d_maps) CK(cudaMalloc(&cp.d_maps, sizeof(TmaDesc) * 2 * VSB_MAX_SRC));
            const int ver = e->stem_version >= 3 ? 3 : 2;
            TmaDesc m3[4];
            memset(m3, 0, sizeof(m3));
            int rc;
            if (ver == 3) {
              // conv output viewed as (c, x % 4, x / 4, y, n): one store per column phase
              const cuuint64_t od[5] = {64, 4, (cuuint64_t)ot.W / 4, (cuuint64_t)ot.H, (cuuint64_t)nb};
              const cuuint64_t os[4] = {128, 512, (cuuint64_t)ot.W * 128, (cuuint64_t)ot.H * ot.W * 128};
              const cuuint32_t ob[5] = {64, 1, 8, 14, 1};
              rc = make_custom_map(e, &m3[0], ot.ptr, 5, od, os, ob, true);
              if (rc) return rc;
              const cuuint32_t ob7[5] = {64, 1, 7, 14, 1};
              rc = make_custom_map(e, &m3[3], ot.ptr, 5, od, os, ob7, true);
              if (rc) return rc;
              rc = make_tensor_map(e, &m3[1], e->tens[nx.out], nb, false, 64, 15, 7, 1);
              if (rc) return rc;
              // network input as (x, y, n): the raw window of a tile, zero-filled outside the image
              const cuuint64_t img = (cuuint64_t)st.H * st.W * 2;
              const cuuint64_t id[5] = {(cuuint64_t)st.W, (cuuint64_t)st.H, (cuuint64_t)nb, 1, 1};
              const cuuint64_t is[4] = {(cuuint64_t)st.W * 2, img, img * nb, img * nb};
              const cuuint32_t ib[5] = {64, 38, 1, 1, 1};
              rc = make_custom_map(e, &m3[2], st.ptr, 5, id, is, ib, false);
              if (rc) return rc;
            } else {
              rc = make_tensor_map(e, &m3[0], ot, nb, false, 64, 14, 16, 1);
              if (rc) return rc;
              rc = make_tensor_map(e, &m3[1], e->tens[nx.out], nb, false, 64, 7, 8, 1);
              if (rc) return rc;
            }
            CK(cudaMemcpy(cp.d_maps + 2, m3, sizeof(m3), cudaMemcpyHostToDevice));
            vsb::ConvStemParams& sp = cp.sparams;
            sp = vsb::ConvStemParams{};
            sp.version = ver;
            sp.in = (const uint16_t*)st.ptr;
            sp.wpacked = cp.d_whalo;
            sp.bias = cp.d_bias_pad;
            sp.out_map = cp.d_maps + 2;
            sp.pool_map = cp.d_maps + 3;
            sp.in_map = cp.d_maps + 4;
            sp.out_map7 = cp.d_maps + 5;
            sp.NB = nb;
            sp.H = ot.H;
            sp.W = ot.W;
            sp.a_stages = 3;
            vsb::conv_stem_tiles(ver, ot.H, ot.W, &sp.tiles_x, &sp.tiles_y);
            cp.use_stem2 = true;
          }
          if (!e->no_tma_epilogue) {
            if (!cp.d_maps) CK(cudaMalloc(&cp.d_maps, sizeof(TmaDesc) * 2 * VSB_MAX_SRC));
            TmaDesc m;
            int rc = make_tensor_map(e, &m, ot, nb, false, 64, 8, 16, 1);
            if (rc) return rc;
            CK(cudaMemcpy(cp.d_maps, &m, sizeof(m), cudaMemcpyHostToDevice));
            h.out_map = cp.d_maps;
          }
        }
      }
      continue;
    }
    if (cp.grouped_halo) {
      const vsb_op& op = e->ops[i];
      const TensorBuf& ot = e->tens[op.out];
      const TensorBuf& st = e->tens[op.src[0]];
      const int tx = (ot.W + 7) / 8, ty = (ot.H + 15) / 16;
      const double eff = (double)ot.W * ot.H / ((double)tx * 8 * ty * 16);
      cp.use_halo = false;
      if (eff >= 0.5) {
        vsb::ConvHaloParams& h = cp.hparams;
        h = vsb::ConvHaloParams{};
        const int HW = 8 + 2 * op.dil, HH = 16 + 2 * op.dil;
        TmaDesc hm[VSB_MAX_SRC];
        memset(hm, 0, sizeof(hm));
        for (int s = 0; s < op.n_src; ++s) {
          int rc = make_tensor_map(e, &hm[s], e->tens[op.src[s]], nb, false, 64, HW, HH, 1);
          if (rc) return rc;
        }
        CK(cudaMemcpy(cp.d_maps, hm, sizeof(hm), cudaMemcpyHostToDevice));
        h.map = cp.d_maps;
        h.bias = cp.d_bias_pad;
        h.residual = op.res >= 0 ? (const uint16_t*)e->tens[op.res].ptr : nullptr;
        h.out = ot.ptr;
        h.out_f32 = 0;
        h.relu = op.relu;
        h.cout = op.cout;
        h.BN = 64;
        h.n_tiles = 1;
        h.NB = nb;
        h.H = ot.H;
        h.W = ot.W;
        h.ncs = 1;
        h.dil = op.dil;
        h.tiles_x = tx;
        h.tiles_y = ty;
        h.kc = 64;
        h.b_bytes = 64 * 128;
        h.a_stage_bytes = (int)align_up((size_t)HW * HH * 128, 1024);
        {
          int rc2 = configure_halo_pipeline(e, cp, h, op, ot, nb);
          if (rc2 < 0) return rc2;
          if (rc2 > 0 || h.b_stages != 0) return fail(VSB_ERR_UNSUPPORTED, "op %d: grouped halo launch does not fit", i);
        }
        cp.use_halo = true;
      }
      continue;
    }
    if (!cp.tc) continue;
    const vsb_op& op = e->ops[i];
    const TensorBuf& ot = e->tens[op.out];
    const int gw = cp.ps ? ot.W / 2 : ot.W, gh = cp.ps ? ot.H / 2 : ot.H;  // box grid
    const int per_cls = cp.ps ? 32 : 128;
    int bw = std::min(pow2ceil(gw), cp.ps ? 8 : 16);
    int bh = std::min(pow2ceil(gh), per_cls / bw);
    int ntile = per_cls / (bw * bh);
    TmaDesc maps[VSB_MAX_SRC];
    memset(maps, 0, sizeof(maps));
    for (int s = 0; s < op.n_src; ++s) {
      const TensorBuf& st = e->tens[op.src[s]];
      bool folded;
      if (cp.ps) folded = !op.src_up[s];
      else folded = cp.s2;
      if (folded && ((st.H & 1) || (st.W & 1)))
        return fail(VSB_ERR_UNSUPPORTED, "op %d: folded source with odd dims %dx%d", i, st.H, st.W);
      int rc = make_tensor_map(e, &maps[s], st, nb, folded, cp.kb, bw, bh, ntile);
      if (rc) return rc;
    }
    CK(cudaMemcpy(cp.d_maps, maps, sizeof(maps), cudaMemcpyHostToDevice));
    vsb::ConvTcParams& p = cp.params;
    p.maps = cp.d_maps;
    p.runs = cp.d_runs;
    p.num_runs = (int)cp.runs.size();
    p.num_slabs = cp.num_slabs;
    p.row_bytes = cp.kb * 2;
    p.wpacked = cp.d_wpacked;
    p.bias = cp.d_bias_pad;
    p.residual = op.res >= 0 ? (const uint16_t*)e->tens[op.res].ptr : nullptr;
    p.out = ot.ptr;
    p.out_f32 = ot.dtype;
    p.relu = op.relu;
    p.cout = op.cout;
    p.BN = cp.BN;
    p.n_tiles = cp.n_tiles;
    p.NB = nb;
    p.H = ot.H;
    p.W = ot.W;
    p.ncls_log2 = cp.ps ? 2 : 0;
    p.bw_log2 = ilog2(bw);
    p.bh_log2 = ilog2(bh);
    p.nt_log2 = ilog2(ntile);
    p.tiles_x = (gw + bw - 1) / bw;
    p.tiles_y = (gh + bh - 1) / bh;
    p.tiles_n = (nb + ntile - 1) / ntile;
    p.a_bytes = 128 * p.row_bytes;
    p.stage_bytes = (128 + cp.BN) * p.row_bytes;
    p.num_stages = std::min<int>(vsb::TC_MAX_STAGES, (int)((206 * 1024) / p.stage_bytes));
    if (p.num_stages < 2) return fail(VSB_ERR_UNSUPPORTED, "op %d: too few pipeline stages", i);
    cp.use_halo2 = false;
    if (cp.halo2_ok) {
      vsb::ConvHalo2Params& h = cp.h2params;
      const int P = 2 * h.kc;
      int mt = std::min(h.kc == 64 ? 2 : 4, std::max(1, 256 / cp.BN));
      mt = mt >= 4 ? 4 : (mt >= 2 ? 2 : 1);
      if (cp.BN & (cp.BN - 1)) mt = 1;  // the epilogue maps column -> (tile, channel) with shifts
      int tx = 0, ty = (ot.H + 15) / 16;
      double eff = 0;
      for (;; mt /= 2) {
        tx = (ot.W + 8 * mt - 1) / (8 * mt);
        eff = (double)ot.W * ot.H / ((double)tx * 8 * mt * ty * 16);
        if (eff >= 0.6 || mt == 1) break;
      }
      h.mt = mt;
      h.BN = cp.BN;
      h.n_tiles = cp.n_tiles;
      h.a_stage_bytes = (int)align_up((size_t)(8 * mt + 2) * 18 * P, 1024);
      const size_t budget = 200 * 1024;
      const size_t res_bytes = (size_t)h.nslabs * 9 * h.b_bytes;
      if (cp.n_tiles == 1 && res_bytes + 3 * (size_t)h.a_stage_bytes <= budget) {
        h.b_stages = 0;
        h.a_stages = (int)std::min<size_t>(vsb::HALO_MAX_A_STAGES, (budget - res_bytes) / h.a_stage_bytes);
      } else {
        h.a_stages = 3;
        h.b_stages = (int)std::min<size_t>(vsb::HALO_MAX_B_STAGES, (budget - 3 * (size_t)h.a_stage_bytes) / h.b_bytes);
      }
      h.mma_warps = 1;
      // measured neutral (these layers are bound by the cp.async halo assembly, not by MMA issue):
      // off unless vsb_set_flag("halo2_mma2", 1)
      if (e->halo2_mma2 && h.b_stages == 0 && h.a_stages >= 4 && !e->no_mma2) {
        h.mma_warps = 2;
        h.a_stages &= ~1;  // two half-rings, one per MMA warp
      }
      if (eff >= 0.6 && (h.b_stages == 0 || h.b_stages >= 2)) {
        for (int s = 0; s < op.n_src; ++s) {
          const TensorBuf& st = e->tens[op.src[s]];
          h.src[s].ptr = (const uint16_t*)st.ptr;
          h.src[s].C = st.C;
          h.src[s].Hs = st.H;
          h.src[s].Ws = st.W;
          h.src[s].up = op.src_up[s];
          // sources read at their own resolution arrive by ONE TMA box per slab (same swizzled
          // compact-pitch layout the cp.async loaders write); up-sampled ones keep the loaders
          h.src_map[s] = nullptr;
          if (!op.src_up[s] && !e->no_halo2_tma) {
            TmaDesc m;
            int rc = make_tensor_map(e, &m, st, nb, false, h.kc, 8 * mt + 2, 18, 1);
            if (rc) return rc;
            CK(cudaMemcpy(cp.d_maps + VSB_MAX_SRC + s, &m, sizeof(m), cudaMemcpyHostToDevice));
            h.src_map[s] = cp.d_maps + VSB_MAX_SRC + s;
          }
        }
        h.wpacked = cp.d_whalo;
        h.bias = cp.d_bias_pad;
        h.residual = p.residual;
        h.out = ot.ptr;
        h.out_f32 = ot.dtype;
        h.relu = op.relu;
        h.cout = op.cout;
        h.NB = nb;
        h.H = ot.H;
        h.W = ot.W;
        h.tiles_x = tx;
        h.tiles_y = ty;
        cp.use_halo2 = true;
      }
    }
    cp.use_halo = false;
    if (cp.halo_ok) {
      const int tx = (ot.W + 7) / 8, ty = (ot.H + 15) / 16;
      const double eff = (double)ot.W * ot.H / ((double)tx * 8 * ty * 16);
      vsb::ConvHaloParams& h = cp.hparams;
      h.dil = op.dil;
      const int HW = 8 + 2 * op.dil, HH = 16 + 2 * op.dil;
      const int kc = h.kc ? h.kc : 64;
      h.kc = kc;
      h.ncs = op.cin / kc;
      h.BN = cp.BN;
      h.n_tiles = cp.n_tiles;
      h.b_bytes = cp.BN * 2 * kc;
      h.a_stage_bytes = (int)align_up((size_t)HW * HH * 2 * kc, 1024);
      int fit = configure_halo_pipeline(e, cp, h, op, ot, nb);
      if (fit < 0) return fit;
      if (eff >= 0.6 && fit == 0 && (h.b_stages == 0 || h.b_stages >= 2)) {
        TmaDesc hm;
        const TensorBuf& st = e->tens[op.src[0]];
        int rc = make_tensor_map(e, &hm, st, nb, false, kc, 8 * h.mt + 2 * op.dil, HH, 1);
        if (rc) return rc;
        // the halo map lives in slot VSB_MAX_SRC - 1 of this op's map array (single-source op)
        CK(cudaMemcpy(cp.d_maps + (VSB_MAX_SRC - 1), &hm, sizeof(hm), cudaMemcpyHostToDevice));
        h.map = cp.d_maps + (VSB_MAX_SRC - 1);
        h.wpacked = cp.d_whalo;
        h.bias = cp.d_bias_pad;
        h.residual = p.residual;
        h.out = ot.ptr;
        h.out_f32 = ot.dtype;
        h.relu = op.relu;
        h.cout = op.cout;
        h.NB = nb;
        h.H = ot.H;
        h.W = ot.W;
        h.tiles_x = (ot.W + 8 * h.mt - 1) / (8 * h.mt);
        h.tiles_y = ty;
        cp.use_halo = true;
      }
    }
    cp.use_el = false;
    if (cp.el_ok && (cp.el_kind == 2 || (!(ot.H & 1) && !(ot.W & 1)))) {
      vsb::ConvHaloElParams& h = cp.elparams;
      // the grid the lowered convolution runs on: half the output (space-to-depth output) or the output itself
      const int Hs = cp.el_kind == 1 ? ot.H / 2 : ot.H, Ws = cp.el_kind == 1 ? ot.W / 2 : ot.W;
      const int tx = (Ws + 8 * h.mt - 1) / (8 * h.mt), ty = (Hs + 15) / 16;
      const double eff = (double)Ws * Hs / ((double)tx * 8 * h.mt * ty * 16);
      const int HW = 8 * h.mt + 2, HH = 18;
      h.a_stage_bytes = (int)align_up((size_t)HW * HH * 128, 1024);
      h.b_bytes = h.BN * 128;
      // a slab feeds 16..72 MMAs, far longer than a halo box takes to arrive: two stages; the rest is weight ring
      h.a_stages = e->el_a_stages;
      const bool tma_epi = !e->no_el_tma_epilogue;
      const size_t fixed = vsb::conv_halo_el_smem_bytes(vsb::ConvHaloElParams{}) + (tma_epi ? 32768 : 0);  // control, bias, tables, slack, staging
      const size_t room = 227 * 1024 - fixed - (size_t)h.a_stages * h.a_stage_bytes;
      h.b_stages = (int)std::min<size_t>(vsb::HALO_MAX_B_STAGES, room / h.b_bytes);
      if (eff >= 0.55 && h.b_stages >= 2) {
        TmaDesc maps[VSB_MAX_SRC];
        memset(maps, 0, sizeof(maps));
        bool ok = true;
        for (int s = 0; s < op.n_src && ok; ++s) {
          const TensorBuf& st = e->tens[op.src[s]];
          const bool folded = !(op.mode == 2 && s == 0);  // all but the up-sampled source are read as parity planes
          ok = folded ? (st.H == 2 * Hs && st.W == 2 * Ws) : (st.H == Hs && st.W == Ws);
          if (!ok) break;
          int rc = make_tensor_map(e, &maps[s], st, nb, folded, 64, HW, HH, 1);
          if (rc) return rc;
        }
        TmaDesc om;
        if (ok && tma_epi) {
          int rc = make_tensor_map(e, &om, ot, nb, cp.el_kind == 1, 64, 8, 16, 1);
          if (rc) return rc;
        }
        if (ok) {
          CK(cudaMemcpy(cp.d_el_maps, maps, sizeof(maps), cudaMemcpyHostToDevice));
          h.out_map = nullptr;
          if (tma_epi) {
            CK(cudaMemcpy(cp.d_el_maps + VSB_MAX_SRC, &om, sizeof(om), cudaMemcpyHostToDevice));
            h.out_map = cp.d_el_maps + VSB_MAX_SRC;
          }
          h.map = cp.d_el_maps;
          h.out = ot.ptr;
          h.NB = nb;
          h.H = Hs;
          h.W = Ws;
          h.tiles_x = tx;
          h.tiles_y = ty;
          cp.use_el = true;
        }
      }
    }
    // Per-tap kernel launches (1x1 and stride-2 convolutions): epilogue through shared memory + TMA store
    // where the tile is a plain box of the output tensor (conv_tc.cuh)
    p.out_map = p.res_map = nullptr;
    if (!cp.use_halo && !cp.use_halo2 && !cp.grouped_s2 && !e->no_tc_smem_epilogue && !e->no_tma_epilogue && !cp.ps &&
        ot.dtype == 0 && op.cout % 64 == 0 && cp.BN % 64 == 0 && cp.BN * cp.n_tiles == op.cout) {
      const bool has_res = op.res >= 0;
      const size_t budget = 206 * 1024 - vsb::conv_tc_epilogue_bytes(has_res);
      const int stages = std::min<int>(vsb::TC_MAX_STAGES, (int)(budget / p.stage_bytes));
      if (stages >= 2) {
        TmaDesc m2[2];
        memset(m2, 0, sizeof(m2));
        int rc = make_tensor_map(e, &m2[0], ot, nb, false, 64, bw, bh, ntile);
        if (rc) return rc;
        if (has_res) {
          rc = make_tensor_map(e, &m2[1], e->tens[op.res], nb, false, 64, bw, bh, ntile);
          if (rc) return rc;
        }
        CK(cudaMemcpy(cp.d_maps + VSB_MAX_SRC, m2, sizeof(m2), cudaMemcpyHostToDevice));
        p.out_map = cp.d_maps + VSB_MAX_SRC;
        p.res_map = has_res ? cp.d_maps + VSB_MAX_SRC + 1 : nullptr;
        p.num_stages = stages;
      }
    }
  }
  return VSB_OK;
}

// ---- profiling helpers -----------------------------------------------------
struct ProfScope {
  vsb_engine* e;
  int idx = -1;
  ProfScope(vsb_engine* e_, int cls, int op = -1) : e(e_) {
    e->launches += 1;
    if (!e->profiling) return;
    if (e->ev_used == e->ev_pool.size()) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      e->ev_pool.push_back({a, b});
      e->ev_cls.push_back(cls);
      e->ev_op.push_back(op);
    }
    idx = (int)e->ev_used++;
    e->ev_cls[idx] = cls;
    e->ev_op[idx] = op;
    cudaEventRecord(e->ev_pool[idx].first, e->stream);
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(e->ev_pool[idx].second, e->stream);
  }
};

void prof_collect(vsb_engine* e) {
  if (!e->profiling) return;
  cudaStreamSynchronize(e->stream);
  for (size_t i = 0; i < e->ev_used; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->ev_pool[i].first, e->ev_pool[i].second);
    e->prof_ms[e->ev_cls[i]] += ms;
    e->prof_launches[e->ev_cls[i]] += 1;
    const int op = e->ev_op[i];
    if (op >= 0) {
      if ((int)e->op_ms.size() <= op) {
        e->op_ms.resize(op + 1, 0.f);
        e->op_launches.resize(op + 1, 0);
      }
      e->op_ms[op] += ms;
      e->op_launches[op] += 1;
    }
  }
  e->ev_used = 0;
}

// ---- executor ----------------------------------------------------------------
// Runs op `oi` on images [n0, n0 + nb) of the current batch.
int run_conv(vsb_engine* e, int oi, int n0, int nb) {
  const vsb_op& op = e->ops[oi];
  ConvPlan& cp = e->conv[oi];
  const TensorBuf& ot = e->tens[op.out];
  if (cp.use_el && e->conv_impl == 0 && !e->no_halo && !(op.mode == 2 ? e->no_s2d_up : e->no_el_conv)) {
    vsb::ConvHaloElParams h = cp.elparams;
    h.NB = nb;
    h.n_base = n0;
    h.dbg = e->halo_dbg;
    ProfScope ps(e, PC_CONV_TC, oi);
    CK(vsb::launch_conv_halo_el(h, e->num_sms, e->stream));
    return VSB_OK;
  }
  if (cp.tc && cp.use_halo && e->conv_impl == 0 && !e->no_halo) {
    vsb::ConvHaloParams h = cp.hparams;
    h.NB = nb;
    h.n_base = n0;
    h.dbg = e->halo_dbg;
    h.prof = e->d_halo_prof;
    if (e->d_halo_prof) CK(cudaMemsetAsync(e->d_halo_prof, 0, 32 * 8, e->stream));
    {
      ProfScope ps(e, PC_CONV_TC, oi);
      CK(vsb::launch_conv_halo(h, e->num_sms, e->stream));
    }
    if (e->d_halo_prof) {
      unsigned long long v[32];
      CK(cudaMemcpyAsync(v, e->d_halo_prof, sizeof(v), cudaMemcpyDeviceToHost, e->stream));
      CK(cudaStreamSynchronize(e->stream));
      fprintf(stderr,
              "[halo_prof] op %d BN %d kc %d ncs %d a_stages %d b_stages %d acc %d groups %d mma %d | producer: total %llu wait_a_empty %llu tiles %llu"
              " | mma0: total %llu wait_acc_empty %llu wait_a_full %llu issue %llu | epi0: total %llu acc_full %llu"
              " ld+waits %llu math %llu publish %llu loop %llu tiles %llu\n",
              oi, h.BN, h.kc, h.ncs, h.a_stages, h.b_stages, h.acc_stages, h.epi_groups, h.mma_warps, v[0], v[1], v[2], v[8],
              v[9], v[10], v[11], v[16], v[18], v[19], v[20], v[21], v[22], v[23]);
    }
    return VSB_OK;
  }
  if (cp.stem_tc && cp.use_stem2 && cp.pool_op >= 0 && e->conv_impl == 0 && !e->no_halo) {
    vsb::ConvStemParams sp = cp.sparams;
    sp.NB = nb;
    sp.n_base = n0;
    sp.dbg = e->stem_dbg;
    ProfScope ps(e, PC_STEM, oi);
    CK(vsb::launch_conv_stem(sp, e->num_sms, e->stream));
    return VSB_OK;
  }
  if (cp.stem_tc && e->conv_impl == 0 && !e->no_halo) {
    vsb::ConvHalo2Params h = cp.h2params;
    h.NB = nb;
    h.n_base = n0;
    if (cp.pool_op >= 0) {
      // the pooled tensor is max-reduced into: zero is the identity of max over ReLU outputs
      const TensorBuf& pt = e->tens[e->ops[cp.pool_op].out];
      const size_t per_img = (size_t)pt.H * pt.W * pt.C * 2;
      ProfScope ps(e, PC_POOL, cp.pool_op);
      e->launches -= 1;  // a driver memset, not one of our kernels
      CK(cudaMemsetAsync((uint8_t*)pt.ptr + (size_t)n0 * per_img, 0, (size_t)nb * per_img, e->stream));
    }
    ProfScope ps(e, PC_STEM, oi);
    CK(vsb::launch_conv_halo2(h, e->num_sms, e->stream));
    return VSB_OK;
  }
  const bool dw_simt = cp.depthwise && e->no_dw_tc;  // vsb_set_flag("dw_tc", 0): depthwise on the CUDA-core kernel
  if (cp.grouped_halo && cp.use_halo && e->conv_impl == 0 && !e->no_halo && !dw_simt) {
    for (int b = 0; b < cp.n_blocks; ++b) {
      vsb::ConvHaloParams h = cp.hparams;
      h.NB = nb;
      h.n_base = n0;
      h.map = cp.d_maps + cp.blk_src[b];
      h.cin_off = cp.blk_cin_off[b];
      h.cout_off = b * 64;
      h.wpacked = cp.d_whalo + (size_t)b * 9 * 64 * 128;
      ProfScope ps(e, PC_CONV_TC, oi);
      CK(vsb::launch_conv_halo(h, e->num_sms, e->stream));
    }
    return VSB_OK;
  }
  if (cp.tc && cp.use_halo2 && e->conv_impl == 0 && !e->no_halo) {
    vsb::ConvHalo2Params h = cp.h2params;
    h.NB = nb;
    h.n_base = n0;
    if (e->fuse_op == oi) {
      h.head = e->fuse;
      h.head.s0 += n0;
    }
    ProfScope ps(e, e->fuse_op == oi ? PC_HEAD : PC_CONV_TC, oi);
    CK(vsb::launch_conv_halo2(h, e->num_sms, e->stream));
    return VSB_OK;
  }
  if (cp.grouped_s2 && e->conv_impl == 0 && !dw_simt) {
    for (int b = 0; b < cp.n_blocks; ++b) {
      vsb::ConvTcParams p = cp.params;
      p.NB = nb;
      p.n_base = n0;
      p.tiles_n = (nb + (1 << p.nt_log2) - 1) >> p.nt_log2;
      p.cin_off = b * 64;
      p.cout_off = b * 64;
      p.wpacked = cp.d_wpacked + (size_t)b * 9 * 64 * 128;
      ProfScope ps(e, PC_CONV_TC, oi);
      CK(vsb::launch_conv_tc(p, e->num_sms, e->stream));
    }
    return VSB_OK;
  }
  if (cp.tc && !cp.grouped_s2 && !cp.depthwise && e->conv_impl == 0) {
    vsb::ConvTcParams p = cp.params;
    p.NB = nb;
    p.n_base = n0;
    p.tiles_n = (nb + (1 << p.nt_log2) - 1) >> p.nt_log2;
    ProfScope ps(e, PC_CONV_TC, oi);
    CK(vsb::launch_conv_tc(p, e->num_sms, e->stream));
    return VSB_OK;
  }
  const TensorBuf& s0 = e->tens[op.src[0]];
  if (e->conv_impl != 2 && op.cin == 1 && op.kh == 7 && op.kw == 7 && op.stride == 2 && op.pad == 3 &&
      op.cout == 64 && op.n_src == 1 && ot.dtype == 0 && op.res < 0) {
    ProfScope ps(e, PC_STEM, oi);
    vsb::launch_stem7x7((const uint16_t*)s0.ptr + (size_t)n0 * s0.H * s0.W, nb, s0.H, s0.W, e->d_weights + op.w_off,
                        (const float*)(e->d_weights + op.b_off),
                        (uint16_t*)ot.ptr + (size_t)n0 * ot.H * ot.W * ot.C, op.relu, e->stream);
    CK(cudaGetLastError());
    return VSB_OK;
  }
  vsb::ConvArgs a{};
  a.n_src = op.n_src;
  for (int s = 0; s < op.n_src; ++s) {
    const TensorBuf& st = e->tens[op.src[s]];
    a.src[s].ptr = (const uint8_t*)st.ptr + (size_t)n0 * st.H * st.W * st.C * 2;
    a.src[s].C = st.C;
    a.src[s].H = st.H;
    a.src[s].W = st.W;
    a.src[s].up = op.src_up[s];
  }
  a.NB = nb; a.H = ot.H; a.W = ot.W;
  a.cin = op.cin; a.cout = op.cout; a.kh = op.kh; a.kw = op.kw;
  a.stride = op.stride; a.pad = op.pad; a.dil = op.dil; a.groups = op.groups; a.relu = op.relu;
  a.weights = e->d_weights + op.w_off;
  a.bias = op.b_off >= 0 ? (const float*)(e->d_weights + op.b_off) : nullptr;
  const size_t out_off = (size_t)n0 * ot.H * ot.W * ot.C;
  a.residual = op.res >= 0 ? (const uint8_t*)e->tens[op.res].ptr + out_off * 2 : nullptr;
  a.out = (uint8_t*)ot.ptr + out_off * (ot.dtype ? 4 : 2);
  a.out_f32 = ot.dtype;
  if (cp.depthwise && e->conv_impl != 2) {
    a.weights = cp.d_wdw;
    ProfScope ps(e, PC_OTHER, oi);
    vsb::launch_dwconv3x3(a, e->stream, !e->no_dw_tiled);
    CK(cudaGetLastError());
    return VSB_OK;
  }
  ProfScope ps(e, PC_CONV_SIMT, oi);
  vsb::launch_conv_simt(a, e->stream);
  CK(cudaGetLastError());
  return VSB_OK;
}

// Runs one op (not HEAD) on images [n0, n0 + nb) of the batch.
int run_op(vsb_engine* e, int i, int n0, int nb) {
  const vsb_op& op = e->ops[i];
  switch (op.kind) {
    case VSB_OP_CONV:
      return run_conv(e, i, n0, nb);
    case VSB_OP_MAXPOOL: {
      if (e->conv[i].fused_away && e->conv_impl == 0 && !e->no_halo) return VSB_OK;  // done by the stem epilogue
      const TensorBuf& s = e->tens[op.src[0]];
      const TensorBuf& o = e->tens[op.out];
      ProfScope ps(e, PC_POOL, i);
      vsb::launch_maxpool3x3s2((const uint16_t*)s.ptr + (size_t)n0 * s.H * s.W * s.C, nb, s.H, s.W, s.C,
                               (uint16_t*)o.ptr + (size_t)n0 * o.H * o.W * o.C, e->stream);
      CK(cudaGetLastError());
      return VSB_OK;
    }
    case VSB_OP_GAP: {
      const TensorBuf& s = e->tens[op.src[0]];
      const TensorBuf& o = e->tens[op.out];
      const size_t need = vsb::gap_scratch_bytes(nb, s.C);
      if (need > e->gap_scratch_bytes) {
        CK(cudaStreamSynchronize(e->stream));
        cudaFree(e->d_gap_scratch);
        e->d_gap_scratch = nullptr;
        e->gap_scratch_bytes = 0;
        CK(cudaMalloc(&e->d_gap_scratch, need));
        e->gap_scratch_bytes = need;
      }
      ProfScope ps(e, PC_OTHER, i);
      e->launches += 1;  // two kernels: partial sums + finish
      vsb::launch_gap((const uint16_t*)s.ptr + (size_t)n0 * s.H * s.W * s.C, nb, s.H, s.W, s.C,
                      (uint16_t*)o.ptr + (size_t)n0 * o.C, e->d_gap_scratch, e->stream);
      CK(cudaGetLastError());
      return VSB_OK;
    }
    case VSB_OP_UPSAMPLE: {
      const TensorBuf& s = e->tens[op.src[0]];
      const TensorBuf& o = e->tens[op.out];
      ProfScope ps(e, PC_OTHER, i);
      vsb::launch_upsample((const uint16_t*)s.ptr + (size_t)n0 * s.H * s.W * s.C, nb, s.H, s.W, s.C, o.H, o.W,
                           op.mode, (uint16_t*)o.ptr + (size_t)n0 * o.H * o.W * o.C, e->stream);
      CK(cudaGetLastError());
      return VSB_OK;
    }
    default:
      return fail(VSB_ERR_INVALID, "unknown op kind %d", op.kind);
  }
}

// Runs every op except HEAD on tensor 0 (already filled).  Returns the index of the
// HEAD op through *head_idx (or -1).
// Ops are grouped into segments of "shallow" (output at >= 1/4 resolution) and "deep"
// ops.  Deep segments run on the whole batch (they need many images to fill 148 SMs);
// shallow segments run depth-first over sub-batches small enough that the tensors a
// layer hands to the next one stay in the 126 MB L2 instead of streaming through HBM.
int run_network(vsb_engine* e, int nb, int* head_idx) {
  *head_idx = -1;
  const int n_ops = (int)e->ops.size();
  auto shallow = [&](int i) {
    const vsb_op& op = e->ops[i];
    if (op.kind == VSB_OP_HEAD || op.out < 0) return false;
    const int ds = e->tdesc[op.out].ds_log2;
    return ds >= 0 && ds <= 2;
  };
  int i = 0;
  while (i < n_ops) {
    if (e->ops[i].kind == VSB_OP_HEAD) {
      *head_idx = i++;
      continue;
    }
    const bool sh = shallow(i);
    int j = i;
    size_t max_bytes = 1;
    while (j < n_ops && e->ops[j].kind != VSB_OP_HEAD && shallow(j) == sh) {
      const TensorBuf& o = e->tens[e->ops[j].out];
      max_bytes = std::max(max_bytes, (size_t)o.H * o.W * o.C * (o.dtype ? 4 : 2));
      ++j;
    }
    int sub = nb;
    if (sh && e->sub_batch_mb > 0)
      sub = (int)std::max<size_t>(1, std::min<size_t>(nb, ((size_t)e->sub_batch_mb << 20) / max_bytes));
    for (int n0 = 0; n0 < nb; n0 += sub) {
      const int cnt = std::min(sub, nb - n0);
      for (int k = i; k < j; ++k) {
        int rc = run_op(e, k, n0, cnt);
        if (rc) return rc;
        if (e->sync_each) {  // debugging aid (vsb_set_flag("sync_each", 1)): localise a failing launch
          const cudaError_t se = cudaStreamSynchronize(e->stream);
          if (se != cudaSuccess)
            return fail(VSB_ERR_CUDA, "op %d (kind %d, cin %d, cout %d) failed: %s", k, e->ops[k].kind, e->ops[k].cin,
                        e->ops[k].cout, cudaGetErrorString(se));
        }
      }
    }
    i = j;
  }
  return VSB_OK;
}

int auto_batch(const vsb_engine* e, int64_t Hp, int64_t Wp, int64_t S, bool xplane) {
  if (e->batch_override > 0) return (int)std::min<int64_t>(e->batch_override, S);
  // 128 slices of 1024^2 per launch sequence for every direction.  Round 1 used 32 for the row directions
  // ("larger batches are neutral"); with the per-launch fixed cost now about 20 us against 180 us of work per
  // 32-slice launch, 128 measured 3.3 % faster (266 -> 258 us per slice, tests/exp_stem.sh).  x-plane
  // directions (slice index along x) need 128 anyway: the slicer then reads 128 contiguous bytes per
  // (row, column) and the head merges 1 KB runs of keys instead of 256 B.
  (void)xplane;
  const int64_t target_px = e->row_batch_px > 0 && !xplane ? e->row_batch_px : (128ll << 20);
  int64_t nb = std::max<int64_t>(1, target_px / (Hp * Wp));
  nb = std::min<int64_t>(nb, 256);
  if (nb >= 32) nb = nb / 32 * 32;
  else if (nb >= 8) nb = nb / 8 * 8;
  return (int)std::min<int64_t>(nb, S);
}

int predict_range(vsb_engine* e, int d, int64_t s_begin, int64_t s_end) {
  if (!e->has_plan) return fail(VSB_ERR_STATE, "no plan loaded");
  if (!e->d_vol) return fail(VSB_ERR_STATE, "no volume set");
  vsb_direction g;
  int rc = direction_geometry(e->Z, e->Y, e->X, d, &g);
  if (rc) return rc;
  if (s_begin < 0 || s_end > g.S || s_begin > s_end) return fail(VSB_ERR_INVALID, "bad slice range");
  if (s_begin == s_end) return VSB_OK;
  if (e->vote_mode && !e->d_votes) return fail(VSB_ERR_STATE, "vote buffer missing");
  const int nbmax = auto_batch(e, g.Hp, g.Wp, s_end - s_begin, g.stride_s == 1);
  e->keep_all = false;
  rc = build_workspace(e, (int)g.Hp, (int)g.Wp, nbmax);
  if (rc) return rc;
  const int nbw = e->ws_nb;
  (void)nbw;
  for (int64_t s0 = s_begin; s0 < s_end; s0 += nbmax) {
    const int nb = (int)std::min<int64_t>(nbmax, s_end - s0);
    {
      ProfScope ps(e, PC_SLICER);
      if (e->vol_dtype == 2) vsb::launch_slicer(e->d_vol, g, s0, nb, (uint16_t*)e->tens[0].ptr, e->stream);
      else vsb::launch_slicer_typed(e->d_vol, e->vol_dtype, g, s0, nb, (uint16_t*)e->tens[0].ptr, e->stream);
      CK(cudaGetLastError());
    }
    // fused head: the conv that produces the logits merges them in its own epilogue
    e->fuse_op = -1;
    {
      int hidx = -1;
      for (int i = 0; i < (int)e->ops.size(); ++i)
        if (e->ops[i].kind == VSB_OP_HEAD) hidx = i;
      if (hidx >= 0 && !e->vote_mode && !e->no_fuse_head && !e->no_halo && e->conv_impl == 0 &&
          e->num_classes <= 8 && (e->ops[hidx].factor <= 1) && e->ops[hidx].mode == 0) {
        const int lt = e->ops[hidx].src[0];
        for (int i = 0; i < hidx; ++i)
          if (e->ops[i].kind == VSB_OP_CONV && e->ops[i].out == lt && e->conv[i].tc && e->conv[i].use_halo2 &&
              e->conv[i].n_tiles == 1 && e->conv[i].BN >= 8 && (e->conv[i].BN & (e->conv[i].BN - 1)) == 0)
            e->fuse_op = i;
      }
      if (e->fuse_op >= 0) {
        vsb::HeadFuse& f = e->fuse;
        f.on = 1;
        f.C = e->num_classes;
        f.d = d;
        f.Hc = (int)g.H;
        f.Wc = (int)g.W;
        f.crop_top = (int)g.crop_top;
        f.crop_left = (int)g.crop_left;
        f.s0 = s0;
        f.base = g.base;
        f.stride_s = g.stride_s;
        f.stride_r = g.stride_r;
        f.stride_c = g.stride_c;
        f.keys = e->d_keys;
      }
    }
    int head = -1;
    rc = run_network(e, nb, &head);
    const bool fused = e->fuse_op >= 0;
    e->fuse_op = -1;
    if (rc) return rc;
    if (head < 0) return fail(VSB_ERR_INVALID, "plan has no HEAD op");
    if (fused) continue;
    const vsb_op& hop = e->ops[head];
    vsb::HeadArgs h{};
    h.logits = (const float*)e->tens[hop.src[0]].ptr;
    h.C = e->num_classes;
    h.factor = hop.factor > 0 ? hop.factor : 1;
    h.s2d = hop.mode == 1;
    h.nb = nb;
    h.g = g;
    h.d = d;
    h.s0 = s0;
    h.keys = e->vote_mode ? nullptr : e->d_keys;
    h.votes = e->vote_mode ? e->d_votes : nullptr;
    h.nvox = e->Z * e->Y * e->X;
    {
      ProfScope ps(e, PC_HEAD);
      vsb::launch_head(h, e->stream);
      CK(cudaGetLastError());
    }
  }
  return VSB_OK;
}

}  // namespace

// ============================== C ABI ========================================
extern "C" {

int vsb_abi_version(void) { return VSB_ABI_VERSION; }
int vsb_act_dtype(void) { return VSB_ACT_F16 ? 1 : 0; }
const char* vsb_last_error(void) { return g_err.c_str(); }

int vsb_direction_geometry(int64_t Z, int64_t Y, int64_t X, int32_t d, vsb_direction* out) {
  if (!out) return fail(VSB_ERR_INVALID, "null out");
  return direction_geometry(Z, Y, X, d, out);
}

int vsb_create(int device, vsb_engine** out) {
  if (!out) return fail(VSB_ERR_INVALID, "null out");
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess || n == 0)
    return fail(VSB_ERR_CUDA, "no CUDA device available (%s); libvsb200 has no CPU fallback",
                cudaGetErrorString(err));
  if (device < 0 || device >= n) return fail(VSB_ERR_INVALID, "device %d out of range (%d devices)", device, n);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(VSB_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  vsb_engine* e = new vsb_engine();
  e->device = device;
  e->num_sms = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
  e->stream = e->own_stream;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) return fail(VSB_ERR_CUDA, "cuTensorMapEncodeTiled not found");
  e->encode = (EncodeTiledFn)fn;
  CK(vsb::conv_tc_configure());
  CK(vsb::conv_halo_configure());
  CK(vsb::conv_stem_configure());
  {
    // the slicer normalises with one FMA per pixel; it must reproduce the reference formula for all 256 inputs
    const int bad = vsb::slicer_norm_selfcheck(e->stream);
    if (bad != 0) {
      cudaStreamDestroy(e->own_stream);
      delete e;
      return fail(VSB_ERR_UNSUPPORTED, "slicer normalisation self-check failed (%d of 256 byte values differ)", bad);
    }
  }
  e->no_halo = getenv("VSB_NO_HALO") != nullptr;
  if (getenv("VSB_FUSE_HEAD")) e->no_fuse_head = false;
  if (const char* sb = getenv("VSB_SUB_BATCH_MB")) e->sub_batch_mb = atoi(sb);
  *out = e;
  return VSB_OK;
}

static void close_peers(vsb_engine* e);

static void free_plan(vsb_engine* e) {
  for (ConvPlan& cp : e->conv) {
    cudaFree(cp.d_runs);
    cudaFree(cp.d_wpacked);
    cudaFree(cp.d_whalo);
    cudaFree(cp.d_wdw);
    cudaFree(cp.d_wel);
    cudaFree(cp.d_bias_el);
    cudaFree(cp.d_el_slabs);
    cudaFree(cp.d_el_entries);
    cudaFree(cp.d_el_maps);
    cudaFree(cp.d_bias_pad);
    cudaFree(cp.d_maps);
  }
  e->conv.clear();
  cudaFree(e->d_weights);
  e->d_weights = nullptr;
  e->has_plan = false;
  // the vote volume is sized by the class count of the plan that was loaded when it was allocated
  cudaFree(e->d_votes);
  e->d_votes = nullptr;
  e->votes_bytes = 0;
  e->vote_mode = 0;
}

void vsb_destroy(vsb_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  free_workspace(e);
  cudaFree(e->arena);
  e->arena = nullptr;
  e->arena_bytes = 0;
  free_plan(e);
  close_peers(e);
  cudaFree(e->d_vol_owned);
  cudaFree(e->d_gap_scratch);
  cudaFree(e->d_raw);
  cudaFree(e->d_keys_owned);
  cudaFree(e->d_votes);
  cudaFree(e->d_labels);
  cudaFree(e->d_probs);
  for (auto& ev : e->ev_pool) {
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  cudaStreamDestroy(e->own_stream);
  delete e;
}

int vsb_load_plan(vsb_engine* e, const vsb_tensor_desc* tensors, int32_t n_tensors, const vsb_op* ops,
                  int32_t n_ops, const void* weights, size_t weight_bytes, int32_t num_classes) {
  if (!e || !tensors || !ops || !weights || n_tensors < 1 || n_ops < 1)
    return fail(VSB_ERR_INVALID, "bad plan arguments");
  if (num_classes < 1 || num_classes > 32) return fail(VSB_ERR_UNSUPPORTED, "num_classes %d not in [1,32]", num_classes);
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  free_workspace(e);
  free_plan(e);
  e->tdesc.assign(tensors, tensors + n_tensors);
  e->ops.assign(ops, ops + n_ops);
  e->num_classes = num_classes;
  for (int i = 0; i < n_ops; ++i) {
    const vsb_op& op = e->ops[i];
    if (op.n_src < 0 || op.n_src > VSB_MAX_SRC) return fail(VSB_ERR_INVALID, "op %d: n_src", i);
    for (int s = 0; s < op.n_src; ++s)
      if (op.src[s] < 0 || op.src[s] >= n_tensors) return fail(VSB_ERR_INVALID, "op %d: src id", i);
    if (op.kind != VSB_OP_HEAD && (op.out <= 0 || op.out >= n_tensors)) return fail(VSB_ERR_INVALID, "op %d: out id", i);
    if (op.res >= n_tensors) return fail(VSB_ERR_INVALID, "op %d: res id", i);
    if (op.kind == VSB_OP_HEAD && op.mode == 1) {
      const vsb_tensor_desc& lt = e->tdesc[op.src[0]];
      if (num_classes > 8 || op.factor > 1 || lt.dtype != 1 || lt.ds_log2 != 1 || lt.channels != 4 * num_classes)
        return fail(VSB_ERR_INVALID, "op %d: space-to-depth head needs f32 logits [Hp/2, Wp/2, 4*classes], classes <= 8", i);
    }
    if (op.kind == VSB_OP_CONV) {
      if (op.groups < 1 || op.cin % op.groups || op.cout % op.groups) return fail(VSB_ERR_INVALID, "op %d: groups", i);
      const size_t wb = (size_t)op.cout * op.kh * op.kw * (op.cin / op.groups) * 2;
      if (op.w_off < 0 || (size_t)op.w_off + wb > weight_bytes) return fail(VSB_ERR_INVALID, "op %d: weight range", i);
      if (op.b_off >= 0 && (size_t)op.b_off + (size_t)op.cout * 4 > weight_bytes)
        return fail(VSB_ERR_INVALID, "op %d: bias range", i);
    }
  }
  e->h_weights.assign((const uint8_t*)weights, (const uint8_t*)weights + weight_bytes);
  e->weight_bytes = weight_bytes;
  CK(cudaMalloc(&e->d_weights, weight_bytes));
  CK(cudaMemcpy(e->d_weights, weights, weight_bytes, cudaMemcpyHostToDevice));
  e->conv.assign(n_ops, ConvPlan());
  for (int i = 0; i < n_ops; ++i) {
    int rc = prepare_conv_plan(e, i);
    if (rc) return rc;
    rc = prepare_el_plan(e, i);
    if (rc) return rc;
  }
  e->has_plan = true;
  return VSB_OK;
}

static const int kDtypeBytes[9] = {4, 8, 1, 1, 2, 2, 4, 4, 8};

// (Re)allocate everything that depends on the voxel count; zero the keys.  The volume buffer itself is
// (re)allocated by the callers that own it.
static int resize_voxel_state(vsb_engine* e, int64_t Z, int64_t Y, int64_t X) {
  const int64_t n = Z * Y * X;
  const bool resize = n != e->Z * e->Y * e->X || !e->d_keys_owned;
  if (resize) {
    // the other ranks map the key volume that is freed here: the exchange must be re-opened
    // (vsb_keys_ipc_export / vsb_peers_open / vsb_peers_attach) after every change of the volume size
    close_peers(e);
    cudaFree(e->d_keys_owned);
    e->d_keys_owned = nullptr;
    cudaFree(e->d_votes);
    e->d_votes = nullptr;
    e->votes_bytes = 0;
    cudaFree(e->d_labels);
    e->d_labels = nullptr;
    cudaFree(e->d_probs);
    e->d_probs = nullptr;
    CK(cudaMalloc(&e->d_keys_owned, n * 8));
    e->d_keys = e->d_keys_owned;
  }
  e->Z = Z; e->Y = Y; e->X = X;
  CK(cudaMemsetAsync(e->d_keys, 0, n * 8, e->stream));
  if (e->d_votes) CK(cudaMemsetAsync(e->d_votes, 0, e->votes_bytes, e->stream));
  return VSB_OK;
}

static int ensure_owned_volume(vsb_engine* e, size_t bytes) {
  if (!e->d_vol_owned || e->vol_owned_bytes < bytes) {
    cudaFree(e->d_vol_owned);
    e->d_vol_owned = nullptr;
    e->vol_owned_bytes = 0;
    CK(cudaMalloc(&e->d_vol_owned, align_up(bytes, 256)));
    e->vol_owned_bytes = align_up(bytes, 256);
  }
  return VSB_OK;
}

// Common body of vsb_set_volume / vsb_set_volume_typed / vsb_set_volume_shard: elements
// [v_begin, v_end) of a host volume are uploaded (the whole volume for the first two).
static int set_volume_impl(vsb_engine* e, const void* vol, int dtype, int on_device, int64_t Z, int64_t Y, int64_t X,
                           int64_t v_begin, int64_t v_end) {
  if (!e || !vol || Z <= 0 || Y <= 0 || X <= 0) return fail(VSB_ERR_INVALID, "bad volume arguments");
  if (dtype < 0 || dtype > 8 || !vsb::slicer_typed_supported(dtype))
    return fail(VSB_ERR_UNSUPPORTED,
                "volume dtype code %d is not sliceable (supported: float32, uint8, int8, uint16, int16, int32)", dtype);
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  const int64_t n = Z * Y * X;
  const size_t esz = (size_t)kDtypeBytes[dtype];
  if (on_device) {
    cudaFree(e->d_vol_owned);
    e->d_vol_owned = nullptr;
    e->vol_owned_bytes = 0;
    e->d_vol = (const uint8_t*)vol;
  } else {
    int rc = ensure_owned_volume(e, (size_t)n * esz);
    if (rc) return rc;
    if (v_end > v_begin)
      CK(cudaMemcpyAsync(e->d_vol_owned + (size_t)v_begin * esz, (const uint8_t*)vol + (size_t)v_begin * esz,
                         (size_t)(v_end - v_begin) * esz, cudaMemcpyHostToDevice, e->stream));
    e->d_vol = e->d_vol_owned;
  }
  e->vol_dtype = dtype;
  e->vol_generation += 1;
  return resize_voxel_state(e, Z, Y, X);
}

int vsb_set_volume(vsb_engine* e, const uint8_t* vol, int32_t on_device, int64_t Z, int64_t Y, int64_t X) {
  return set_volume_impl(e, vol, 2, on_device, Z, Y, X, 0, Z * Y * X);
}

int vsb_set_volume_typed(vsb_engine* e, const void* vol, int32_t dtype, int32_t on_device, int64_t Z, int64_t Y,
                         int64_t X) {
  return set_volume_impl(e, vol, dtype, on_device, Z, Y, X, 0, Z * Y * X);
}

int vsb_set_volume_shard(vsb_engine* e, const uint8_t* vol_host, int64_t Z, int64_t Y, int64_t X, int64_t v_begin,
                         int64_t v_end) {
  if (v_begin < 0 || v_end > Z * Y * X || v_begin > v_end) return fail(VSB_ERR_INVALID, "bad voxel range");
  return set_volume_impl(e, vol_host, 2, 0, Z, Y, X, v_begin, v_end);
}

int vsb_volume_pull(vsb_engine* e, vsb_engine* peer, int64_t v_begin, int64_t v_end) {
  if (!e || !peer || !e->d_vol_owned || !peer->d_vol) return fail(VSB_ERR_STATE, "volume_pull: no volume set");
  const int64_t n = e->Z * e->Y * e->X;
  if (n != peer->Z * peer->Y * peer->X || e->vol_dtype != 2 || peer->vol_dtype != 2)
    return fail(VSB_ERR_INVALID, "volume_pull: the two engines hold different volumes");
  if (v_begin < 0 || v_end > n || v_begin > v_end) return fail(VSB_ERR_INVALID, "bad voxel range");
  if (v_begin == v_end || e == peer) return VSB_OK;
  CK(cudaSetDevice(e->device));
  CK(cudaMemcpyPeerAsync(e->d_vol_owned + v_begin, e->device, peer->d_vol + v_begin, peer->device,
                         (size_t)(v_end - v_begin), e->stream));
  return VSB_OK;
}

int vsb_reset_keys(vsb_engine* e) {
  if (!e || !e->d_keys) return fail(VSB_ERR_STATE, "no volume set");
  CK(cudaSetDevice(e->device));
  const int64_t n = e->Z * e->Y * e->X;
  CK(cudaMemsetAsync(e->d_keys, 0, n * 8, e->stream));
  if (e->d_votes) CK(cudaMemsetAsync(e->d_votes, 0, e->votes_bytes, e->stream));
  return VSB_OK;
}

int vsb_predict_range(vsb_engine* e, int32_t d, int64_t s_begin, int64_t s_end) {
  if (!e) return fail(VSB_ERR_INVALID, "null engine");
  CK(cudaSetDevice(e->device));
  return predict_range(e, d, s_begin, s_end);
}

int vsb_predict(vsb_engine* e, uint32_t dir_mask, int32_t skip_duplicates) {
  if (!e) return fail(VSB_ERR_INVALID, "null engine");
  CK(cudaSetDevice(e->device));
  if (dir_mask == 0 || dir_mask >= (1u << 12)) return fail(VSB_ERR_INVALID, "bad direction mask");
  for (int d = 0; d < 12; ++d) {
    if (!(dir_mask & (1u << d))) continue;
    if (skip_duplicates && !e->vote_mode && (d == 3 || d == 6 || d == 9 || d == 10)) {
      // identical image sets to d = 1, 4, 7, 0; only skipped
      // when the earlier twin is part of the same request
      const int twin = d == 3 ? 1 : (d == 6 ? 4 : (d == 9 ? 7 : 0));
      if (dir_mask & (1u << twin)) continue;
    }
    vsb_direction g;
    int rc = direction_geometry(e->Z, e->Y, e->X, d, &g);
    if (rc) return rc;
    rc = predict_range(e, d, 0, g.S);
    if (rc) return rc;
  }
  return VSB_OK;
}

int vsb_keys(vsb_engine* e, void** dev_ptr, int64_t* count) {
  if (!e || !e->d_keys) return fail(VSB_ERR_STATE, "no volume set");
  if (dev_ptr) *dev_ptr = e->d_keys;
  if (count) *count = e->Z * e->Y * e->X;
  return VSB_OK;
}

int vsb_bind_keys(vsb_engine* e, void* dev_ptr) {
  if (!e || !e->d_keys_owned) return fail(VSB_ERR_STATE, "no volume set");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  e->d_keys = dev_ptr ? (unsigned long long*)dev_ptr : e->d_keys_owned;
  return VSB_OK;
}

int vsb_unpack_device(vsb_engine* e, uint8_t* labels_dev, uint16_t* probs_dev) {
  if (!e || !e->d_keys) return fail(VSB_ERR_STATE, "no volume set");
  CK(cudaSetDevice(e->device));
  e->launches += 1;
  vsb::launch_unpack(e->d_keys, e->Z * e->Y * e->X, labels_dev, probs_dev, e->stream);
  CK(cudaGetLastError());
  return VSB_OK;
}

int vsb_fetch(vsb_engine* e, uint8_t* labels, uint16_t* probs) {
  if (!e || !e->d_keys) return fail(VSB_ERR_STATE, "no volume set");
  if (!labels) return fail(VSB_ERR_INVALID, "null labels");
  CK(cudaSetDevice(e->device));
  const int64_t n = e->Z * e->Y * e->X;
  if (!e->d_labels) CK(cudaMalloc(&e->d_labels, align_up(n, 256)));
  if (probs && !e->d_probs) CK(cudaMalloc(&e->d_probs, align_up(n * 2, 256)));
  e->launches += 1;
  vsb::launch_unpack(e->d_keys, n, e->d_labels, probs ? e->d_probs : nullptr, e->stream);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(labels, e->d_labels, n, cudaMemcpyDeviceToHost, e->stream));
  if (probs) CK(cudaMemcpyAsync(probs, e->d_probs, n * 2, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  prof_collect(e);
  return VSB_OK;
}

int vsb_set_vote_mode(vsb_engine* e, int32_t on) {
  if (!e) return fail(VSB_ERR_INVALID, "null engine");
  if (on && !e->has_plan) return fail(VSB_ERR_STATE, "no plan loaded");
  if (on && !e->d_keys) return fail(VSB_ERR_STATE, "no volume set");
  CK(cudaSetDevice(e->device));
  e->vote_mode = on ? 1 : 0;
  const size_t need = align_up((size_t)e->Z * e->Y * e->X * e->num_classes, 4);
  if (on && (!e->d_votes || e->votes_bytes < need)) {
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(e->d_votes);
    e->d_votes = nullptr;
    e->votes_bytes = 0;
    CK(cudaMalloc(&e->d_votes, need));
    e->votes_bytes = need;
    CK(cudaMemsetAsync(e->d_votes, 0, need, e->stream));
  }
  return VSB_OK;
}

int vsb_fetch_votes(vsb_engine* e, uint8_t* votes) {
  if (!e || !e->d_votes || !votes) return fail(VSB_ERR_STATE, "vote mode not active");
  CK(cudaSetDevice(e->device));
  CK(cudaMemcpyAsync(votes, e->d_votes, (size_t)e->Z * e->Y * e->X * e->num_classes, cudaMemcpyDeviceToHost,
                     e->stream));
  CK(cudaStreamSynchronize(e->stream));
  prof_collect(e);
  return VSB_OK;
}

int vsb_synchronize(vsb_engine* e) {
  if (!e) return fail(VSB_ERR_INVALID, "null engine");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  prof_collect(e);
  return VSB_OK;
}

int vsb_set_stream(vsb_engine* e, void* stream) {
  if (!e) return fail(VSB_ERR_INVALID, "null engine");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  prof_collect(e);
  e->stream = stream ? (cudaStream_t)stream : e->own_stream;
  return VSB_OK;
}

int vsb_launch_count(vsb_engine* e, int64_t* count, int32_t reset) {
  if (!e || !count) return fail(VSB_ERR_INVALID, "null argument");
  *count = e->launches;
  if (reset) e->launches = 0;
  return VSB_OK;
}

int vsb_set_batch(vsb_engine* e, int32_t n) {
  if (!e || n < 0 || n > 1024) return fail(VSB_ERR_INVALID, "bad batch");
  e->batch_override = n;
  return VSB_OK;
}

int vsb_set_flag(vsb_engine* e, const char* name, int32_t value) {
  if (!e || !name) return fail(VSB_ERR_INVALID, "null argument");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  const std::string n(name);
  if (n == "halo") e->no_halo = value == 0;
  else if (n == "fuse_head") e->no_fuse_head = value == 0;
  else if (n == "sub_batch_mb") e->sub_batch_mb = value;
  else if (n == "halo_dbg") e->halo_dbg = value;
  else if (n == "halo_prof") {
    if (value && !e->d_halo_prof) CK(cudaMalloc(&e->d_halo_prof, 32 * 8));
    if (!value && e->d_halo_prof) { cudaFree(e->d_halo_prof); e->d_halo_prof = nullptr; }
  }
  else if (n == "sync_each") e->sync_each = value != 0;
  else if (n == "fuse_pool") { e->no_fuse_pool = value == 0; free_workspace(e); }
  else if (n == "stem") { e->stem_version = value; free_workspace(e); }
  else if (n == "stem_dbg") e->stem_dbg = value;
  else if (n == "s2d_up") e->no_s2d_up = value == 0;
  else if (n == "el_conv") e->no_el_conv = value == 0;
  else if (n == "dw_tc") e->no_dw_tc = value == 0;
  else if (n == "res_inplace") { e->no_res_inplace = value == 0; free_workspace(e); }
  else if (n == "el_tma_epilogue") { e->no_el_tma_epilogue = value == 0; free_workspace(e); }
  else if (n == "el_a_stages") { e->el_a_stages = std::max(2, std::min(value, 4)); free_workspace(e); }
  else if (n == "row_batch_mpx") e->row_batch_px = (int64_t)value << 20;
  else if (n == "dw_tiled") e->no_dw_tiled = value == 0;
  else if (n == "tc_smem_epilogue") { e->no_tc_smem_epilogue = value == 0; free_workspace(e); }
  else if (n == "halo_ab") { e->halo_ab_override = value; free_workspace(e); }
  else if (n == "halo_mt_bn") { e->halo_mt_max_bn = value; free_workspace(e); }
  else if (n == "halo_mt") { e->no_halo_mt = value < 2; free_workspace(e); }
  else if (n == "halo2_tma") { e->no_halo2_tma = value == 0; free_workspace(e); }
  else if (n == "halo2_mma2") { e->halo2_mma2 = value != 0; free_workspace(e); }
  else if (n == "mma_warps") { e->no_mma2 = value < 2; free_workspace(e); }
  else if (n == "epi_groups") { e->no_epi_groups = value == 0; free_workspace(e); }
  else if (n == "tma_epilogue") { e->no_tma_epilogue = value == 0; free_workspace(e); }
  else if (n == "halo_a_stages") { e->halo_a_stages_max = std::max(2, std::min(value, (int)vsb::HALO_MAX_A_STAGES)); free_workspace(e); }
  else return fail(VSB_ERR_INVALID, "unknown flag '%s'", name);
  return VSB_OK;
}

int vsb_set_conv_impl(vsb_engine* e, int32_t impl) {
  if (!e || impl < 0 || impl > 2) return fail(VSB_ERR_INVALID, "bad impl");
  e->conv_impl = impl;
  return VSB_OK;
}

static int slice_batch_impl(vsb_engine* e, int32_t d, int64_t s0, int32_t nb, uint16_t* out, bool typed_path);
int vsb_slice_batch(vsb_engine* e, int32_t d, int64_t s0, int32_t nb, uint16_t* out) {
  return slice_batch_impl(e, d, s0, nb, out, false);
}
int vsb_slice_batch_generic(vsb_engine* e, int32_t d, int64_t s0, int32_t nb, uint16_t* out) {
  return slice_batch_impl(e, d, s0, nb, out, true);
}
static int slice_batch_impl(vsb_engine* e, int32_t d, int64_t s0, int32_t nb, uint16_t* out, bool typed_path) {
  if (!e || !e->d_vol || !out || nb < 1) return fail(VSB_ERR_STATE, "no volume set / bad args");
  CK(cudaSetDevice(e->device));
  vsb_direction g;
  int rc = direction_geometry(e->Z, e->Y, e->X, d, &g);
  if (rc) return rc;
  if (s0 < 0 || s0 + nb > g.S) return fail(VSB_ERR_INVALID, "bad slice range");
  uint16_t* dbuf = nullptr;
  const size_t bytes = (size_t)nb * g.Hp * g.Wp * 2;
  CK(cudaMalloc(&dbuf, bytes));
  if (e->vol_dtype == 2 && !typed_path) vsb::launch_slicer(e->d_vol, g, s0, nb, dbuf, e->stream);
  else vsb::launch_slicer_typed(e->d_vol, e->vol_dtype, g, s0, nb, dbuf, e->stream);
  cudaError_t err = cudaGetLastError();
  if (err == cudaSuccess) err = cudaMemcpyAsync(out, dbuf, bytes, cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  cudaFree(dbuf);
  CK(err);
  return VSB_OK;
}

int vsb_merge_injected(vsb_engine* e, int32_t d, const float* probs, const uint8_t* labels) {
  if (!e || !e->d_keys || !probs || !labels) return fail(VSB_ERR_STATE, "no volume set / bad args");
  CK(cudaSetDevice(e->device));
  vsb_direction g;
  int rc = direction_geometry(e->Z, e->Y, e->X, d, &g);
  if (rc) return rc;
  const int64_t n = g.S * g.H * g.W;
  float* dp = nullptr;
  uint8_t* dl = nullptr;
  CK(cudaMalloc(&dp, n * 4));
  cudaError_t err = cudaMalloc(&dl, align_up(n, 256));
  if (err == cudaSuccess) err = cudaMemcpyAsync(dp, probs, n * 4, cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(dl, labels, n, cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess) {
    vsb::launch_merge_injected(dp, dl, g, d, e->d_keys, e->stream);
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  cudaFree(dp);
  cudaFree(dl);
  CK(err);
  return VSB_OK;
}

int vsb_forward_logits(vsb_engine* e, const float* images, int32_t nb, int32_t Hp, int32_t Wp, float* logits) {
  if (!e || !e->has_plan) return fail(VSB_ERR_STATE, "no plan loaded");
  if (!images || !logits || nb < 1 || Hp < 32 || Wp < 32 || (Hp & 31) || (Wp & 31))
    return fail(VSB_ERR_INVALID, "bad forward arguments (dims must be multiples of 32)");
  CK(cudaSetDevice(e->device));
  e->keep_all = true;  // debug tensors stay addressable
  int rc = build_workspace(e, Hp, Wp, nb);
  if (rc) return rc;
  const int64_t n = (int64_t)nb * Hp * Wp;
  float* dimg = nullptr;
  CK(cudaMalloc(&dimg, n * 4));
  cudaError_t err = cudaMemcpyAsync(dimg, images, n * 4, cudaMemcpyHostToDevice, e->stream);
  if (err == cudaSuccess) {
    vsb::launch_f32_to_act(dimg, (uint16_t*)e->tens[0].ptr, n, e->stream);
    err = cudaGetLastError();
  }
  int head = -1;
  if (err == cudaSuccess) {
    rc = run_network(e, nb, &head);
    if (rc) {
      cudaStreamSynchronize(e->stream);
      cudaFree(dimg);
      return rc;
    }
  }
  std::vector<float> s2d;  // space-to-depth logits are un-shuffled on the host (test hook, not the hot path)
  if (err == cudaSuccess && head >= 0) {
    const vsb_op& hop = e->ops[head];
    const TensorBuf& lt = e->tens[hop.src[0]];
    const size_t cnt = (size_t)nb * lt.H * lt.W * lt.C;
    float* dst = logits;
    if (hop.mode == 1) {
      s2d.resize(cnt);
      dst = s2d.data();
    }
    err = cudaMemcpyAsync(dst, lt.ptr, cnt * 4, cudaMemcpyDeviceToHost, e->stream);
  }
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  cudaFree(dimg);
  CK(err);
  if (!s2d.empty()) {
    const int C = e->num_classes, Hh = Hp / 2, Wh = Wp / 2;
    for (int64_t n = 0; n < nb; ++n)
      for (int i = 0; i < Hh; ++i)
        for (int j = 0; j < Wh; ++j) {
          const float* src = s2d.data() + ((n * Hh + i) * Wh + j) * 4 * C;
          for (int q = 0; q < 4; ++q)
            memcpy(logits + ((n * Hp + 2 * i + (q >> 1)) * Wp + 2 * j + (q & 1)) * C, src + q * C, (size_t)C * 4);
        }
  }
  prof_collect(e);
  if (head < 0) return fail(VSB_ERR_INVALID, "plan has no HEAD op");
  return VSB_OK;
}

int vsb_debug_tensor(vsb_engine* e, int32_t t, float* out, int64_t capacity, int64_t* shape4) {
  if (!e || e->tens.empty() || t < 0 || t >= (int)e->tens.size() || !e->tens[t].ptr)
    return fail(VSB_ERR_STATE, "tensor not available");
  CK(cudaSetDevice(e->device));
  const TensorBuf& b = e->tens[t];
  const int64_t n = (int64_t)e->ws_nb * b.H * b.W * b.C;
  if (shape4) {
    shape4[0] = e->ws_nb; shape4[1] = b.H; shape4[2] = b.W; shape4[3] = b.C;
  }
  if (!out) return VSB_OK;
  if (capacity < n) return fail(VSB_ERR_INVALID, "capacity %lld < %lld", (long long)capacity, (long long)n);
  float* d = nullptr;
  CK(cudaMalloc(&d, n * 4));
  vsb::launch_to_f32(b.ptr, b.dtype, d, n, e->stream);
  cudaError_t err = cudaGetLastError();
  if (err == cudaSuccess) err = cudaMemcpyAsync(out, d, n * 4, cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  cudaFree(d);
  CK(err);
  return VSB_OK;
}

int vsb_clip_to_uint8(vsb_engine* e, const void* data, int32_t dtype, int64_t n, double mean, double lower,
                      double upper, uint8_t* out) {
  if (!e || !data || !out || n <= 0 || dtype < 0 || dtype > 8) return fail(VSB_ERR_INVALID, "bad clip arguments");
  if (!(upper > lower)) return fail(VSB_ERR_INVALID, "clip range is empty (upper <= lower)");
  CK(cudaSetDevice(e->device));
  const int* esz = kDtypeBytes;
  // stream the volume through the GPU in chunks so a float64 1024^3 volume never needs 8 GiB at once
  const int64_t chunk = 256ll << 20;
  void* din = nullptr;
  uint8_t* dout = nullptr;
  unsigned long long* dcnt = nullptr;
  CK(cudaMalloc(&din, (size_t)std::min(chunk, n) * esz[dtype]));
  cudaError_t err = cudaMalloc(&dout, (size_t)std::min(chunk, n));
  if (err == cudaSuccess) err = cudaMalloc(&dcnt, 16);
  if (err == cudaSuccess) err = cudaMemsetAsync(dcnt, 0, 16, e->stream);
  for (int64_t off = 0; err == cudaSuccess && off < n; off += chunk) {
    const int64_t m = std::min(chunk, n - off);
    err = cudaMemcpyAsync(din, (const uint8_t*)data + off * esz[dtype], (size_t)m * esz[dtype], cudaMemcpyHostToDevice,
                          e->stream);
    if (err != cudaSuccess) break;
    e->launches += 1;
    vsb::launch_clip_count(din, dtype, m, mean, lower, upper, dout, dcnt, e->stream);
    err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaMemcpyAsync(out + off, dout, (size_t)m, cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  }
  cudaFree(din);
  cudaFree(dout);
  cudaFree(dcnt);
  CK(err);
  return VSB_OK;
}

// ---- pre-processing on the GPU: raw volume -> statistics -> clipped uint8 volume ----
int vsb_raw_upload(vsb_engine* e, const void* data, int32_t dtype, int64_t n) {
  if (!e || !data || n <= 0 || dtype < 0 || dtype > 8) return fail(VSB_ERR_INVALID, "bad raw-volume arguments");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  cudaFree(e->d_raw);
  e->d_raw = nullptr;
  e->raw_n = 0;
  CK(cudaMalloc(&e->d_raw, (size_t)n * kDtypeBytes[dtype]));
  e->raw_dtype = dtype;
  e->raw_n = n;
  CK(cudaMemcpyAsync(e->d_raw, data, (size_t)n * kDtypeBytes[dtype], cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));  // the host buffer may go away
  return VSB_OK;
}

int vsb_raw_release(vsb_engine* e) {
  if (!e) return fail(VSB_ERR_INVALID, "null engine");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  cudaFree(e->d_raw);
  e->d_raw = nullptr;
  e->raw_n = 0;
  return VSB_OK;
}

int vsb_raw_moments(vsb_engine* e, double* out4) {
  if (!e || !e->d_raw || !out4) return fail(VSB_ERR_STATE, "no raw volume uploaded");
  CK(cudaSetDevice(e->device));
  const int np = vsb::moments_partials();
  double* dpart = nullptr;
  CK(cudaMalloc(&dpart, sizeof(double) * np));
  std::vector<double> h(np);
  double sum = 0, cnt = 0, sq = 0, nans = 0, mean = 0;
  cudaError_t err = cudaSuccess;
  for (int pass = 1; pass <= 2 && err == cudaSuccess; ++pass) {
    e->launches += 1;
    vsb::launch_moments(e->d_raw, e->raw_dtype, e->raw_n, pass, mean, dpart, e->stream);
    err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaMemcpyAsync(h.data(), dpart, sizeof(double) * np, cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) break;
    // block partials are combined in block order: the same bits on every run
    double a = 0, b = 0;
    for (int i = 0; i < np; i += 2) { a += h[i]; b += h[i + 1]; }
    if (pass == 1) {
      sum = a; cnt = b;
      mean = cnt > 0 ? sum / cnt : 0.0;
      if (e->raw_dtype == 0) mean = (double)(float)mean;  // numpy returns (and subtracts) a float32 mean for float32 data
    } else {
      sq = a; nans = b;
    }
  }
  cudaFree(dpart);
  CK(err);
  double sd = cnt > 0 ? sqrt(sq / cnt) : 0.0;
  if (e->raw_dtype == 0) sd = (double)(float)sd;
  out4[0] = cnt; out4[1] = mean; out4[2] = sd; out4[3] = nans;
  return VSB_OK;
}

int vsb_raw_clip_to_volume(vsb_engine* e, double mean, double lower, double upper, int64_t Z, int64_t Y, int64_t X,
                           uint8_t* out_host, uint64_t* counts2) {
  if (!e || !e->d_raw) return fail(VSB_ERR_STATE, "no raw volume uploaded");
  if (Z <= 0 || Y <= 0 || X <= 0 || Z * Y * X != e->raw_n) return fail(VSB_ERR_INVALID, "shape does not match the raw volume");
  if (!(upper > lower)) return fail(VSB_ERR_INVALID, "clip range is empty (upper <= lower)");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  const int64_t n = e->raw_n;
  int rc = ensure_owned_volume(e, (size_t)n);
  if (rc) return rc;
  unsigned long long* dcnt = nullptr;
  CK(cudaMalloc(&dcnt, 16));
  cudaError_t err = cudaMemsetAsync(dcnt, 0, 16, e->stream);
  unsigned long long hc[2] = {0, 0};
  if (err == cudaSuccess) {
    e->launches += 1;
    vsb::launch_clip_count(e->d_raw, e->raw_dtype, n, mean, lower, upper, e->d_vol_owned, dcnt, e->stream);
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaMemcpyAsync(hc, dcnt, 16, cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess && out_host) err = cudaMemcpyAsync(out_host, e->d_vol_owned, (size_t)n, cudaMemcpyDeviceToHost, e->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
  cudaFree(dcnt);
  cudaFree(e->d_raw);
  e->d_raw = nullptr;
  e->raw_n = 0;
  CK(err);
  if (counts2) { counts2[0] = hc[0]; counts2[1] = hc[1]; }
  // the clipped volume is now the engine's resident uint8 volume (no second upload)
  e->d_vol = e->d_vol_owned;
  e->vol_dtype = 2;
  e->vol_generation += 1;
  return resize_voxel_state(e, Z, Y, X);
}

int vsb_volume_generation(vsb_engine* e, int64_t* gen) {
  if (!e || !gen) return fail(VSB_ERR_INVALID, "null argument");
  *gen = e->vol_generation;
  return VSB_OK;
}

int vsb_keys_ipc_export(vsb_engine* e, uint8_t* handle64) {
  if (!e || !handle64) return fail(VSB_ERR_INVALID, "null argument");
  if (!e->d_keys_owned || e->d_keys != e->d_keys_owned)
    return fail(VSB_ERR_STATE, "IPC export needs the engine-owned key volume (no vsb_bind_keys)");
  CK(cudaSetDevice(e->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, e->d_keys_owned));
  memcpy(handle64, &h, 64);
  return VSB_OK;
}

static void close_peers(vsb_engine* e) {
  for (int r = 0; r < 8; ++r) {
    if (e->peer_keys[r] && r != e->my_rank && e->peers_ipc) cudaIpcCloseMemHandle(e->peer_keys[r]);
    e->peer_keys[r] = nullptr;
  }
  e->peers_ipc = false;
  e->n_ranks = 1;
  e->my_rank = 0;
}

int vsb_peers_open(vsb_engine* e, int32_t n_ranks, int32_t my_rank, const uint8_t* handles64) {
  if (!e || !handles64 || n_ranks < 1 || n_ranks > 8 || my_rank < 0 || my_rank >= n_ranks)
    return fail(VSB_ERR_INVALID, "bad peer arguments (1..8 ranks)");
  if (!e->d_keys_owned) return fail(VSB_ERR_STATE, "no volume set");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  close_peers(e);
  e->n_ranks = n_ranks;
  e->my_rank = my_rank;
  e->peers_ipc = true;
  for (int r = 0; r < n_ranks; ++r) {
    if (r == my_rank) {
      e->peer_keys[r] = e->d_keys_owned;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles64 + (size_t)r * 64, 64);
    void* p = nullptr;
    cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) {
      close_peers(e);
      return fail(VSB_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(err));
    }
    e->peer_keys[r] = (unsigned long long*)p;
  }
  return VSB_OK;
}

// Same-process variant (one host process driving several GPUs, e.g. VolSeg2dPredictor with the
// additive `cuda_devices` setting): the peers are engine handles of this process, their key volumes are
// read through ordinary peer access.
int vsb_peers_attach(vsb_engine* e, int32_t n_ranks, int32_t my_rank, vsb_engine* const* engines) {
  if (!e || !engines || n_ranks < 1 || n_ranks > 8 || my_rank < 0 || my_rank >= n_ranks || engines[my_rank] != e)
    return fail(VSB_ERR_INVALID, "bad peer arguments (1..8 engines, engines[my_rank] must be this engine)");
  if (!e->d_keys_owned || e->d_keys != e->d_keys_owned) return fail(VSB_ERR_STATE, "no volume set / foreign key buffer bound");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  close_peers(e);
  const int64_t n = e->Z * e->Y * e->X;
  for (int r = 0; r < n_ranks; ++r) {
    vsb_engine* q = engines[r];
    if (!q || !q->d_keys_owned || q->Z * q->Y * q->X != n) return fail(VSB_ERR_STATE, "peer %d holds no volume of this size", r);
    if (q->device != e->device) {
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, e->device, q->device));
      if (!can) return fail(VSB_ERR_UNSUPPORTED, "GPU %d cannot access GPU %d", e->device, q->device);
      const cudaError_t err = cudaDeviceEnablePeerAccess(q->device, 0);
      if (err == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (err != cudaSuccess)
        return fail(VSB_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", e->device, q->device, cudaGetErrorString(err));
    }
  }
  e->n_ranks = n_ranks;
  e->my_rank = my_rank;
  e->peers_ipc = false;
  for (int r = 0; r < n_ranks; ++r) e->peer_keys[r] = engines[r]->d_keys_owned;
  return VSB_OK;
}

// reduce + unpack this rank's voxel shard (over the attached / opened peers, or alone) and copy it to
// the host: labels_host / probs_host point at the SHARD's first element.  Synchronises the stream.
int vsb_fetch_shard(vsb_engine* e, int64_t v_begin, int64_t v_end, uint8_t* labels_host, uint16_t* probs_host) {
  if (!e || !e->d_keys_owned || !labels_host) return fail(VSB_ERR_STATE, "no volume set / null output");
  const int64_t n = e->Z * e->Y * e->X, m = v_end - v_begin;
  if (v_begin < 0 || v_end > n || m < 0 || (v_begin & 1)) return fail(VSB_ERR_INVALID, "bad shard (begin must be even)");
  if (m == 0) return VSB_OK;
  CK(cudaSetDevice(e->device));
  if (!e->d_labels) CK(cudaMalloc(&e->d_labels, align_up(n, 256)));
  if (probs_host && !e->d_probs) CK(cudaMalloc(&e->d_probs, align_up(n * 2, 256)));
  int rc = vsb_reduce_unpack_shard(e, v_begin, v_end, e->d_labels, probs_host ? e->d_probs : nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(labels_host, e->d_labels, (size_t)m, cudaMemcpyDeviceToHost, e->stream));
  if (probs_host) CK(cudaMemcpyAsync(probs_host, e->d_probs, (size_t)m * 2, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  prof_collect(e);
  return VSB_OK;
}

int vsb_peers_close(vsb_engine* e) {
  if (!e) return fail(VSB_ERR_INVALID, "null engine");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  close_peers(e);
  return VSB_OK;
}

int vsb_reduce_unpack_shard(vsb_engine* e, int64_t v_begin, int64_t v_end, uint8_t* labels_dev, uint16_t* probs_dev) {
  if (!e || !e->d_keys_owned || !labels_dev) return fail(VSB_ERR_STATE, "no volume set / null output");
  const int64_t n = e->Z * e->Y * e->X;
  if (v_begin < 0 || v_end > n || v_begin > v_end || (v_begin & 1))
    return fail(VSB_ERR_INVALID, "bad shard [%lld, %lld) (begin must be even)", (long long)v_begin, (long long)v_end);
  for (int r = 0; r < e->n_ranks; ++r)
    if (!e->peer_keys[r] && !(e->n_ranks == 1)) return fail(VSB_ERR_STATE, "peer %d not opened", r);
  if (e->n_ranks > 1 && e->peer_keys[e->my_rank] != e->d_keys_owned)
    return fail(VSB_ERR_STATE, "peer table is stale (the key volume was re-allocated): re-open the exchange");
  CK(cudaSetDevice(e->device));
  const unsigned long long* ks[8];
  if (e->n_ranks == 1) ks[0] = e->d_keys;
  else for (int r = 0; r < e->n_ranks; ++r) ks[r] = e->peer_keys[r];
  e->launches += 1;
  vsb::launch_reduce_unpack(ks, e->n_ranks, v_begin, v_end - v_begin, labels_dev, probs_dev, e->stream);
  CK(cudaGetLastError());
  return VSB_OK;
}

int vsb_set_profiling(vsb_engine* e, int32_t on) {
  if (!e) return fail(VSB_ERR_INVALID, "null engine");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  prof_collect(e);
  e->profiling = on != 0;
  for (int i = 0; i < PC_N; ++i) {
    e->prof_ms[i] = 0.f;
    e->prof_launches[i] = 0;
  }
  e->op_ms.clear();
  e->op_launches.clear();
  return VSB_OK;
}

int vsb_op_ms(vsb_engine* e, int32_t op, float* ms, int64_t* launches) {
  if (!e || op < 0) return fail(VSB_ERR_INVALID, "bad op");
  prof_collect(e);
  const bool have = op < (int)e->op_ms.size();
  if (ms) *ms = have ? e->op_ms[op] : 0.f;
  if (launches) *launches = have ? e->op_launches[op] : 0;
  return VSB_OK;
}

int vsb_stage_ms(vsb_engine* e, int32_t stage, float* ms, int64_t* launches) {
  if (!e || stage < 0 || stage >= PC_N) return fail(VSB_ERR_INVALID, "bad stage");
  prof_collect(e);
  if (ms) *ms = e->prof_ms[stage];
  if (launches) *launches = e->prof_launches[stage];
  return VSB_OK;
}

}  // extern "C"
