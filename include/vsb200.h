/* vsb200.h -- C ABI of libvsb200.so, the B200 (sm_100a) engine that replaces the
 * prediction hot path of DiamondLightSource/volume-segmantics.
 *
 * The reference has no FFI of its own: its boundary is the Python class
 * VolSeg2dPredictor (volume_segmantics/model/operations/vol_seg_2d_predictor.py).
 * Each entry point below names the reference statements it replaces
 * (file:line relative to the reference repository).  All pointers are
 * caller-owned, all sizes explicit, no C++ or torch types cross this line.
 * Every function returns VSB_OK (0) or a negative error code; the message is
 * available from vsb_last_error().  Nothing here ever aborts the process and
 * nothing here has a CPU fallback: without a CUDA device vsb_create() fails.
 *
 * Threading: a handle is not thread-safe; one handle per process per GPU.
 */
#ifndef VSB200_H_
#define VSB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSB_ABI_VERSION 2

#define VSB_OK 0
#define VSB_ERR_INVALID -1   /* bad argument / plan */
#define VSB_ERR_CUDA -2      /* CUDA runtime / driver error */
#define VSB_ERR_STATE -3     /* call order (no plan, no volume, ...) */
#define VSB_ERR_UNSUPPORTED -4

typedef struct vsb_engine vsb_engine;

/* ---- network plan ------------------------------------------------------
 * The host (Python) folds BatchNorm into the convolutions and lowers the
 * smp network (model_2d.py:10-39) into a flat op list over numbered tensors.
 * Tensor 0 is the network input: [nb, Hp, Wp, 1] 16-bit (see vsb_act_dtype), written by the slicer.
 * Spatial size of tensor t is (Hp >> ds_log2, Wp >> ds_log2); ds_log2 == -1
 * means 1x1 (global pooled).  Layout of every tensor is NHWC.               */
typedef struct {
  int32_t channels;
  int32_t ds_log2;
  int32_t dtype; /* 0 = 16-bit activation format of the build (vsb_act_dtype), 1 = f32 */
  int32_t reserved;
} vsb_tensor_desc;

enum {
  VSB_OP_CONV = 1,      /* conv (+folded BN bias) (+residual) (+ReLU)        */
  VSB_OP_MAXPOOL = 2,   /* 3x3 stride 2 pad 1                                 */
  VSB_OP_GAP = 3,       /* global average pool -> 1x1                         */
  VSB_OP_UPSAMPLE = 4,  /* mode 0: bilinear x`factor` align_corners=True;
                           mode 1: broadcast a 1x1 tensor to the out size     */
  VSB_OP_HEAD = 5       /* logits -> softmax -> argmax -> fp16 -> merge       */
};

#define VSB_MAX_SRC 6

typedef struct {
  int32_t kind;
  int32_t out;                 /* output tensor id (HEAD: unused, -1)         */
  int32_t n_src;               /* conv input = channel concat of the sources  */
  int32_t src[VSB_MAX_SRC];
  int32_t src_up[VSB_MAX_SRC]; /* 1: source is nearest-upsampled x2 first     */
  int32_t res;                 /* residual tensor id added before ReLU, or -1 */
  int32_t cin, cout;
  int32_t kh, kw, stride, pad, dil, groups;
  int32_t relu;
  int32_t mode, factor;        /* UPSAMPLE: mode/factor. HEAD: factor = bilinear
                                  upsampling of the logits (1 = none).
                                  CONV: mode 2 = the blob also holds, at byte
                                  offset factor * 256, the 16-bit weights
                                  [4*cout][3][3][C0 + 4*(cin - C0)] of the same
                                  layer as a convolution between space-to-depth
                                  tensors at half the output resolution (source
                                  0 up-sampled with C0 channels, then the
                                  (py, px, c) sub-pixels of the other sources;
                                  output channel (a*2+b)*cout + c = pixel
                                  (2i+a, 2j+b)); the engine may run either form */
  int64_t w_off;               /* byte offset in the weight blob of 16-bit
                                  [cout][kh][kw][cin/groups] (OHWI)           */
  int64_t b_off;               /* byte offset of f32 [cout] bias              */
} vsb_op;

/* Direction d = 3*k + a: k quarter turns of np.rot90 in the Z-Y plane
 * (vol_seg_2d_predictor.py:108), a = Axis.Z/Y/X swap (base_data_utils.py:132-138).
 * A slice-space pixel (s, r, c) of direction d is voxel
 *     base + s*stride_s + r*stride_r + c*stride_c      (element offsets)
 * of the C-ordered (Z,Y,X) volume.                                           */
typedef struct {
  int64_t S, H, W;          /* slices, rows, cols of the direction's images   */
  int64_t Hp, Wp;           /* padded to multiples of 32 (augmentations.py:30-44) */
  int64_t pad_top, pad_left;   /* int(p/2): albumentations PadIfNeeded centre */
  int64_t crop_top, crop_left; /* round-half-even(p/2): torchvision center_crop
                                  (base_data_utils.py:125-129)                */
  int64_t base, stride_s, stride_r, stride_c;
} vsb_direction;

int vsb_abi_version(void);
/* 16-bit format of activations and weights in this build: 0 = bfloat16,
 * 1 = IEEE half.  The weight blob handed to vsb_load_plan must use it.       */
int vsb_act_dtype(void);
const char* vsb_last_error(void);

/* VolSeg2dPredictor.__init__ (vol_seg_2d_predictor.py:19-26): bind a GPU.   */
int vsb_create(int device, vsb_engine** out);
void vsb_destroy(vsb_engine* e);

/* create_model_from_file (model_2d.py:42-57) after host-side BN folding.    */
int vsb_load_plan(vsb_engine* e, const vsb_tensor_desc* tensors, int32_t n_tensors,
                  const vsb_op* ops, int32_t n_ops, const void* weights, size_t weight_bytes,
                  int32_t num_classes);

/* Geometry of one direction; pure host arithmetic (also used by the tests). */
int vsb_direction_geometry(int64_t Z, int64_t Y, int64_t X, int32_t d, vsb_direction* out);

/* data_vol handed to _predict_* (vol_seg_2d_predictor.py:31,67,100): a uint8
 * (Z,Y,X) C-ordered volume.  on_device != 0: `vol` is a device pointer that is
 * adopted without a copy (caller keeps it alive); else it is copied H2D.
 * Allocates and zeroes the packed-key volume (8 B / voxel).                  */
int vsb_set_volume(vsb_engine* e, const uint8_t* vol, int32_t on_device, int64_t Z, int64_t Y,
                   int64_t X);

/* The same for volumes that are not uint8 (datasets.py:129-135: integer slices of ANY bit depth are
 * cast to float32 and divided by 255, float32 slices are used as they are; then (x - 0.449) / 0.226 in
 * fp32).  dtype: 0 float32, 2 uint8, 3 int8, 4 uint16, 5 int16, 7 int32 -- the types cv2.copyMakeBorder
 * (albumentations PadIfNeeded, augmentations.py:61-65) pads without conversion.  float64 / float16
 * volumes fail in the reference too (double input to a float model), so they are refused here.         */
int vsb_set_volume_typed(vsb_engine* e, const void* vol, int32_t dtype, int32_t on_device, int64_t Z, int64_t Y,
                         int64_t X);

/* Residency token: incremented whenever the engine's resident volume changes, so the host shim can tell
 * that the array it is asked to predict is the one a previous call left in HBM (vsb_raw_clip_to_volume). */
int vsb_volume_generation(vsb_engine* e, int64_t* generation);

/* Zero the key volume (start of a new prediction on the same data).         */
int vsb_reset_keys(vsb_engine* e);

/* The inner hot loop, vol_seg_2d_predictor.py:40-58 + :60-64 + :90-98, for the
 * slices [s_begin, s_end) of direction d (a multi-GPU work item):
 * slicer -> network -> softmax/argmax -> fp16 -> inverse rotation -> packed
 * key atomicMax.  Asynchronous on the engine's stream.                       */
int vsb_predict_range(vsb_engine* e, int32_t d, int64_t s_begin, int64_t s_end);

/* _predict_single_axis / _predict_3_ways_max_probs / _predict_12_ways_max_probs
 * (vol_seg_2d_predictor.py:31-116): every direction whose bit is set in
 * dir_mask, all slices.  skip_duplicates != 0 drops directions 3,6,9,10 whose
 * images duplicate an earlier direction (result-identical).      */
int vsb_predict(vsb_engine* e, uint32_t dir_mask, int32_t skip_duplicates);

/* Device pointer / element count of the uint64 key volume, so the host can
 * run the one NCCL max-reduce over it.  vsb_bind_keys lets the
 * host substitute its own device buffer (e.g. a torch tensor) of the same
 * size, which the engine then uses instead of its own.                       */
int vsb_keys(vsb_engine* e, void** dev_ptr, int64_t* count);
int vsb_bind_keys(vsb_engine* e, void* dev_ptr);

/* Fused multi-GPU exchange over NVLink peer memory: every rank exports a
 * CUDA-IPC handle of its key volume, opens the others', and then max-reduces + unpacks
 * its own voxel shard [v_begin, v_end) with ONE kernel that loads the peers' keys
 * directly (no NCCL all-reduce, no second pass over the keys).  The caller orders the
 * ranks with a stream-ordered barrier before (all ranks finished predicting) and after
 * (all ranks finished reading) the call.  handles64: n_ranks x 64 bytes, rank-major.   */
int vsb_keys_ipc_export(vsb_engine* e, uint8_t* handle64);
int vsb_peers_open(vsb_engine* e, int32_t n_ranks, int32_t my_rank, const uint8_t* handles64);
int vsb_peers_close(vsb_engine* e);
/* One host process driving several GPUs (the additive `cuda_devices` setting of the drop-in
 * VolSeg2dPredictor): the peers are engine handles of THIS process, read through ordinary
 * peer access.  engines[my_rank] must be `e`; every engine must already hold a volume of the same size.
 *   vsb_set_volume_shard  allocate the whole uint8 volume, upload only voxels [v_begin, v_end) of it
 *   vsb_volume_pull       copy voxels [v_begin, v_end) of `peer`'s volume into this engine's (NVLink
 *                         peer copy on this engine's stream): N uploads of 1/N + an all-gather over
 *                         NVLink instead of N uploads of the whole volume over PCIe
 *   vsb_fetch_shard       max-reduce + unpack voxels [v_begin, v_end) over all peers' keys and copy
 *                         them to the host (pointers address the SHARD's first element): every GPU
 *                         downloads its own part of the result over its own PCIe link
 * The caller orders the engines (all predictions finished before any vsb_fetch_shard; all shards
 * fetched before the next vsb_reset_keys), e.g. with vsb_synchronize + a host barrier.                 */
int vsb_peers_attach(vsb_engine* e, int32_t n_ranks, int32_t my_rank, vsb_engine* const* engines);
int vsb_set_volume_shard(vsb_engine* e, const uint8_t* vol_host, int64_t Z, int64_t Y, int64_t X, int64_t v_begin,
                         int64_t v_end);
int vsb_volume_pull(vsb_engine* e, vsb_engine* peer, int64_t v_begin, int64_t v_end);
int vsb_fetch_shard(vsb_engine* e, int64_t v_begin, int64_t v_end, uint8_t* labels_host, uint16_t* probs_fp16_host);
int vsb_reduce_unpack_shard(vsb_engine* e, int64_t v_begin, int64_t v_end, uint8_t* labels_dev,
                            uint16_t* probs_fp16_dev);

/* Unpack keys -> labels uint8 (Z,Y,X) [+ probs fp16 bits] and copy to host
 * (the `return labels, probs` of vol_seg_2d_predictor.py:65,88,116).
 * probs may be NULL.  Synchronises the stream.                              */
int vsb_fetch(vsb_engine* e, uint8_t* labels, uint16_t* probs_fp16);
/* Same, into device buffers (no D2H, asynchronous).                         */
int vsb_unpack_device(vsb_engine* e, uint8_t* labels_dev, uint16_t* probs_fp16_dev);

/* One-hot vote path (vol_seg_2d_predictor.py:118-136): per-class uint8 vote
 * counts (C,Z,Y,X).  vsb_predict* with vote mode on accumulates votes
 * instead of keys.                                                           */
int vsb_set_vote_mode(vsb_engine* e, int32_t on);
int vsb_fetch_votes(vsb_engine* e, uint8_t* votes);

int vsb_synchronize(vsb_engine* e);

/* Run on a caller-owned CUDA stream (cudaStream_t), e.g. torch's current stream,
 * so host-side events and NCCL collectives order with the engine's kernels.
 * NULL returns to the engine's own stream.                                   */
int vsb_set_stream(vsb_engine* e, void* cuda_stream);
/* Number of kernels this engine has launched (optionally reset).            */
int vsb_launch_count(vsb_engine* e, int64_t* count, int32_t reset);

/* Slices processed per launch (reference: 4, config.py:31). 0 = automatic.   */
int vsb_set_batch(vsb_engine* e, int32_t slices_per_batch);
/* 0: tcgen05 implicit-GEMM convolutions wherever legal (default);
 * 1: CUDA-core convolutions everywhere (bring-up / cross-check);
 * 2: as 1 but also without the dedicated 7x7 stem kernel.                   */
int vsb_set_conv_impl(vsb_engine* e, int32_t impl);
/* Tuning / cross-check switches (every alternative computes the same arithmetic; tests/test_engine_paths_gpu.py
 * checks that the pipeline switches give bit-identical logits):
 *   "halo" (1)            halo-tile conv kernels; 0: per-tap TMA kernel everywhere
 *   "tma_epilogue" (1)    epilogue through shared memory + TMA store where weights stay resident; 0: per-thread stores
 *   "tc_smem_epilogue" (1) the same for the per-tap kernel (1x1 and stride-2 convolutions); 0: per-thread stores
 *   "epi_groups" (1)      two epilogue groups on alternate tiles for BN <= 64; 0: one group
 *   "mma_warps" (2)       two MMA-issuing warps for resident-weight, one-slab launches; 1: one
 *   "halo_mt" (2)         two tiles per stage for streamed-weight launches with BN <= "halo_mt_bn" (128); 1: one
 *   "halo_a_stages" (8)   maximum depth of the halo ring, 2..8
 *   "halo2_tma" (1)       non-up-sampled sources of the cp.async halo kernel fetched by TMA; 0: by the loader warps
 *   "halo2_mma2" (0)      second MMA warp in the cp.async halo kernel (measured neutral)
 *   "fuse_pool" (1)       3x3/2 max-pool inside the stem's launch; 0: separate kernel
 *   "stem" (3)            stem + pool kernel: 3 = raw input window read in place by shifted no-swizzle descriptors, pool
 *                         computed inside the CTA; 2 = im2col rows built by loader warps, pool inside the CTA;
 *                         1 = im2col + pooling by red.global.max into a zeroed tensor (round-1 kernel)
 *   "s2d_up" (1)          decoder `upsample x2 + concat -> conv3x3` layers whose op carries mode 2 run as a
 *                         space-to-depth convolution at half resolution (summed up-sample taps: same arithmetic up
 *                         to summation order and one rounding of the summed weights); 0: parity-split kernels
 *   "el_tma_epilogue" (1) epilogue of that kernel through shared memory + TMA store (folded output map); 0: per-thread stores
 *   "el_a_stages" (2)     halo ring depth of that kernel, 2..4 (the rest of the shared memory is weight ring)
 *   "res_inplace" (1)     halo kernel, 64-channel residual layers: the residual tile lands in the output staging buffer
 *                         (frees shared memory for a fourth halo stage and the second MMA warp); 0: separate buffers
 *   "el_conv" (1)         32 -> 32 and dense stride-2 3x3 convolutions on that kernel (exact re-arrangements); 0: round-1 kernels
 *   "dw_tc" (1)           depthwise 3x3 convolutions as block-diagonal tensor-core convolutions; 0: CUDA-core kernel
 *   "fuse_head" (0)       softmax/argmax/merge inside the last conv's epilogue (bit-identical, measured slower)
 *   "sub_batch_mb" (0)    L2 budget for depth-first sub-batches, 0 = off
 * Debugging aids: "sync_each" (synchronise after every op and name the one that failed), "halo_prof" (per-launch
 * cycle accounting of the conv_halo roles on stderr), "halo_dbg" (timing experiments that skip loads / MMAs /
 * stores -- results are wrong while set).                                                                        */
int vsb_set_flag(vsb_engine* e, const char* name, int32_t value);

/* ---- test hooks (bit-exact criteria of BASELINE.json) ---------------------
 * Slicer only: padded+normalised 16-bit (vsb_act_dtype) images [nb, Hp, Wp] of direction d,
 * slices [s0, s0+nb) (datasets.py:120-142 + augmentations.py:46-65).         */
int vsb_slice_batch(vsb_engine* e, int32_t d, int64_t s0, int32_t nb, uint16_t* out_act16_host);
/* The same through the generic (typed) slicer kernels, whatever the dtype of the resident volume.     */
int vsb_slice_batch_generic(vsb_engine* e, int32_t d, int64_t s0, int32_t nb, uint16_t* out_act16_host);
/* Merge only: inject per-direction slice-space results [S,H,W] (cropped
 * dims) fp32 max-prob + uint8 label (host pointers) for direction d.         */
int vsb_merge_injected(vsb_engine* e, int32_t d, const float* probs, const uint8_t* labels);
/* Network only: images f32 [nb,Hp,Wp] (host, already padded+normalised; they
 * are rounded to the 16-bit format as the slicer would) -> logits f32 [nb,Hp,Wp,C] (host). */
int vsb_forward_logits(vsb_engine* e, const float* images, int32_t nb, int32_t Hp, int32_t Wp,
                       float* logits_out);
/* Copy tensor `t` of the last vsb_forward_logits call to host as f32 NHWC.   */
int vsb_debug_tensor(vsb_engine* e, int32_t t, float* out, int64_t capacity_elems,
                     int64_t* shape4);

/* With profiling on, every kernel launch is bracketed by CUDA events on the
 * engine's stream; vsb_stage_ms returns the summed milliseconds and launch
 * count per kernel class since profiling was switched on:
 * 0 slicer, 1 tcgen05 conv, 2 CUDA-core conv, 3 stem, 4 max-pool,
 * 5 head+merge, 6 other (pool / up-sample).                                  */
int vsb_stage_ms(vsb_engine* e, int32_t stage, float* ms, int64_t* launches);
int vsb_set_profiling(vsb_engine* e, int32_t on);
/* Same, per op index of the loaded plan (conv / pool ops).                   */
int vsb_op_ms(vsb_engine* e, int32_t op, float* ms, int64_t* launches);

/* ---- clip_to_uint8 (base_data_utils.py:243-287), SURVEY 8f-1 --------------
 * Elementwise part of the reference's pre-processing with host buffers: NaN -> mean, clip to
 * [lower, upper], rescale to 0..255, truncate to uint8 -- the same IEEE operations in the same order
 * and PRECISION as numpy: float32 arithmetic for float32 data (numpy keeps the array dtype; mean / lower
 * / upper are then float32 values), float64 for float64 and integer data (astype(float)).  Bit-exact
 * given the same statistics.  `data` and `out` are HOST pointers, streamed through the GPU in chunks;
 * dtype: 0 f32, 1 f64, 2 u8, 3 i8, 4 u16, 5 i16, 6 u32, 7 i32, 8 i64.                                   */
int vsb_clip_to_uint8(vsb_engine* e, const void* data, int32_t dtype, int64_t n, double mean,
                      double lower, double upper, uint8_t* out);


/* ---- BaseDataManager._preprocess_data on the GPU (base_data_manager.py:29-42, base_data_utils.py:243-287)
 * One upload of the raw volume, then:
 *   vsb_raw_moments          out4 = {count of non-NaN voxels, nanmean, nanstd, count of NaNs}; two-pass
 *                            float64 reduction with a fixed reduction order (bit-reproducible run to run;
 *                            agrees with numpy's pairwise summation to rounding, not bit for bit).  For
 *                            float32 data mean and std are rounded to float32 as numpy returns them.
 *   vsb_raw_clip_to_volume   NaN -> mean, clip to [lower, upper], rescale to 0..255, truncate -- in float32
 *                            arithmetic for float32 data and float64 otherwise, exactly the operations
 *                            numpy performs (bit-exact given the same mean / lower / upper).  counts2 =
 *                            {voxels > upper, voxels < lower} (the reference's two log lines).  The uint8
 *                            result BECOMES THE ENGINE'S RESIDENT VOLUME (no second upload) and is also
 *                            copied to out_host when that is not NULL.  Frees the raw volume.
 * dtype codes as vsb_clip_to_uint8.                                                                        */
int vsb_raw_upload(vsb_engine* e, const void* data_host, int32_t dtype, int64_t n);
int vsb_raw_moments(vsb_engine* e, double* out4);
int vsb_raw_clip_to_volume(vsb_engine* e, double mean, double lower, double upper, int64_t Z, int64_t Y, int64_t X,
                           uint8_t* out_host, uint64_t* counts2);
int vsb_raw_release(vsb_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* VSB200_H_ */
