// Halo tcgen05 convolution kernel (see conv_halo.cuh).  Persistent, one CTA per SM:
//   warp 0      A producer: one TMA halo box per (tile, 64-channel slab)
//   warp 1      TMEM allocator + MMA issuer: 9 taps x 4 K-steps per slab
//   warp 2      B producer: weight images, once (resident) or through a ring
//   warps 3..10 epilogue (TMEM -> registers -> bias/residual/ReLU -> NHWC)
#include "conv_halo.cuh"

#include "conv_epilogue.cuh"
#include "kernels.h"

namespace vsb {

namespace {

struct HaloCtl {
  uint64_t a_full[HALO_MAX_A_STAGES];
  uint64_t a_empty[HALO_MAX_A_STAGES];
  uint64_t b_full[HALO_MAX_B_STAGES];
  uint64_t b_empty[HALO_MAX_B_STAGES];
  uint64_t w_full;
  uint64_t acc_full[8];
  uint64_t acc_empty[8];
  uint64_t res_full[2];
  uint64_t res_empty[2];
  uint64_t out_full[2];
  uint64_t out_empty[2];
  uint32_t tmem_base;
  uint32_t pad[3];
};
// Cycle accounting slots (CTA 0, one lane per role): producer 0..3 = {total, wait a_empty, tiles, -};
// MMA 4..7 = {total, wait acc_empty, wait a_full, issue}; epilogue group 0 issuer 8..15 =
// {total, wait store-read, wait acc_full, tmem_ld+bar1, math+sts, arrive+fence+bar2, store issue, tiles}.
__device__ __forceinline__ long long dev_clock() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
  return t;
}
struct ProfClock {
  // accumulates in registers (slots are compile-time constants); flush() adds them to global
  // memory once, so the measurement does not put a global round trip into every lap
  unsigned long long* out;
  long long t;
  unsigned long long a[8];
  __device__ __forceinline__ ProfClock(unsigned long long* o) : out(o), t(o ? dev_clock() : 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0;
  }
  template <int SLOT>
  __device__ __forceinline__ void lap() {
    if (out) {
      const long long n = dev_clock();
      a[SLOT] += (unsigned long long)(n - t);
      t = n;
    }
  }
  template <int SLOT>
  __device__ __forceinline__ void count() {
    if (out) a[SLOT] += 1;
  }
  __device__ __forceinline__ void flush(int base) {
    if (out) {
#pragma unroll
      for (int i = 0; i < 8; ++i) out[base + i] += a[i];
    }
  }
};
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// elementwise max of two packed 16-bit activation pairs / 16-byte max-reduction to global
__device__ __forceinline__ uint32_t act_max2(uint32_t a, uint32_t b) {
  uint32_t r;
#if VSB_ACT_F16
  asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
#else
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
#endif
  return r;
}
__device__ __forceinline__ void red_max_act8(uint16_t* dst, uint4 v) {
#if VSB_ACT_F16
  asm volatile("red.global.max.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
#else
  asm volatile("red.global.max.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
#endif
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
constexpr int kHaloCtlBytes = 1024;
constexpr int kHaloBiasBytes = 2048 * 4;
// Stem (MODE 1): raw input window of one tile = 37 rows x 24 pixels (48 B), fetched with
// 8-byte cp.async two tiles ahead of the im2col build.
constexpr int kStemRawRows = 37, kStemRawPitch = 48, kStemRawBytes = 2048, kStemRawRing = 4;

__device__ __forceinline__ void halo_decode(const ConvHaloParams& p, int t, int& n_tile, int& X0, int& Y0,
                                            int& n) {
  auto fdiv = [](uint32_t v, const FastDiv& f) { return f.m ? __umulhi(v, f.m) : v; };
  uint32_t sp = fdiv((uint32_t)t, p.div_n_tiles);
  n_tile = t - (int)(sp * p.div_n_tiles.d);
  uint32_t q = fdiv(sp, p.div_tx);
  const int tx = (int)(sp - q * p.div_tx.d);
  sp = q;
  q = fdiv(sp, p.div_ty);
  const int ty = (int)(sp - q * p.div_ty.d);
  n = (int)q + p.n_base;
  X0 = tx * (p.mt == 2 ? 16 : 8);
  Y0 = ty * 16;
}


// Descriptor + small offset (16-byte units).  The 14-bit address field lives in the low
// word and never carries out for shared-memory addresses, so a 32-bit add suffices.
__device__ __forceinline__ uint64_t desc_add(uint64_t d, uint32_t off16) {
  return (d & 0xffffffff00000000ull) | (uint32_t)((uint32_t)d + off16);
}

// K-major operand descriptor for the compact layouts: pitch 32 -> SW32 (6), 64 -> SW64 (4),
// 128 -> SW128 (2); sbo = byte distance between consecutive 8-row groups.
template <int PITCH>
__device__ __forceinline__ uint64_t desc_compact(uint32_t addr, uint32_t sbo) {
  constexpr uint64_t layout = PITCH == 128 ? 2ull : (PITCH == 64 ? 4ull : 6ull);
  return (uint64_t)((addr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         (layout << 61);
}
// Swizzle<B,4,3> of a byte offset relative to a 1024-byte aligned base.
template <int PITCH>
__device__ __forceinline__ uint32_t swz(uint32_t off) {
  constexpr uint32_t mask = PITCH == 128 ? 7u : (PITCH == 64 ? 3u : 1u);
  return off ^ (((off >> 7) & mask) << 4);
}
// All MMAs of one 64-channel (or narrower) slab against RESIDENT weights, issued
// back-to-back by the elected lane: NTAPS taps x KSTEPS K-steps, compile-time offsets.
template <int NTAPS, int KSTEPS>
__device__ __forceinline__ void issue_slab_resident(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                    uint32_t b_step, uint32_t row_units, uint32_t col_units,
                                                    uint32_t idesc, bool accumulate_first) {
#pragma unroll
  for (int tap = 0; tap < NTAPS; ++tap) {
    const uint64_t at = desc_add(a_desc, (tap / 3) * row_units + (tap % 3) * col_units);
    const uint64_t bt = desc_add(b_desc, tap * b_step);
#pragma unroll
    for (int k = 0; k < KSTEPS; ++k)
      umma_bf16_ss(d_tmem, desc_add(at, 2 * k), desc_add(bt, 2 * k), idesc,
                   (tap | k) != 0 ? 1u : (accumulate_first ? 1u : 0u));
  }
}

// Same, skipping the K-steps whose weights are all zero (bit tap*KSTEPS+k of `mask` clear).
template <int NTAPS, int KSTEPS>
__device__ __forceinline__ void issue_slab_masked(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t b_step,
                                                  uint32_t row_units, uint32_t col_units, uint32_t idesc,
                                                  bool accumulate_first, uint64_t mask) {
  uint32_t acc = accumulate_first ? 1u : 0u;
#pragma unroll
  for (int tap = 0; tap < NTAPS; ++tap) {
    const uint64_t at = desc_add(a_desc, (tap / 3) * row_units + (tap % 3) * col_units);
    const uint64_t bt = desc_add(b_desc, tap * b_step);
#pragma unroll
    for (int k = 0; k < KSTEPS; ++k)
      if ((mask >> (tap * KSTEPS + k)) & 1ull) {
        umma_bf16_ss(d_tmem, desc_add(at, 2 * k), desc_add(bt, 2 * k), idesc, acc);
        acc = 1u;
      }
  }
}

}  // namespace

template <int KC>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ ConvHaloParams p) {
  constexpr int P = 2 * KC;        // bytes per halo pixel / weight row (128: SW128, 64: SW64)
  constexpr int KSTEPS = KC / 16;  // tcgen05.mma K-steps per slab and tap
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_area = smem + (size_t)p.a_stages * p.a_stage_bytes;
  const size_t b_area_bytes = p.b_stages ? (size_t)p.b_stages * p.b_bytes : (size_t)p.ncs * 9 * p.b_bytes;
  uint8_t* out_stage = b_area + b_area_bytes;
  // res_inplace: the residual tile is fetched INTO the staging buffer of its tile and overwritten by the output
  // (each thread reads and writes the same 16-byte chunks) -- half the staging memory, which buys the fourth
  // halo stage / second MMA warp for the 64-channel residual layers
  uint8_t* res_stage = p.res_inplace ? out_stage : out_stage + (size_t)p.out_bufs * p.out_buf_bytes;
  HaloCtl* ctl = reinterpret_cast<HaloCtl*>(out_stage + (size_t)(p.out_bufs + (p.res_inplace ? 0 : p.res_bufs)) * p.out_buf_bytes);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctl) + kHaloCtlBytes);
  float* const bias_l = bias_s;  // launch-relative index (tile-local channel)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.n_tiles * p.tiles_x * p.tiles_y * p.NB;
  const int MT = p.mt == 2 ? 2 : 1;  // 8 x 16 tiles per stage (2 only with streamed weights + direct epilogue)
  const int HW = 8 * MT + 2 * p.dil, HH = 16 + 2 * p.dil;
  const bool resident = p.b_stages == 0;
  const uint32_t acc_cols = 512u / (uint32_t)p.acc_stages;
  const int G = p.epi_groups == 2 ? 2 : 1;     // epilogue groups (alternate tiles)
  const int MW = p.mma_warps == 2 ? 2 : 1;     // MMA issuing warps (alternate tiles)
  const bool smem_epi = p.out_map != nullptr;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(&ctl->a_full[i], 1);
      mbar_init(&ctl->a_empty[i], 1);
    }
    for (int i = 0; i < HALO_MAX_B_STAGES; ++i) {
      mbar_init(&ctl->b_full[i], 1);
      mbar_init(&ctl->b_empty[i], 1);
    }
    mbar_init(&ctl->w_full, 1);
    // shared-memory epilogue: one arrival per epilogue warp; direct epilogue: one per thread
    const int epi_arrivals = smem_epi ? HALO_EPI_WARPS / G : 32 * HALO_EPI_WARPS;
    for (int i = 0; i < 8; ++i) {
      mbar_init(&ctl->acc_full[i], 1);
      mbar_init(&ctl->acc_empty[i], epi_arrivals);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->res_full[i], 1);
      mbar_init(&ctl->res_empty[i], epi_arrivals);
      mbar_init(&ctl->out_full[i], epi_arrivals);
      mbar_init(&ctl->out_empty[i], 1);
    }
    fence_mbar_init();
  }
  // bias_s is indexed by absolute output channel; this launch covers [cout_off, cout_off + n_tiles*BN)
  bias_s -= p.cout_off;
  for (int i = threadIdx.x; i < p.n_tiles * p.BN; i += HALO_THREADS) bias_s[p.cout_off + i] = p.bias[p.cout_off + i];
  if (warp == 1) tmem_alloc<512>(&ctl->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================== A producer: halo boxes (+ residual tiles) =====================
    int as = 0;
    uint32_t aph = 0;
    const uint32_t a_box_bytes = (uint32_t)(HW * HH * P);
    int rb = 0;
    uint32_t rph = 0;
    unsigned long long* pr = (p.prof && blockIdx.x == 0 && lane == 0) ? p.prof : nullptr;
    const long long pstart = pr ? dev_clock() : 0;
    ProfClock pc(pr);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n_tile, X0, Y0, n;
      halo_decode(p, t, n_tile, X0, Y0, n);
      pc.count<2>();
      if (p.res_map && !p.res_inplace) {
        // residual tile of this output tile, in the epilogue's staging layout
        mbar_wait(&ctl->res_empty[rb], rph ^ 1);
        if (elect_one()) {
          const int groups = (p.BN + 63) >> 6;
          const uint32_t gbytes = p.BN >= 64 ? 16384u : 8192u;  // 128 rows of 64 (or 32) channels
          mbar_arrive_expect_tx(&ctl->res_full[rb], groups * gbytes);
          for (int g = 0; g < groups; ++g)
            tma_load_5d(p.res_map, &ctl->res_full[rb], res_stage + (size_t)rb * p.out_buf_bytes + g * gbytes,
                        p.cout_off + n_tile * p.BN + g * 64, X0, 0, Y0, n);
        }
        __syncwarp();
        if (++rb == p.res_bufs) {
          rb = 0;
          rph ^= 1;
        }
      }
      for (int cs = 0; cs < p.ncs; ++cs) {
        pc.lap<3>();
        mbar_wait(&ctl->a_empty[as], aph ^ 1);
        pc.lap<1>();
        if (elect_one()) {
          if (p.dbg & 1) {
            mbar_arrive(&ctl->a_full[as]);
          } else {
            mbar_arrive_expect_tx(&ctl->a_full[as], a_box_bytes);
            tma_load_5d(p.map, &ctl->a_full[as], a_ring + (size_t)as * p.a_stage_bytes, p.cin_off + cs * KC, X0 - p.dil, 0,
                        Y0 - p.dil, n);
          }
        }
        __syncwarp();
        if (++as == p.a_stages) {
          as = 0;
          aph ^= 1;
        }
      }
    }
    if (pr) {
      pc.a[0] = (unsigned long long)(dev_clock() - pstart);
      pc.flush(0);
    }
  } else if (warp == 2) {
    // ===================== B producer: weight images; then TMA store issuer =====================
    if (resident) {
      if (elect_one()) {
        const uint32_t total = (uint32_t)(p.ncs * 9 * p.b_bytes);
        mbar_arrive_expect_tx(&ctl->w_full, total);
        for (int i = 0; i < p.ncs * 9; ++i)
          bulk_load_1d(b_area + (size_t)i * p.b_bytes, p.wpacked + (size_t)i * p.b_bytes, (uint32_t)p.b_bytes,
                       &ctl->w_full);
      }
      __syncwarp();
      // Shared-memory epilogue (resident launches only): wait until an epilogue group has filled a
      // staging buffer, hand it to TMA and give the buffer back once TMA has read it.
      if (smem_epi && lane == 0) {
        int ob = 0;
        uint32_t oph = 0;
        // in-place residual: this thread also fetches the residual tiles -- the tile that reuses a staging
        // buffer is requested the moment TMA has read the previous output out of it
        auto fetch_res = [&](int tt, int buf) {
          if (tt >= total_tiles) return;
          int n_tile, X0, Y0, n;
          halo_decode(p, tt, n_tile, X0, Y0, n);
          const int groups = (p.BN + 63) >> 6;
          const uint32_t gbytes = p.BN >= 64 ? 16384u : 8192u;
          mbar_arrive_expect_tx(&ctl->res_full[buf], groups * gbytes);
          for (int g = 0; g < groups; ++g)
            tma_load_5d(p.res_map, &ctl->res_full[buf], out_stage + (size_t)buf * p.out_buf_bytes + g * gbytes,
                        p.cout_off + n_tile * p.BN + g * 64, X0, 0, Y0, n);
        };
        if (p.res_inplace)
          for (int i = 0; i < p.out_bufs; ++i) fetch_res(blockIdx.x + i * (int)gridDim.x, i);
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
          mbar_wait(&ctl->out_full[ob], oph);
          if (!(p.dbg & 4)) {
            int n_tile, X0, Y0, n;
            halo_decode(p, t, n_tile, X0, Y0, n);
            if (p.out_f32) {
              tma_store_5d(p.out_map, out_stage + (size_t)ob * p.out_buf_bytes, 0, X0, 0, Y0, n);
            } else {
              const uint32_t gbytes = p.BN >= 64 ? 16384u : 8192u;
              for (int g = 0; g < ((p.BN + 63) >> 6); ++g)
                tma_store_5d(p.out_map, out_stage + (size_t)ob * p.out_buf_bytes + g * gbytes, p.cout_off + g * 64, X0,
                             0, Y0, n);
            }
          }
          tma_store_commit();
          // this warp has nothing else to do; releasing one store late would make the two
          // epilogue groups wait for each other
          tma_store_wait_read<0>();
          if (p.res_inplace) fetch_res(t + p.out_bufs * (int)gridDim.x, ob);
          else mbar_arrive(&ctl->out_empty[ob]);
          if (++ob == p.out_bufs) {
            ob = 0;
            oph ^= 1;
          }
        }
        tma_store_wait_all<0>();
      }
    } else {
      int bs = 0;
      uint32_t bph = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int n_tile = t % p.n_tiles;
        const uint8_t* wsrc = p.wpacked + (size_t)n_tile * p.ncs * 9 * p.b_bytes;
        for (int i = 0; i < p.ncs * 9; ++i) {
          mbar_wait(&ctl->b_empty[bs], bph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&ctl->b_full[bs], (uint32_t)p.b_bytes);
            bulk_load_1d(b_area + (size_t)bs * p.b_bytes, wsrc + (size_t)i * p.b_bytes, (uint32_t)p.b_bytes,
                         &ctl->b_full[bs]);
          }
          __syncwarp();
          if (++bs == p.b_stages) {
            bs = 0;
            bph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 || warp == HALO_MMA2_WARP) {
    // ===================== MMA issuer(s) =====================
    // Issuing one tcgen05.mma costs a single warp ~50 cycles of dependent instruction latency
    // and every tile ~1000 more (waits, fences, commits) -- measured with p.prof -- which exceeds
    // the 32-cycle execution of an N <= 64 MMA.  With mma_warps == 2 (resident weights, one
    // slab per tile) two warps issue alternate tiles into alternate accumulator stages.
    const int mw = warp == 1 ? 0 : 1;
    if (mw < MW) {
      int as = mw, bs = 0, acc = mw;
      uint32_t aph = 0, bph = 0, acc_phase = 0;
      const uint32_t idesc = umma_idesc_act(128, p.BN);
      const uint64_t a_desc0 = desc_compact<P>(smem_u32(a_ring), (uint32_t)HW * P);
      const uint64_t b_desc0 = desc_compact<P>(smem_u32(b_area), 8 * P);
      const uint32_t a_step = (uint32_t)p.a_stage_bytes >> 4;
      const uint32_t b_step = (uint32_t)p.b_bytes >> 4;
      const uint32_t row_units = (uint32_t)(p.dil * HW * (P / 16));  // one dilated halo row, in 16-byte units
      const uint32_t col_units = (uint32_t)(p.dil * (P / 16));
      if (resident) mbar_wait(&ctl->w_full, 0);
      unsigned long long* pr = (p.prof && blockIdx.x == 0 && lane == 0 && mw == 0) ? p.prof : nullptr;
      const long long mstart = pr ? dev_clock() : 0;
      ProfClock pc(pr);
      for (int t = blockIdx.x + mw * gridDim.x; t < total_tiles; t += MW * gridDim.x) {
        pc.lap<3>();
        mbar_wait(&ctl->acc_empty[acc], acc_phase ^ 1);
        pc.lap<1>();
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_cols;
        for (int cs = 0; cs < p.ncs; ++cs) {
          pc.lap<3>();
          mbar_wait(&ctl->a_full[as], aph);
          pc.lap<2>();
          tc_fence_after_sync();
          const uint64_t a_stage_desc = desc_add(a_desc0, as * a_step);
          if (resident) {
            if (elect_one()) {
              if (p.dbg & 2)
                umma_bf16_ss(d_tmem, a_stage_desc, desc_add(b_desc0, cs * 9 * b_step), idesc, cs != 0 ? 1u : 0u);
              else if (p.use_kmask)
                issue_slab_masked<9, KSTEPS>(d_tmem, a_stage_desc, desc_add(b_desc0, cs * 9 * b_step), b_step, row_units,
                                             col_units, idesc, cs != 0, p.kmask[cs]);
              else
                issue_slab_resident<9, KSTEPS>(d_tmem, a_stage_desc, desc_add(b_desc0, cs * 9 * b_step), b_step,
                                               row_units, col_units, idesc, cs != 0);
              umma_commit(&ctl->a_empty[as]);
              if (cs == p.ncs - 1) umma_commit(&ctl->acc_full[acc]);
            }
            __syncwarp();
          } else {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&ctl->b_full[bs], bph);
              tc_fence_after_sync();
              if (elect_one()) {
                const uint64_t at = desc_add(a_stage_desc, (tap / 3) * row_units + (tap % 3) * col_units);
                const uint64_t bt = desc_add(b_desc0, bs * b_step);
                if (MT == 2) {
                  // second tile of the stage: 8 pixels to the right, next accumulator column range
                  const uint32_t d2 = d_tmem + (uint32_t)p.BN;
                  const uint64_t at2 = desc_add(at, 8 * (P / 16));
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k) {
                    const uint32_t accf = (tap | k) != 0 ? 1u : (cs != 0 ? 1u : 0u);
                    umma_bf16_ss(d_tmem, desc_add(at, 2 * k), desc_add(bt, 2 * k), idesc, accf);
                    umma_bf16_ss(d2, desc_add(at2, 2 * k), desc_add(bt, 2 * k), idesc, accf);
                  }
                } else {
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k)
                    umma_bf16_ss(d_tmem, desc_add(at, 2 * k), desc_add(bt, 2 * k), idesc,
                                 (tap | k) != 0 ? 1u : (cs != 0 ? 1u : 0u));
                }
                umma_commit(&ctl->b_empty[bs]);
                if (tap == 8) {
                  umma_commit(&ctl->a_empty[as]);
                  if (cs == p.ncs - 1) umma_commit(&ctl->acc_full[acc]);
                }
              }
              __syncwarp();
              if (++bs == p.b_stages) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
          as += MW;  // MW == 2 only with ncs == 1: the slab sequence is the tile sequence
          if (as >= p.a_stages) {
            as -= p.a_stages;
            aph ^= 1;
          }
        }
        acc += MW;
        if (acc >= p.acc_stages) {
          acc -= p.acc_stages;
          acc_phase ^= 1;
        }
      }
      if (pr) {
        pc.a[0] = (unsigned long long)(dev_clock() - mstart);
        pc.a[4] = pc.a[5] = pc.a[6] = pc.a[7] = 0;
        pc.flush(8);
      }
    }
  } else if (smem_epi) {
    // ===================== epilogue through shared memory (stores by the store warp) =====================
    // One group of eight warps (columns split in halves) or, for BN <= 64, two groups of four
    // warps on alternate tiles (a thread owns all BN columns of its row).  With two groups the
    // accumulator stage, residual buffer and staging buffer of a tile all have its group's parity.
    const int quarter = warp & 3;
    const int half = (warp - 3) >> 2;
    const int grp = G == 2 ? half : 0;
    const int row = quarter * 32 + lane;
    const bool has_res = p.res_map != nullptr;
    const uint32_t out_addr0 = smem_u32(out_stage), res_addr0 = smem_u32(res_stage);
    const uint32_t f32_pitch = (uint32_t)p.cout * 4u;
    const int c_begin = G == 2 ? 0 : half * 32, c_step = G == 2 ? 32 : 64;
    const float* bt = bias_l;  // n_tiles == 1 on this path
    int acc = grp, ob = grp, rb = grp;
    uint32_t acc_phase = 0, oph = 0, rph = 0;
    unsigned long long* pr = (p.prof && blockIdx.x == 0 && threadIdx.x == 96) ? p.prof : nullptr;
    const long long estart = pr ? dev_clock() : 0;
    ProfClock pc(pr);
    for (int t = blockIdx.x + grp * gridDim.x; t < total_tiles; t += G * gridDim.x) {
      pc.count<7>();
      const uint32_t ost = out_addr0 + (uint32_t)ob * p.out_buf_bytes;
      const uint32_t rst = res_addr0 + (uint32_t)rb * p.out_buf_bytes;
      pc.lap<6>();
      mbar_wait(&ctl->acc_full[acc], acc_phase);
      pc.lap<2>();
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (uint32_t)acc * acc_cols + ((uint32_t)(quarter * 32) << 16);
      if (!p.out_f32) {
        // both 32-column chunks of a BN = 64 tile are requested before the one wait
        uint32_t v[2][32];
        const int c1 = c_begin + c_step;
        tmem_ld_32x32b_x32(taddr + c_begin, v[0]);
        if (c1 < p.BN) tmem_ld_32x32b_x32(taddr + c1, v[1]);
        tmem_ld_wait();
        if (has_res) mbar_wait(&ctl->res_full[rb], rph);
        // TMA has read the previous tile out of this buffer (in-place residual: implied by the residual's arrival)
        if (!p.res_inplace) mbar_wait(&ctl->out_empty[ob], oph ^ 1);
        pc.lap<3>();
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int c = c_begin + ci * c_step;
          if (c >= p.BN) break;
          // staging rows hold 64 channels (128 B, SW128) or, for BN == 32, 32 channels (64 B, SW64)
          const bool wide = p.BN >= 64;
          const uint32_t row_off = wide ? (uint32_t)(c >> 6) * 16384u + (uint32_t)row * 128u : (uint32_t)row * 64u;
          const uint32_t swz_row = wide ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);
          const int j0 = (c & 63) >> 3;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t off = row_off + (((uint32_t)(j0 + q) ^ swz_row) << 4);
            const float4 b0 = *reinterpret_cast<const float4*>(bt + c + q * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(bt + c + q * 8 + 4);
            float f0 = __uint_as_float(v[ci][q * 8 + 0]) + b0.x, f1 = __uint_as_float(v[ci][q * 8 + 1]) + b0.y;
            float f2 = __uint_as_float(v[ci][q * 8 + 2]) + b0.z, f3 = __uint_as_float(v[ci][q * 8 + 3]) + b0.w;
            float f4 = __uint_as_float(v[ci][q * 8 + 4]) + b1.x, f5 = __uint_as_float(v[ci][q * 8 + 5]) + b1.y;
            float f6 = __uint_as_float(v[ci][q * 8 + 6]) + b1.z, f7 = __uint_as_float(v[ci][q * 8 + 7]) + b1.w;
            if (has_res) {
              const uint4 rv = lds128(rst + off);
              float2 r;
              r = unpack_act2(rv.x); f0 += r.x; f1 += r.y;
              r = unpack_act2(rv.y); f2 += r.x; f3 += r.y;
              r = unpack_act2(rv.z); f4 += r.x; f5 += r.y;
              r = unpack_act2(rv.w); f6 += r.x; f7 += r.y;
            }
            uint4 pk;
            if (p.relu) {
              pk.x = pack2<true>(f0, f1); pk.y = pack2<true>(f2, f3); pk.z = pack2<true>(f4, f5); pk.w = pack2<true>(f6, f7);
            } else {
              pk.x = pack2<false>(f0, f1); pk.y = pack2<false>(f2, f3); pk.z = pack2<false>(f4, f5); pk.w = pack2<false>(f6, f7);
            }
            sts128(ost + off, pk);
          }
        }
      } else {
        // f32 logits (cout <= 32): dense rows of cout floats, no swizzle
        const bool active = G == 2 || half == 0;
        uint32_t v[32];
        if (active) {
          tmem_ld_32x32b_x32(taddr, v);
          tmem_ld_wait();
        }
        mbar_wait(&ctl->out_empty[ob], oph ^ 1);
        pc.lap<3>();
        if (active) {
          const uint32_t row_off = (uint32_t)row * f32_pitch;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (q * 4 < p.cout) {
              const float4 b0 = *reinterpret_cast<const float4*>(bt + q * 4);
              uint4 o;
              float f0 = __uint_as_float(v[q * 4 + 0]) + b0.x, f1 = __uint_as_float(v[q * 4 + 1]) + b0.y;
              float f2 = __uint_as_float(v[q * 4 + 2]) + b0.z, f3 = __uint_as_float(v[q * 4 + 3]) + b0.w;
              if (p.relu) {
                f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); f2 = fmaxf(f2, 0.f); f3 = fmaxf(f3, 0.f);
              }
              o.x = __float_as_uint(f0); o.y = __float_as_uint(f1); o.z = __float_as_uint(f2); o.w = __float_as_uint(f3);
              sts128(ost + row_off + q * 16, o);
            }
          }
        }
      }
      pc.lap<4>();
      // publish: staging buffer to the async proxy / store warp, accumulator stage and
      // residual buffer back to their producers -- one arrival per warp
      tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&ctl->out_full[ob]);
        mbar_arrive(&ctl->acc_empty[acc]);
        if (has_res && !p.res_inplace) mbar_arrive(&ctl->res_empty[rb]);
      }
      pc.lap<5>();
      acc += G;
      if (acc >= p.acc_stages) {
        acc -= p.acc_stages;
        acc_phase ^= 1;
      }
      if (G == 2) {
        oph ^= 1;  // own buffer, every tile
      } else if (++ob == p.out_bufs) {
        ob = 0;
        oph ^= 1;
      }
      if (has_res) {
        rb += G;
        if (rb >= p.res_bufs) {
          rb -= p.res_bufs;
          rph ^= 1;
        }
      }
    }
    if (pr) {
      pc.a[0] = (unsigned long long)(dev_clock() - estart);
      pc.flush(16);
    }
  } else {
    // ===================== direct epilogue (per-thread global stores) =====================
    int acc = 0;
    uint32_t acc_phase = 0;
    const int quarter = warp & 3;
    const int half = (warp - 3) >> 2;
    const int row = quarter * 32 + lane;
    const int xi = row & 7, yi = row >> 3;
    const EpiOut eo{p.out, p.residual, p.out_f32, p.relu, p.cout};
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n_tile, X0, Y0, n;
      halo_decode(p, t, n_tile, X0, Y0, n);
      const int ch0 = p.cout_off + n_tile * p.BN;
      for (int m = 0; m < MT; ++m) {
        const int ox = X0 + m * 8 + xi, oy = Y0 + yi;
        const bool valid = ox < p.W && oy < p.H;
        const int64_t pix = ((int64_t)n * p.H + oy) * p.W + ox;
        uint4 rpre0[4], rpre1[4];
        bool have0 = false, have1 = false;
        if (valid) {
          if (half * 32 < p.BN) have0 = prefetch_residual32(eo, pix, ch0 + half * 32, rpre0);
          if (half * 32 + 64 < p.BN) have1 = prefetch_residual32(eo, pix, ch0 + half * 32 + 64, rpre1);
        }
        if (m == 0) {
          mbar_wait(&ctl->acc_full[acc], acc_phase);
          tc_fence_after_sync();
        }
        const uint32_t taddr =
            tmem_base + (uint32_t)acc * acc_cols + (uint32_t)(m * p.BN) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {  // BN <= 256: at most four 32-column chunks per thread
          const int c = half * 32 + ci * 64;
          if (c >= p.BN) break;
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c, v);
          tmem_ld_wait();
          if (valid) {
            if (ci == 0) epilogue_chunk32_pre(eo, v, bias_s, pix, ch0 + c, rpre0, have0);
            else if (ci == 1) epilogue_chunk32_pre(eo, v, bias_s, pix, ch0 + c, rpre1, have1);
            else epilogue_chunk32(eo, v, bias_s, pix, ch0 + c);
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&ctl->acc_empty[acc]);
      if (++acc == p.acc_stages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// =====================================================================================
// Entry-list halo kernel (see conv_halo.cuh, ConvHaloElParams): TMA halo boxes from several
// sources, weight images streamed with per-image GEMM column ranges, direct epilogue that
// un-shuffles the space-to-depth columns into the plain output tensor.
//   warp 0      A producer: one TMA halo box per (tile, slab ref)
//   warp 1      TMEM allocator + MMA issuer
//   warp 2      B producer: one bulk copy per entry through the weight ring
//   warps 3..10 epilogue
// =====================================================================================
namespace {
struct HaloElTables {
  HaloSlabRef slabs[HALO_EL_MAX_SLABS];
  HaloEntry entries[HALO_EL_MAX_ENTRIES];
};
constexpr int kHaloElBiasBytes = 1024 * 4;  // n_tiles * BN <= 1024
static_assert(sizeof(HaloEntry) == 16 && (sizeof(HaloSlabRef) * HALO_EL_MAX_SLABS) % 16 == 0,
              "the MMA warps read entries with 16-byte shared-memory loads");
__device__ __forceinline__ void halo_el_decode(const ConvHaloElParams& p, int t, int& n_tile, int& X0, int& Y0, int& n) {
  auto fdiv = [](uint32_t v, const FastDiv& f) { return f.m ? __umulhi(v, f.m) : v; };
  uint32_t sp = fdiv((uint32_t)t, p.div_n_tiles);
  n_tile = t - (int)(sp * p.div_n_tiles.d);
  uint32_t q = fdiv(sp, p.div_tx);
  const int tx = (int)(sp - q * p.div_tx.d);
  sp = q;
  q = fdiv(sp, p.div_ty);
  const int ty = (int)(sp - q * p.div_ty.d);
  n = (int)q + p.n_base;
  X0 = tx * (8 * p.mt);
  Y0 = ty * 16;
}
}  // namespace

template <int MT>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_el_kernel(const __grid_constant__ ConvHaloElParams p) {
  constexpr int P = 128, KSTEPS = 4;
  constexpr int HW = 8 * MT + 2, HH = 18;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_area = smem + (size_t)p.a_stages * p.a_stage_bytes;
  uint8_t* stage = b_area + (size_t)p.b_stages * p.b_bytes;  // 2 x 16 KB when out_map, 1024-aligned
  HaloCtl* ctl = reinterpret_cast<HaloCtl*>(stage + (p.out_map ? 32768 : 0));
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctl) + kHaloCtlBytes);
  HaloElTables* tab = reinterpret_cast<HaloElTables*>(reinterpret_cast<uint8_t*>(bias_s) + kHaloElBiasBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.n_tiles * p.tiles_x * p.tiles_y * p.NB;
  // Accumulators: one slot of BN columns per (stage, 8 x 16 tile of the stage); slot = stage * MT + m.
  // MT * BN <= 256: two stages; MT * BN == 512 (BN = 256, MT = 2): one stage, and because every slot has its
  // own full / empty barriers the MMA warp of tile m restarts as soon as the epilogue has drained ITS slot.
  const int acc_stages = MT * p.BN <= 256 ? 2 : 1;

  if (threadIdx.x == 0) {
    // MT == 2: two MMA-issuing warps, one per 8 x 16 tile of the stage (a single warp issues an MMA every
    // ~50 cycles, an N <= 96 MMA executes in 40..56); both commit to the operand rings' barriers
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(&ctl->a_full[i], 1);
      mbar_init(&ctl->a_empty[i], MT);
    }
    for (int i = 0; i < HALO_MAX_B_STAGES; ++i) {
      mbar_init(&ctl->b_full[i], 1);
      mbar_init(&ctl->b_empty[i], MT);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&ctl->acc_full[i], 1);
      mbar_init(&ctl->acc_empty[i], 32 * HALO_EPI_WARPS);
    }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < p.n_tiles * p.BN; i += HALO_THREADS) bias_s[i] = p.bias[i];
  for (int i = threadIdx.x; i < p.n_slabs; i += HALO_THREADS) tab->slabs[i] = p.slabs[i];
  for (int i = threadIdx.x; i < p.n_entries; i += HALO_THREADS) tab->entries[i] = p.entries[i];
  if (warp == 1) tmem_alloc<512>(&ctl->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================== A producer =====================
    int as = 0;
    uint32_t aph = 0;
    constexpr uint32_t a_box_bytes = (uint32_t)(HW * HH * P);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n_tile, X0, Y0, n;
      halo_el_decode(p, t, n_tile, X0, Y0, n);
      for (int sr = p.tile_begin[n_tile]; sr < p.tile_begin[n_tile + 1]; ++sr) {
        const HaloSlabRef s = tab->slabs[sr];
        mbar_wait(&ctl->a_empty[as], aph ^ 1);
        if (elect_one()) {
          if (p.dbg & 1) {
            mbar_arrive(&ctl->a_full[as]);
          } else {
            mbar_arrive_expect_tx(&ctl->a_full[as], a_box_bytes);
            tma_load_5d(p.map + s.map, &ctl->a_full[as], a_ring + (size_t)as * p.a_stage_bytes, s.c, X0 - 1, s.p, Y0 - 1, n);
          }
        }
        __syncwarp();
        if (++as == p.a_stages) {
          as = 0;
          aph ^= 1;
        }
      }
    }
  } else if (warp == 2) {
    // ===================== B producer =====================
    int bs = 0;
    uint32_t bph = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n_tile, X0, Y0, n;
      halo_el_decode(p, t, n_tile, X0, Y0, n);
      const int e_begin = tab->slabs[p.tile_begin[n_tile]].e_begin;
      const int e_end = tab->slabs[p.tile_begin[n_tile + 1] - 1].e_end;
      for (int ei = e_begin; ei < e_end; ++ei) {  // the entries of a tile's slab refs are consecutive
        const HaloEntry en = tab->entries[ei];
        const uint32_t bytes = en.grp & 0x7fffffffu;
        if (!bytes) continue;  // not the first entry of its group
        mbar_wait(&ctl->b_empty[bs], bph ^ 1);
        if (elect_one()) {
          if (p.dbg & 2) {
            mbar_arrive(&ctl->b_full[bs]);
          } else {
            mbar_arrive_expect_tx(&ctl->b_full[bs], bytes);
            bulk_load_1d(b_area + (size_t)bs * p.b_bytes, p.wpacked + en.w_off, bytes, &ctl->b_full[bs]);
          }
        }
        __syncwarp();
        if (++bs == p.b_stages) {
          bs = 0;
          bph ^= 1;
        }
      }
    }
  } else if (warp == 1 || (MT == 2 && warp == HALO_MMA2_WARP)) {
    // ===================== MMA issuer(s) =====================
    // (a loop run by lane 0 alone instead of elect-per-entry measured 10 % slower)
    const int mw = warp == 1 ? 0 : 1;  // the 8 x 16 tile of the stage this warp issues
    int as = 0, bs = 0, acc = 0;
    uint32_t aph = 0, bph = 0, acc_phase = 0;
    const uint64_t a_desc0 = desc_add(desc_compact<P>(smem_u32(a_ring), (uint32_t)HW * P), mw * 8 * (P / 16));
    const uint64_t b_desc0 = desc_compact<P>(smem_u32(b_area), 8 * P);
    const uint32_t a_step = (uint32_t)p.a_stage_bytes >> 4;
    const uint32_t b_step = (uint32_t)p.b_bytes >> 4;
    const uint32_t entries_s = smem_u32(tab->entries);
    const bool one_mma = (p.dbg & 8) != 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n_tile, X0, Y0, n;
      halo_el_decode(p, t, n_tile, X0, Y0, n);
      const int slot = acc * MT + mw;
      mbar_wait(&ctl->acc_empty[slot], acc_phase ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + (uint32_t)(slot * p.BN);
      const int sr_end = p.tile_begin[n_tile + 1];
      uint32_t accf = 0;  // the first entry of a tile covers all BN columns (host guarantee)
      for (int sr = p.tile_begin[n_tile]; sr < sr_end; ++sr) {
        const HaloSlabRef s = tab->slabs[sr];
        // entries are read with explicit 16-byte shared-memory loads, one entry ahead (the compiler turned the
        // struct read into three generic 4-byte loads at the head of a ~120-instruction dependent chain per entry)
        uint4 nxt = lds128(entries_s + (uint32_t)s.e_begin * 16u);
        mbar_wait(&ctl->a_full[as], aph);
        tc_fence_after_sync();
        const uint64_t a_stage_desc = desc_add(a_desc0, as * a_step);
        for (int ei = s.e_begin; ei < s.e_end; ++ei) {
          HaloEntry en;
          en.ab_off16 = nxt.x;
          en.w_off = nxt.y;
          en.ncol0_n = nxt.z;
          en.grp = nxt.w;
          if (ei + 1 < s.e_end) nxt = lds128(entries_s + (uint32_t)(ei + 1) * 16u);
          if (en.grp & 0x7fffffffu) {  // first entry of a group: its images have landed
            mbar_wait(&ctl->b_full[bs], bph);
            tc_fence_after_sync();
          }
          const bool last = (en.grp >> 31) != 0;
          if (elect_one()) {
            const uint32_t ncol0 = en.ncol0_n & 0xfffu, nn = (en.ncol0_n >> 16) & 0xfffu, kmask = en.ncol0_n >> 28;
            const uint32_t idesc = umma_idesc_act(128, (int)nn);
            const uint64_t at = desc_add(a_stage_desc, en.ab_off16 & 0xffffu);
            const uint64_t bt = desc_add(b_desc0, bs * b_step + (en.ab_off16 >> 16));
            const uint32_t d = d_tmem + ncol0;
            if (kmask == 15u && !one_mma) {  // the common case, without per-K-step tests
              umma_bf16_ss(d, at, bt, idesc, accf);
              umma_bf16_ss(d, desc_add(at, 2), desc_add(bt, 2), idesc, 1u);
              umma_bf16_ss(d, desc_add(at, 4), desc_add(bt, 4), idesc, 1u);
              umma_bf16_ss(d, desc_add(at, 6), desc_add(bt, 6), idesc, 1u);
            } else {
              uint32_t af = accf;
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                if ((kmask >> k) & 1u) {  // K-steps whose weights are all zero are not issued
                  umma_bf16_ss(d, desc_add(at, 2 * k), desc_add(bt, 2 * k), idesc, af);
                  af = 1u;
                  if (one_mma) break;
                }
            }
            if (last) umma_commit(&ctl->b_empty[bs]);
            if (ei == s.e_end - 1) {
              umma_commit(&ctl->a_empty[as]);
              if (sr == sr_end - 1) umma_commit(&ctl->acc_full[slot]);
            }
          }
          __syncwarp();
          accf = 1;
          if (last && ++bs == p.b_stages) {
            bs = 0;
            bph ^= 1;
          }
        }
        if (++as == p.a_stages) {
          as = 0;
          aph ^= 1;
        }
      }
      if (++acc == acc_stages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp < 3 + HALO_EPI_WARPS && p.out_map) {
    // ===================== epilogue through shared memory + TMA store =====================
    // Work unit = (8 x 16 tile m, 64-column group g): 128 pixels x 64 channels = 16 KB, staged in the SW128
    // layout of the FOLDED output map (px*C + c, x/2, py, y/2, n): the 64 columns are the two horizontal
    // sub-pixels of 32 channels, or 64 channels of one sub-pixel, of row parity a -- one box {64, 8, 1, 16, 1}.
    // Buffers alternate; one barrier per unit: [all: stage unit j] -> [thread 0: wait until the stores of
    // units < j have been read] -> barrier -> [thread 0: store unit j] (conv_tc.cu has the same scheme).
    int acc = 0;
    uint32_t acc_phase = 0;
    const int quarter = warp & 3;
    const int half = (warp - 3) >> 2;
    const int row = quarter * 32 + lane;
    const bool issuer = warp == 3 && lane == 0;
    const uint32_t stage0 = smem_u32(stage);
    const bool relu = p.relu != 0;
    uint32_t unit = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n_tile, X0, Y0, n;
      halo_el_decode(p, t, n_tile, X0, Y0, n);
      const int ch0 = n_tile * p.BN;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const int slot = acc * MT + m;
        mbar_wait(&ctl->acc_full[slot], acc_phase);
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + (uint32_t)(slot * p.BN) + ((uint32_t)(quarter * 32) << 16);
        for (int g = 0; g * 64 < p.BN; ++g, ++unit) {
          const int c = g * 64 + half * 32;
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c, v);
          tmem_ld_wait();
          if (g * 64 + 64 >= p.BN) {  // this thread's last read of the slot
            tc_fence_before_sync();
            mbar_arrive(&ctl->acc_empty[slot]);
          }
          const float* bs = bias_s + ch0 + c;
          const uint32_t rowa = stage0 + (unit & 1u) * 16384u + (uint32_t)row * 128u;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b0 = *reinterpret_cast<const float4*>(bs + q * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(bs + q * 8 + 4);
            const float f0 = __uint_as_float(v[q * 8 + 0]) + b0.x, f1 = __uint_as_float(v[q * 8 + 1]) + b0.y;
            const float f2 = __uint_as_float(v[q * 8 + 2]) + b0.z, f3 = __uint_as_float(v[q * 8 + 3]) + b0.w;
            const float f4 = __uint_as_float(v[q * 8 + 4]) + b1.x, f5 = __uint_as_float(v[q * 8 + 5]) + b1.y;
            const float f6 = __uint_as_float(v[q * 8 + 6]) + b1.z, f7 = __uint_as_float(v[q * 8 + 7]) + b1.w;
            uint4 pk;
            if (relu) {
              pk.x = pack2<true>(f0, f1); pk.y = pack2<true>(f2, f3); pk.z = pack2<true>(f4, f5); pk.w = pack2<true>(f6, f7);
            } else {
              pk.x = pack2<false>(f0, f1); pk.y = pack2<false>(f2, f3); pk.z = pack2<false>(f4, f5); pk.w = pack2<false>(f6, f7);
            }
            sts128(rowa + ((uint32_t)((half * 4 + q) ^ (row & 7)) << 4), pk);
          }
          fence_proxy_async_smem();
          if (issuer) tma_store_wait_read<0>();
          named_bar_sync(1, 32 * HALO_EPI_WARPS);
          if (issuer && !(p.dbg & 4)) {
            const int chs = ch0 + g * 64;
            if (p.s2d_out) {
              const int cls = chs >> p.cout_log2, cpl = chs & (p.cout - 1);
              tma_store_5d(p.out_map, stage + (unit & 1u) * 16384u, (cls & 1) * p.cout + cpl, X0 + m * 8, cls >> 1, Y0, n);
            } else {
              tma_store_5d(p.out_map, stage + (unit & 1u) * 16384u, chs, X0 + m * 8, 0, Y0, n);
            }
            tma_store_commit();
          }
        }
      }
      if (++acc == acc_stages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (issuer) tma_store_wait_all<0>();
  } else if (warp < 3 + HALO_EPI_WARPS) {
    // ===================== direct epilogue (per-thread global stores) =====================
    int acc = 0;
    uint32_t acc_phase = 0;
    const int quarter = warp & 3;
    const int half = (warp - 3) >> 2;
    const int row = quarter * 32 + lane;
    const int xi = row & 7, yi = row >> 3;
    const EpiOut eo{p.out, nullptr, 0, p.relu, p.cout};
    const int OW = 2 * p.W, OH = 2 * p.H;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n_tile, X0, Y0, n;
      halo_el_decode(p, t, n_tile, X0, Y0, n);
      const int ch0 = n_tile * p.BN;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const int slot = acc * MT + m;
        mbar_wait(&ctl->acc_full[slot], acc_phase);
        tc_fence_after_sync();
        const int ox = X0 + m * 8 + xi, oy = Y0 + yi;
        const bool valid = ox < p.W && oy < p.H && !(p.dbg & 4);
        const uint32_t taddr = tmem_base + (uint32_t)(slot * p.BN) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = half * 32 + ci * 64;
          if (c >= p.BN) break;
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c, v);
          tmem_ld_wait();
          if (valid) {
            // space-to-depth columns [chs, chs + 32): sub-pixel (a, b) = class, plain channels cpl .. cpl + 31
            const int chs = ch0 + c;
            if (p.s2d_out) {
              const int cls = chs >> p.cout_log2, cpl = chs & (p.cout - 1);
              const int64_t pix = ((int64_t)n * OH + 2 * oy + (cls >> 1)) * OW + 2 * ox + (cls & 1);
              epilogue_chunk32(eo, v, bias_s + (chs - cpl), pix, cpl);
            } else {
              epilogue_chunk32(eo, v, bias_s, ((int64_t)n * p.H + oy) * p.W + ox, chs);
            }
          }
        }
        tc_fence_before_sync();
        mbar_arrive(&ctl->acc_empty[slot]);
      }
      if (++acc == acc_stages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// =====================================================================================
// Generalised halo kernel: cp.async-assembled A tiles (see conv_halo.cuh).
//   warps 0..3   A loaders (128 threads): 16-byte cp.async copies, zero-filled outside
//                the image, written at the swizzled position of a compact-pitch pixel row;
//                a stage is published after cp.async.wait_group + fence.proxy.async
//   warp 4       MMA issuer + TMEM allocator
//   warp 5       B producer (weights resident or ring, as above)
//   warps 6..13  epilogue
// MODE 0: 3x3 convolution, KC channels per slab, MT output tiles per stage.
// MODE 1: the 7x7 stride-2 single-channel stem (smp ResNetEncoder conv1): the loader
//         writes the im2col row of each output pixel (49 taps, zero padded to K = 64) and
//         one tap of four K-steps is issued per tile (KC = 64, MT = 1).
// =====================================================================================
namespace {
// logits (accumulator columns 0..C-1 + bias) of one padded pixel -> merged key.
// Arithmetic identical to head_kernel (kernels_simple.cu), so fused and unfused paths
// produce the same keys.
__device__ __forceinline__ void head_fused_pixel(const HeadFuse& hf, const uint32_t* v8, const float* bias_s,
                                                 int n, int oy, int ox) {
  const int r = oy - hf.crop_top, c = ox - hf.crop_left;
  if ((unsigned)r >= (unsigned)hf.Hc || (unsigned)c >= (unsigned)hf.Wc) return;
  float l[8];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    l[k] = k < hf.C ? __uint_as_float(v8[k]) + bias_s[k] : -INFINITY;
    m = fmaxf(m, l[k]);
  }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    l[k] = k < hf.C ? expf(l[k] - m) : 0.f;
    sum += l[k];
  }
  float best = -1.f;
  int lab = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float pk = __fdiv_rn(l[k], sum);
    if (k < hf.C && pk > best) {
      best = pk;
      lab = k;
    }
  }
  const int64_t vox = hf.base + (hf.s0 + n) * hf.stride_s + (int64_t)r * hf.stride_r + (int64_t)c * hf.stride_c;
  const uint16_t h = __half_as_ushort(__float2half_rn(best));
  atomicMax(hf.keys + vox, pack_key(h, hf.d, (uint32_t)lab, __float_as_uint(best)));
}

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
}  // namespace

template <int MODE, int KC, int MT>
__global__ void __launch_bounds__(HALO2_THREADS, 1)
conv_halo2_kernel(const __grid_constant__ ConvHalo2Params p) {
  constexpr int P = 2 * KC;                 // pixel pitch in bytes
  constexpr int KS = KC / 16;               // K-steps per slab
  constexpr int HW = 8 * MT + 2, HH = 18;   // halo tile (MODE 0)
  constexpr int NTAPS = MODE == 0 ? 9 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_area = smem + (size_t)p.a_stages * p.a_stage_bytes;
  const size_t b_area_bytes =
      p.b_stages ? (size_t)p.b_stages * p.b_bytes : (size_t)p.nslabs * NTAPS * p.b_bytes;
  HaloCtl* ctl = reinterpret_cast<HaloCtl*>(b_area + b_area_bytes);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctl) + kHaloCtlBytes);
  uint8_t* stem_raw = reinterpret_cast<uint8_t*>(bias_s) + kHaloBiasBytes;     // MODE 1: raw input ring
  uint8_t* pool_stage = stem_raw + kStemRawRing * kStemRawBytes;               // MODE 1 + pool: 2 x 16 KB

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.n_tiles * p.tiles_x * p.tiles_y * p.NB;
  const bool resident = p.b_stages == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(&ctl->a_full[i], 32 * HALO2_LOAD_WARPS);
      mbar_init(&ctl->a_empty[i], 1);
    }
    for (int i = 0; i < HALO_MAX_B_STAGES; ++i) {
      mbar_init(&ctl->b_full[i], 1);
      mbar_init(&ctl->b_empty[i], 1);
    }
    mbar_init(&ctl->w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->acc_full[i], 1);
      mbar_init(&ctl->acc_empty[i], (MODE == 1 && p.pool_out) ? 32 * 4 : 32 * 8);
    }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < p.n_tiles * p.BN; i += HALO2_THREADS) bias_s[i] = p.bias[i];
  if (warp == HALO2_LOAD_WARPS) tmem_alloc<512>(&ctl->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  // tile index -> (n_tile, X0, Y0, image)
  auto decode = [&](int t, int& n_tile, int& X0, int& Y0, int& n) {
    auto fdiv = [](uint32_t v, const FastDiv& f) { return f.m ? __umulhi(v, f.m) : v; };
    uint32_t sp = fdiv((uint32_t)t, p.div_n_tiles);
    n_tile = t - (int)(sp * p.div_n_tiles.d);
    uint32_t q = fdiv(sp, p.div_tx);
    const int tx = (int)(sp - q * p.div_tx.d);
    sp = q;
    q = fdiv(sp, p.div_ty);
    const int ty = (int)(sp - q * p.div_ty.d);
    n = (int)q + p.n_base;
    X0 = tx * (8 * MT);
    Y0 = ty * 16;
  };

  if (warp < HALO2_LOAD_WARPS) {
    // ===================== A loaders =====================
    const int ptid = threadIdx.x;  // 0..127
    int as = 0;
    uint32_t aph = 0;
    // Up to `depth` cp.async groups (= stages) stay in flight per thread; the oldest is
    // published (wait_group -> fence.proxy.async -> arrive) once `depth` newer ones exist.
    // depth = a_stages / 2 leaves the other half of the ring published ahead of the MMA
    // warp (depth = a_stages - 1 would run loader and MMA in lock-step).
    const bool MW2 = MODE == 0 && p.mma_warps == 2;
    const int ring_len = MW2 ? p.a_stages / 2 : p.a_stages;
    // (with half-rings a stage is reused after ring_len issues of its parity, so at most ring_len - 1
    // groups may stay unpublished or the loader would wait for a stage nobody was told about)
    const int depth = MW2 ? (ring_len > 1 ? ring_len - 1 : 1) : (p.a_stages / 2 > 1 ? p.a_stages / 2 : 1);
    int inflight = 0;
    int rpos[2] = {0, 0};
    uint32_t rph[2] = {0, 0};
    int tile_k = 0;
    uint8_t fifo[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // stages in issue order (at most a_stages <= 8 unpublished); 0xFF = fetched by TMA, nothing to publish
    int fifo_head = 0, fifo_tail = 0;
    auto publish_oldest = [&]() {
      switch (depth) {  // all but the `depth` most recent groups have landed
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        default: cp_async_wait<4>(); break;
      }
      if (fifo[fifo_head] != 0xFF) {
        fence_proxy_async_smem();
        mbar_arrive(&ctl->a_full[fifo[fifo_head]]);
      }
      fifo_head = (fifo_head + 1) & 7;
      --inflight;
    };
    const uint32_t ring_addr = smem_u32(a_ring);
    // Per-thread copy table: which halo pixel / 16-byte chunk this thread fetches in its
    // it-th copy of every stage, and where it lands (identical for every stage and slab).
    constexpr int CH8 = KC / 8;  // 16-byte chunks per pixel
    constexpr int NITER = MODE == 0 ? (HW * HH * CH8 + 32 * HALO2_LOAD_WARPS - 1) / (32 * HALO2_LOAD_WARPS) : 1;
    uint32_t ltab[NITER];
#pragma unroll
    for (int it = 0; it < NITER; ++it) {
      const int idx = it * 32 * HALO2_LOAD_WARPS + ptid;
      const int j = idx % CH8, pix = idx / CH8;
      const int hy = pix / HW, hx = pix - hy * HW;
      ltab[it] = ((swz<P>((uint32_t)(pix * P + j * 16)) >> 4) << 16) | ((uint32_t)hy << 10) | ((uint32_t)hx << 4) | (uint32_t)j;
    }
    int stem_k = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n_tile, X0, Y0, n;
      decode(t, n_tile, X0, Y0, n);
      if (MODE == 1) {
        // Raw input windows run kStemRawRing - 2 tiles ahead (8-byte cp.async, zero-filled outside
        // the image); the im2col row of each output pixel is then assembled from shared memory:
        // K chunk ky (16 bytes) = input columns 2*ox-4 .. 2*ox+3 of row 2*oy+ky-3, i.e. four
        // aligned 32-bit words; column 2*ox-4 is outside the 7x7 window and meets a zero weight.
        const HaloSrc& sv = p.src[0];
        auto fetch_raw = [&](int tt, int slot) {
          if (tt < total_tiles) {
            int nt2, X2, Y2, n2;
            decode(tt, nt2, X2, Y2, n2);
            const uint16_t* img2 = sv.ptr + (int64_t)n2 * sv.Hs * sv.Ws;
            const uint32_t dst0 = smem_u32(stem_raw) + (uint32_t)slot * kStemRawBytes;
            for (int i = ptid; i < kStemRawRows * 6; i += 32 * HALO2_LOAD_WARPS) {
              const int rr = i / 6, cc = i - rr * 6;
              const int iy = 2 * Y2 - 3 + rr, ix = 2 * X2 - 4 + 4 * cc;
              const bool ok = iy >= 0 && iy < sv.Hs && ix >= 0 && ix < sv.Ws;
              const uint16_t* g = ok ? img2 + (int64_t)iy * sv.Ws + ix : sv.ptr;
              asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst0 + rr * kStemRawPitch + cc * 8), "l"(g),
                           "r"(ok ? 8u : 0u)
                           : "memory");
            }
          }
          cp_async_commit();
        };
        if (t == (int)blockIdx.x) {  // prologue: windows of the first two tiles
          fetch_raw(t, 0);
          fetch_raw(t + (int)gridDim.x, 1);
        }
        const int k = stem_k++;  // CTA-local tile count
        fetch_raw(t + 2 * (int)gridDim.x, (k + 2) % kStemRawRing);
        cp_async_wait<2>();  // this tile's window has landed (two newer groups may be in flight)
        named_bar_sync(5, 32 * HALO2_LOAD_WARPS);  // ... for every loader thread
        const int m = ptid, oxl = m & 7, oyl = m >> 3;
        const uint32_t raw = smem_u32(stem_raw) + (uint32_t)(k % kStemRawRing) * kStemRawBytes;
        uint32_t vals[32];
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
          const uint32_t rowa = raw + (uint32_t)(2 * oyl + ky) * kStemRawPitch + (uint32_t)oxl * 4;
#pragma unroll
          for (int wq = 0; wq < 4; ++wq)
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(vals[ky * 4 + wq]) : "r"(rowa + wq * 4));
        }
#pragma unroll
        for (int i = 28; i < 32; ++i) vals[i] = 0;
        mbar_wait(&ctl->a_empty[as], aph ^ 1);
        const uint32_t row_addr = ring_addr + (uint32_t)as * p.a_stage_bytes + m * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((j ^ (m & 7)) << 4)),
                       "r"(vals[4 * j]), "r"(vals[4 * j + 1]), "r"(vals[4 * j + 2]), "r"(vals[4 * j + 3])
                       : "memory");
        fence_proxy_async_smem();
        mbar_arrive(&ctl->a_full[as]);
        if (++as == p.a_stages) {
          as = 0;
          aph ^= 1;
        }
        continue;
      }
      // With two MMA warps (alternate tiles) the ring is split in two halves, one per tile parity,
      // so a stage only ever meets one consumer warp (see the parity-aliasing note in engine.cu).
      const int par = MW2 ? (tile_k & 1) : 0;
      ++tile_k;
      for (int s = 0; s < p.nslabs; ++s) {
        const HaloSrc& sv = p.src[p.slab_src[s]];
        const int c0 = p.slab_c0[s];
        const int as = par * ring_len + rpos[par];
        mbar_wait(&ctl->a_empty[as], rph[par] ^ 1);
        const uint32_t stage_addr = ring_addr + (uint32_t)as * p.a_stage_bytes;
        if (p.src_map[p.slab_src[s]]) {
          // whole halo of this slab by one TMA box (zero fill outside the image): thread 0 arrives
          // with the byte count, the other loader threads arrive at once; the stage is complete
          // when the bytes have landed -- it takes no part in the cp.async publishing queue
          if (ptid == 0) {
            mbar_arrive_expect_tx(&ctl->a_full[as], (uint32_t)(HW * HH * P));
            tma_load_5d(p.src_map[p.slab_src[s]], &ctl->a_full[as], a_ring + (size_t)as * p.a_stage_bytes, c0, X0 - 1, 0,
                        Y0 - 1, n);
          } else {
            mbar_arrive(&ctl->a_full[as]);
          }
          // keep the publishing queue in step with the ring: an (empty) group and a placeholder entry,
          // so an older cp.async stage is still published after `depth` further ring stages
          cp_async_commit();
          fifo[fifo_tail] = 0xFF;
          fifo_tail = (fifo_tail + 1) & 7;
          if (++inflight > depth) publish_oldest();
          if (++rpos[par] == ring_len) {
            rpos[par] = 0;
            rph[par] ^= 1;
          }
          continue;
        }
        const uint16_t* img = sv.ptr + (int64_t)n * sv.Hs * sv.Ws * sv.C + c0;
        const int up = sv.up;
        const uint32_t row_elems = (uint32_t)sv.Ws * sv.C, px_elems = (uint32_t)sv.C;
#pragma unroll
        for (int it = 0; it < NITER; ++it) {
          const uint32_t e = ltab[it];  // [31:16] swizzled smem offset / 16, [15:10] hy, [9:4] hx, [3:0] chunk j
          if (it * 32 * HALO2_LOAD_WARPS + ptid < HW * HH * CH8) {
            const int y = Y0 - 1 + (int)((e >> 10) & 63), x = X0 - 1 + (int)((e >> 4) & 63);
            const bool ok = (unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W;
            const uint32_t sy = (uint32_t)(up ? y >> 1 : y), sx = (uint32_t)(up ? x >> 1 : x);
            const uint16_t* g = img + (sy * row_elems + sx * px_elems + (e & 15) * 8);
            cp_async_16(stage_addr + ((e >> 16) << 4), ok ? g : sv.ptr, ok ? 16u : 0u);
          }
        }
        cp_async_commit();
        fifo[fifo_tail] = (uint8_t)as;
        fifo_tail = (fifo_tail + 1) & 7;
        if (++inflight > depth) publish_oldest();
        if (++rpos[par] == ring_len) {
          rpos[par] = 0;
          rph[par] ^= 1;
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    while (fifo_head != fifo_tail) {
      if (fifo[fifo_head] != 0xFF) mbar_arrive(&ctl->a_full[fifo[fifo_head]]);
      fifo_head = (fifo_head + 1) & 7;
    }
  } else if (warp == HALO2_LOAD_WARPS + 1) {
    // ===================== B producer =====================
    if (resident) {
      if (elect_one()) {
        mbar_arrive_expect_tx(&ctl->w_full, (uint32_t)(p.nslabs * NTAPS * p.b_bytes));
        for (int i = 0; i < p.nslabs * NTAPS; ++i)
          bulk_load_1d(b_area + (size_t)i * p.b_bytes, p.wpacked + (size_t)i * p.b_bytes, (uint32_t)p.b_bytes,
                       &ctl->w_full);
      }
      __syncwarp();
    } else {
      int bs = 0;
      uint32_t bph = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int n_tile = t % p.n_tiles;
        const uint8_t* wsrc = p.wpacked + (size_t)n_tile * p.nslabs * NTAPS * p.b_bytes;
        for (int i = 0; i < p.nslabs * NTAPS; ++i) {
          mbar_wait(&ctl->b_empty[bs], bph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&ctl->b_full[bs], (uint32_t)p.b_bytes);
            bulk_load_1d(b_area + (size_t)bs * p.b_bytes, wsrc + (size_t)i * p.b_bytes, (uint32_t)p.b_bytes,
                         &ctl->b_full[bs]);
          }
          __syncwarp();
          if (++bs == p.b_stages) {
            bs = 0;
            bph ^= 1;
          }
        }
      }
    }
  } else if (warp == HALO2_LOAD_WARPS || warp == HALO2_MMA2_WARP) {
    // ===================== MMA issuer(s) =====================
    // mma_warps == 2 (MODE 0, resident weights): alternate tiles, own half of the A ring and own
    // accumulator stage per warp -- one warp cannot issue N <= 64 MMAs as fast as they execute.
    const bool MW2 = MODE == 0 && p.mma_warps == 2;
    const int mw = warp == HALO2_LOAD_WARPS ? 0 : 1;
    if (mw == 1 && !MW2) goto mma_done;
    {
    const int ring_len = MW2 ? p.a_stages / 2 : p.a_stages;
    const int ring_base = MW2 ? mw * ring_len : 0;
    int rpos = 0, bs = 0, acc = MW2 ? mw : 0;
    uint32_t aph = 0, bph = 0, acc_phase = 0;
    const uint32_t idesc = umma_idesc_act(128, p.BN);
    const uint64_t a_desc0 = desc_compact<P>(smem_u32(a_ring), MODE == 0 ? HW * P : 8 * P);
    const uint64_t b_desc0 = desc_compact<P>(smem_u32(b_area), 8 * P);
    const uint32_t a_step = (uint32_t)p.a_stage_bytes >> 4;
    const uint32_t b_step = (uint32_t)p.b_bytes >> 4;
    constexpr uint32_t PX = P / 16;  // one pixel, in descriptor address units
    if (resident) mbar_wait(&ctl->w_full, 0);
    for (int t = blockIdx.x + (MW2 ? mw * (int)gridDim.x : 0); t < total_tiles; t += (MW2 ? 2 : 1) * gridDim.x) {
      mbar_wait(&ctl->acc_empty[acc], acc_phase ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
      for (int s = 0; s < p.nslabs; ++s) {
        const int as = ring_base + rpos;
        mbar_wait(&ctl->a_full[as], aph);
        tc_fence_after_sync();
        const uint64_t a_stage_desc = desc_add(a_desc0, as * a_step);
        if (resident) {
          if (elect_one()) {
            const uint64_t b_slab = desc_add(b_desc0, s * NTAPS * b_step);
#pragma unroll
            for (int tap = 0; tap < NTAPS; ++tap) {
              const uint64_t bt = desc_add(b_slab, tap * b_step);
#pragma unroll
              for (int k = 0; k < KS; ++k) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  const uint64_t at = desc_add(a_stage_desc, ((tap / 3) * HW + (tap % 3) + mt * 8) * PX + 2 * k);
                  umma_bf16_ss(d_tmem + mt * p.BN, at, desc_add(bt, 2 * k), idesc,
                               (tap | k) != 0 ? 1u : (s != 0 ? 1u : 0u));
                }
              }
            }
            umma_commit(&ctl->a_empty[as]);
            if (s == p.nslabs - 1) umma_commit(&ctl->acc_full[acc]);
          }
          __syncwarp();
        } else {
#pragma unroll
          for (int tap = 0; tap < NTAPS; ++tap) {
            mbar_wait(&ctl->b_full[bs], bph);
            tc_fence_after_sync();
            if (elect_one()) {
              const uint64_t bt = desc_add(b_desc0, bs * b_step);
#pragma unroll
              for (int k = 0; k < KS; ++k) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  const uint64_t at = desc_add(a_stage_desc, ((tap / 3) * HW + (tap % 3) + mt * 8) * PX + 2 * k);
                  umma_bf16_ss(d_tmem + mt * p.BN, at, desc_add(bt, 2 * k), idesc,
                               (tap | k) != 0 ? 1u : (s != 0 ? 1u : 0u));
                }
              }
              umma_commit(&ctl->b_empty[bs]);
              if (tap == NTAPS - 1) {
                umma_commit(&ctl->a_empty[as]);
                if (s == p.nslabs - 1) umma_commit(&ctl->acc_full[acc]);
              }
            }
            __syncwarp();
            if (++bs == p.b_stages) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
        if (++rpos == ring_len) {
          rpos = 0;
          aph ^= 1;
        }
      }
      if (MW2) {
        acc_phase ^= 1;  // own accumulator stage, every tile
      } else if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    }
  mma_done:;
  } else {
    // ===================== epilogue =====================
    int acc = 0;
    uint32_t acc_phase = 0;
    const int quarter = warp & 3;
    const int half = (warp - (HALO2_LOAD_WARPS + 2)) >> 2;
    const int row = quarter * 32 + lane;
    const int xi = row & 7, yi = row >> 3;
    const EpiOut eo{p.out, p.residual, p.out_f32, p.relu, p.cout};
    const int ncols = MT * p.BN;
    const int bn_log2 = 31 - __clz(p.BN);
    if (MODE == 1 && p.pool_out) {
      // ---- stem + fused max-pool: BN == 64, full tiles (H % 16 == 0, W % 8 == 0), ReLU ----
      // Two groups of four warps take alternate tiles (accumulator stage = group), each with
      // its own 16 KB staging tile; a thread owns all 64 channels of its pixel.
      const int grp = half;
      const uint32_t stage = smem_u32(pool_stage) + (uint32_t)grp * 16384u;
      const int et = (quarter << 5) | lane;  // 0..127 within the group
      const int Hq = p.H >> 1, Wq = p.W >> 1;
      uint32_t gphase = 0;
      for (int t = blockIdx.x + grp * gridDim.x; t < total_tiles; t += 2 * gridDim.x) {
        int n_tile, X0, Y0, n;
        decode(t, n_tile, X0, Y0, n);
        mbar_wait(&ctl->acc_full[grp], gphase);
        gphase ^= 1;
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + (uint32_t)grp * 256u + ((uint32_t)(quarter * 32) << 16);
        uint32_t v[2][32];
        tmem_ld_32x32b_x32(taddr, v[0]);
        tmem_ld_32x32b_x32(taddr + 32, v[1]);
        tmem_ld_wait();
        tc_fence_before_sync();
        mbar_arrive(&ctl->acc_empty[grp]);
        if (p.out_map && et == 0) tma_store_wait_read<0>();  // ... and so has the TMA store issued from it
        named_bar_sync(1 + 2 * grp, 128);  // the group has finished reading the previous tile's staging
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + h2 * 32 + q * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + h2 * 32 + q * 8 + 4);
            uint4 pk;
            pk.x = pack2<true>(__uint_as_float(v[h2][q * 8 + 0]) + b0.x, __uint_as_float(v[h2][q * 8 + 1]) + b0.y);
            pk.y = pack2<true>(__uint_as_float(v[h2][q * 8 + 2]) + b0.z, __uint_as_float(v[h2][q * 8 + 3]) + b0.w);
            pk.z = pack2<true>(__uint_as_float(v[h2][q * 8 + 4]) + b1.x, __uint_as_float(v[h2][q * 8 + 5]) + b1.y);
            pk.w = pack2<true>(__uint_as_float(v[h2][q * 8 + 6]) + b1.z, __uint_as_float(v[h2][q * 8 + 7]) + b1.w);
            sts128(stage + (uint32_t)row * 128u + ((uint32_t)((h2 * 4 + q) ^ (row & 7)) << 4), pk);
          }
        if (p.out_map) fence_proxy_async_smem();
        named_bar_sync(2 + 2 * grp, 128);
        // (a) the stem output itself: one TMA store of the staged tile, or 16 rows x 1 KB with lanes
        // along the bytes of a row
        uint16_t* obase = reinterpret_cast<uint16_t*>(p.out) + (((int64_t)n * p.H + Y0) * p.W + X0) * 64;
        if (p.out_map) {
          if (et == 0) {
            tma_store_5d(p.out_map, pool_stage + grp * 16384, 0, X0, 0, Y0, n);
            tma_store_commit();
          }
        } else
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int idx = et + 128 * k;
          const int rp = idx >> 3, ch = idx & 7;  // tile pixel (row-major, 8 per image row), 16-byte chunk
          const uint4 val = lds128(stage + (uint32_t)rp * 128u + ((uint32_t)(ch ^ (rp & 7)) << 4));
          *reinterpret_cast<uint4*>(obase + ((int64_t)(rp >> 3) * p.W + (rp & 7)) * 64 + ch * 8) = val;
        }
        // (b) pooled partial maxima: 9 x 5 pooled pixels touch this tile
        for (int i = et; i < 9 * 5 * 8; i += 128) {
          const int ch = i & 7, pp = i >> 3;
          const int pyl = pp / 5, pxl = pp - pyl * 5;
          const int PY = (Y0 >> 1) + pyl, PX = (X0 >> 1) + pxl;
          if (PY >= Hq || PX >= Wq) continue;
          uint4 m = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy) {
            const int ry = 2 * pyl + dy;
            if (ry < 0 || ry > 15) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
              const int rx = 2 * pxl + dx;
              if (rx < 0 || rx > 7) continue;
              const int rp = ry * 8 + rx;
              const uint4 val = lds128(stage + (uint32_t)rp * 128u + ((uint32_t)(ch ^ (rp & 7)) << 4));
              m.x = act_max2(m.x, val.x);
              m.y = act_max2(m.y, val.y);
              m.z = act_max2(m.z, val.z);
              m.w = act_max2(m.w, val.w);
            }
          }
          red_max_act8(p.pool_out + (((int64_t)n * Hq + PY) * Wq + PX) * 64 + ch * 8, m);
        }
      }
      if (p.out_map && et == 0) tma_store_wait_all<0>();
    } else
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n_tile, X0, Y0, n;
      decode(t, n_tile, X0, Y0, n);
      const int oy = Y0 + yi;
      const int64_t rowpix = ((int64_t)n * p.H + oy) * p.W;
      const int ch0 = n_tile * p.BN;
      mbar_wait(&ctl->acc_full[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(quarter * 32) << 16);
      for (int c = half * 32; c < ncols; c += 64) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c, v);
        tmem_ld_wait();
        if (oy < p.H) {
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const int col = c + g8 * 8;
            if (col >= ncols) break;
            // BN is a power of two whenever MT > 1 (MT <= 256 / BN is only granted then)
            const int mt = MT == 1 ? 0 : col >> bn_log2;
            const int ch = MT == 1 ? col : col & (p.BN - 1);
            const int ox = X0 + mt * 8 + xi;
            if (ox < p.W) {
              if (p.head.on) {
                if (ch == 0) head_fused_pixel(p.head, &v[g8 * 8], bias_s, n - p.n_base, oy, ox);
              } else {
                epilogue_group8(eo, &v[g8 * 8], bias_s, rowpix + ox, ch0 + ch);
              }
            }
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&ctl->acc_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == HALO2_LOAD_WARPS) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

size_t conv_halo2_smem_bytes(const ConvHalo2Params& p) {
  const size_t b = p.b_stages ? (size_t)p.b_stages * p.b_bytes : (size_t)p.nslabs * (p.stem ? 1 : 9) * p.b_bytes;
  return (size_t)p.a_stages * p.a_stage_bytes + b + kHaloCtlBytes + kHaloBiasBytes + 1024 +
         (p.stem ? kStemRawRing * kStemRawBytes + (p.pool_out ? 2 * 16384 : 0) : 0);
}

namespace {
template <int MODE, int KC, int MT>
cudaError_t halo2_launch_one(const ConvHalo2Params& p, int grid, size_t smem, cudaStream_t st, bool configure) {
  if (configure)
    return cudaFuncSetAttribute(conv_halo2_kernel<MODE, KC, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                227 * 1024);
  conv_halo2_kernel<MODE, KC, MT><<<grid, HALO2_THREADS, smem, st>>>(p);
  return cudaGetLastError();
}
cudaError_t halo2_dispatch(const ConvHalo2Params& p, int grid, size_t smem, cudaStream_t st, bool configure) {
  if (p.stem) return halo2_launch_one<1, 64, 1>(p, grid, smem, st, configure);
  const int key = p.kc * 10 + p.mt;
  switch (key) {
    case 641: return halo2_launch_one<0, 64, 1>(p, grid, smem, st, configure);
    case 642: return halo2_launch_one<0, 64, 2>(p, grid, smem, st, configure);
    case 321: return halo2_launch_one<0, 32, 1>(p, grid, smem, st, configure);
    case 322: return halo2_launch_one<0, 32, 2>(p, grid, smem, st, configure);
    case 324: return halo2_launch_one<0, 32, 4>(p, grid, smem, st, configure);
    case 161: return halo2_launch_one<0, 16, 1>(p, grid, smem, st, configure);
    case 162: return halo2_launch_one<0, 16, 2>(p, grid, smem, st, configure);
    case 164: return halo2_launch_one<0, 16, 4>(p, grid, smem, st, configure);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace

cudaError_t launch_conv_halo2(const ConvHalo2Params& p0, int num_sms, cudaStream_t st) {
  ConvHalo2Params p = p0;
  const int64_t total_tiles = (int64_t)p.n_tiles * p.tiles_x * p.tiles_y * p.NB;
  const int64_t dmax = p.n_tiles > p.tiles_x ? (p.n_tiles > p.tiles_y ? p.n_tiles : p.tiles_y)
                                             : (p.tiles_x > p.tiles_y ? p.tiles_x : p.tiles_y);
  if (total_tiles * dmax >= (1ll << 32)) return cudaErrorInvalidValue;  // FastDiv exactness bound
  p.div_n_tiles = make_fastdiv((uint32_t)p.n_tiles);
  p.div_tx = make_fastdiv((uint32_t)p.tiles_x);
  p.div_ty = make_fastdiv((uint32_t)p.tiles_y);
  const int grid = total_tiles < num_sms ? (int)total_tiles : num_sms;
  return halo2_dispatch(p, grid, conv_halo2_smem_bytes(p), st, false);
}

size_t conv_halo_el_smem_bytes(const ConvHaloElParams& p) {
  return (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.b_stages * p.b_bytes + (p.out_map ? 32768 : 0) + kHaloCtlBytes +
         kHaloElBiasBytes + sizeof(HaloElTables) + 1024;
}

cudaError_t launch_conv_halo_el(const ConvHaloElParams& p0, int num_sms, cudaStream_t st) {
  ConvHaloElParams p = p0;
  const int64_t total_tiles = (int64_t)p.n_tiles * p.tiles_x * p.tiles_y * p.NB;
  const int64_t dmax = p.n_tiles > p.tiles_x ? (p.n_tiles > p.tiles_y ? p.n_tiles : p.tiles_y)
                                             : (p.tiles_x > p.tiles_y ? p.tiles_x : p.tiles_y);
  if (total_tiles * dmax >= (1ll << 32)) return cudaErrorInvalidValue;  // FastDiv exactness bound
  if (p.n_slabs > HALO_EL_MAX_SLABS || p.n_entries > HALO_EL_MAX_ENTRIES || p.n_tiles > HALO_EL_MAX_NTILES ||
      p.n_tiles * p.BN > 1024 || p.mt * p.BN > 512 || conv_halo_el_smem_bytes(p) > 227 * 1024)
    return cudaErrorInvalidValue;
  p.div_n_tiles = make_fastdiv((uint32_t)p.n_tiles);
  p.div_tx = make_fastdiv((uint32_t)p.tiles_x);
  p.div_ty = make_fastdiv((uint32_t)p.tiles_y);
  const int grid = total_tiles < num_sms ? (int)total_tiles : num_sms;
  if (p.mt == 2) conv_halo_el_kernel<2><<<grid, HALO_THREADS, conv_halo_el_smem_bytes(p), st>>>(p);
  else conv_halo_el_kernel<1><<<grid, HALO_THREADS, conv_halo_el_smem_bytes(p), st>>>(p);
  return cudaGetLastError();
}

size_t conv_halo_smem_bytes(const ConvHaloParams& p) {
  const size_t b = p.b_stages ? (size_t)p.b_stages * p.b_bytes : (size_t)p.ncs * 9 * p.b_bytes;
  return (size_t)p.a_stages * p.a_stage_bytes + b + (size_t)(p.out_bufs + (p.res_inplace ? 0 : p.res_bufs)) * p.out_buf_bytes +
         kHaloCtlBytes + kHaloBiasBytes + 1024;
}

cudaError_t conv_halo_configure() {
  cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_halo_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_halo_el_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_halo_el_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  ConvHalo2Params q{};
  q.stem = 1;
  if (e == cudaSuccess) e = halo2_dispatch(q, 0, 0, nullptr, true);
  q.stem = 0;
  for (int kc : {16, 32, 64})
    for (int mt : {1, 2, 4}) {
      if (kc == 64 && mt == 4) continue;
      q.kc = kc;
      q.mt = mt;
      if (e == cudaSuccess) e = halo2_dispatch(q, 0, 0, nullptr, true);
    }
  return e;
}

cudaError_t launch_conv_halo(const ConvHaloParams& p0, int num_sms, cudaStream_t st) {
  ConvHaloParams p = p0;
  const int64_t total_tiles = (int64_t)p.n_tiles * p.tiles_x * p.tiles_y * p.NB;
  const int64_t dmax = p.n_tiles > p.tiles_x ? (p.n_tiles > p.tiles_y ? p.n_tiles : p.tiles_y)
                                             : (p.tiles_x > p.tiles_y ? p.tiles_x : p.tiles_y);
  if (total_tiles * dmax >= (1ll << 32)) return cudaErrorInvalidValue;  // FastDiv exactness bound
  p.div_n_tiles = make_fastdiv((uint32_t)p.n_tiles);
  p.div_tx = make_fastdiv((uint32_t)p.tiles_x);
  p.div_ty = make_fastdiv((uint32_t)p.tiles_y);
  const int grid = total_tiles < num_sms ? (int)total_tiles : num_sms;
  if (p.kc == 32) conv_halo_kernel<32><<<grid, HALO_THREADS, conv_halo_smem_bytes(p), st>>>(p);
  else conv_halo_kernel<64><<<grid, HALO_THREADS, conv_halo_smem_bytes(p), st>>>(p);
  return cudaGetLastError();
}

}  // namespace vsb
