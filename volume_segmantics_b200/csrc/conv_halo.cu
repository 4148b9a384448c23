// Halo tcgen05 convolution kernel (see conv_halo.cuh).  Persistent, one CTA per SM:
//   warp 0      A producer: one TMA halo box per (tile, 64-channel slab)
//   warp 1      TMEM allocator + MMA issuer: 9 taps x 4 K-steps per slab
//   warp 2      B producer: weight images, once (resident) or through a ring
//   warps 3..10 epilogue (TMEM -> registers -> bias/residual/ReLU -> NHWC)
#include "conv_halo.cuh"

#include "conv_epilogue.cuh"

namespace vsb {

namespace {

struct HaloCtl {
  uint64_t a_full[HALO_MAX_A_STAGES];
  uint64_t a_empty[HALO_MAX_A_STAGES];
  uint64_t b_full[HALO_MAX_B_STAGES];
  uint64_t b_empty[HALO_MAX_B_STAGES];
  uint64_t w_full;
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  uint32_t pad[3];
};
constexpr int kHaloCtlBytes = 1024;
constexpr int kHaloBiasBytes = 2048 * 4;

__device__ __forceinline__ void halo_decode(const ConvHaloParams& p, int t, int& n_tile, int& X0, int& Y0,
                                            int& n) {
  n_tile = t % p.n_tiles;
  int sp = t / p.n_tiles;
  const int tx = sp % p.tiles_x;
  sp /= p.tiles_x;
  const int ty = sp % p.tiles_y;
  n = sp / p.tiles_y;
  X0 = tx * 8;
  Y0 = ty * 16;
}


// Descriptor + small offset (16-byte units).  The 14-bit address field lives in the low
// word and never carries out for shared-memory addresses, so a 32-bit add suffices.
__device__ __forceinline__ uint64_t desc_add(uint64_t d, uint32_t off16) {
  return (d & 0xffffffff00000000ull) | (uint32_t)((uint32_t)d + off16);
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

}  // namespace

__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ ConvHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_area = smem + (size_t)p.a_stages * p.a_stage_bytes;
  const size_t b_area_bytes = p.b_stages ? (size_t)p.b_stages * p.b_bytes : (size_t)p.ncs * 9 * p.b_bytes;
  HaloCtl* ctl = reinterpret_cast<HaloCtl*>(b_area + b_area_bytes);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctl) + kHaloCtlBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.n_tiles * p.tiles_x * p.tiles_y * p.NB;
  const int HW = 8 + 2 * p.dil, HH = 16 + 2 * p.dil;
  const bool resident = p.b_stages == 0;

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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->acc_full[i], 1);
      mbar_init(&ctl->acc_empty[i], 32 * HALO_EPI_WARPS);
    }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < p.n_tiles * p.BN; i += HALO_THREADS) bias_s[i] = p.bias[i];
  if (warp == 1) tmem_alloc<512>(&ctl->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================== A producer: halo boxes =====================
    int as = 0;
    uint32_t aph = 0;
    const uint32_t a_box_bytes = (uint32_t)(HW * HH * 128);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n_tile, X0, Y0, n;
      halo_decode(p, t, n_tile, X0, Y0, n);
      for (int cs = 0; cs < p.ncs; ++cs) {
        mbar_wait(&ctl->a_empty[as], aph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&ctl->a_full[as], a_box_bytes);
          tma_load_5d(p.map, &ctl->a_full[as], a_ring + (size_t)as * p.a_stage_bytes, cs * 64, X0 - p.dil, 0,
                      Y0 - p.dil, n);
        }
        __syncwarp();
        if (++as == p.a_stages) {
          as = 0;
          aph ^= 1;
        }
      }
    }
  } else if (warp == 2) {
    // ===================== B producer: weight images =====================
    if (resident) {
      if (elect_one()) {
        const uint32_t total = (uint32_t)(p.ncs * 9 * p.b_bytes);
        mbar_arrive_expect_tx(&ctl->w_full, total);
        for (int i = 0; i < p.ncs * 9; ++i)
          bulk_load_1d(b_area + (size_t)i * p.b_bytes, p.wpacked + (size_t)i * p.b_bytes, (uint32_t)p.b_bytes,
                       &ctl->w_full);
      }
      __syncwarp();
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
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int as = 0, bs = 0, acc = 0;
    uint32_t aph = 0, bph = 0, acc_phase = 0;
    const uint32_t idesc = umma_idesc_act(128, p.BN);
    const uint32_t sbo = (uint32_t)HW * 128;
    const uint64_t a_desc0 = umma_smem_desc_sw128(smem_u32(a_ring), sbo);
    const uint64_t b_desc0 = umma_smem_desc_sw128(smem_u32(b_area), 1024);
    const uint32_t a_step = (uint32_t)p.a_stage_bytes >> 4;
    const uint32_t b_step = (uint32_t)p.b_bytes >> 4;
    const uint32_t row_units = (uint32_t)(p.dil * HW * 8);  // one dilated halo row, in 16-byte units
    const uint32_t col_units = (uint32_t)(p.dil * 8);
    if (resident) mbar_wait(&ctl->w_full, 0);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      mbar_wait(&ctl->acc_empty[acc], acc_phase ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
      for (int cs = 0; cs < p.ncs; ++cs) {
        mbar_wait(&ctl->a_full[as], aph);
        tc_fence_after_sync();
        const uint64_t a_stage_desc = desc_add(a_desc0, as * a_step);
        if (resident) {
          if (elect_one()) {
            issue_slab_resident<9, 4>(d_tmem, a_stage_desc, desc_add(b_desc0, cs * 9 * b_step), b_step, row_units,
                                      col_units, idesc, cs != 0);
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
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ss(d_tmem, desc_add(at, 2 * k), desc_add(bt, 2 * k), idesc,
                             (tap | k) != 0 ? 1u : (cs != 0 ? 1u : 0u));
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
        if (++as == p.a_stages) {
          as = 0;
          aph ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue =====================
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
      const int ox = X0 + xi, oy = Y0 + yi;
      const bool valid = ox < p.W && oy < p.H;
      const int64_t pix = ((int64_t)n * p.H + oy) * p.W + ox;
      const int ch0 = n_tile * p.BN;
      mbar_wait(&ctl->acc_full[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(quarter * 32) << 16);
      for (int c = half * 32; c < p.BN; c += 64) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c, v);
        tmem_ld_wait();
        if (valid) epilogue_chunk32(eo, v, bias_s, pix, ch0 + c);
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
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// =====================================================================================
// Generalised halo kernel: cp.async-assembled A tiles (see conv_halo.cuh).
//   warps 0..3   A loaders (128 threads): 16-byte cp.async copies, zero-filled outside
//                the image, written at the SW128-swizzled position of a 128-byte-pitch
//                row; a stage is published after cp.async.wait_group + fence.proxy.async
//   warp 4       MMA issuer + TMEM allocator
//   warp 5       B producer (weights resident or ring, as above)
//   warps 6..13  epilogue
// =====================================================================================
namespace {
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void halo2_decode(const ConvHalo2Params& p, int t, int& n_tile, int& X0, int& Y0,
                                             int& n) {
  n_tile = t % p.n_tiles;
  int sp = t / p.n_tiles;
  const int tx = sp % p.tiles_x;
  sp /= p.tiles_x;
  const int ty = sp % p.tiles_y;
  n = sp / p.tiles_y;
  X0 = tx * 8;
  Y0 = ty * 16;
}
}  // namespace

// MODE 0: 3x3 halo convolution.  MODE 1: the 7x7 stride-2 single-channel stem
// (smp ResNetEncoder conv1): the loader writes the im2col row of each output pixel
// (49 taps, zero padded to K = 64) and one tap of four K-steps is issued per tile.
template <int MODE, int KS>
__global__ void __launch_bounds__(HALO2_THREADS, 1)
conv_halo2_kernel(const __grid_constant__ ConvHalo2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_area = smem + (size_t)p.a_stages * p.a_stage_bytes;
  const size_t b_area_bytes =
      p.b_stages ? (size_t)p.b_stages * p.b_bytes : (size_t)p.nslabs * (MODE == 0 ? 9 : 1) * p.b_bytes;
  HaloCtl* ctl = reinterpret_cast<HaloCtl*>(b_area + b_area_bytes);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctl) + kHaloCtlBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.n_tiles * p.tiles_x * p.tiles_y * p.NB;
  constexpr int HW = 10, HH = 18;
  constexpr int NTAPS = MODE == 0 ? 9 : 1;
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
      mbar_init(&ctl->acc_empty[i], 32 * 8);
    }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < p.n_tiles * p.BN; i += HALO2_THREADS) bias_s[i] = p.bias[i];
  if (warp == HALO2_LOAD_WARPS) tmem_alloc<512>(&ctl->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp < HALO2_LOAD_WARPS) {
    // ===================== A loaders =====================
    const int ptid = threadIdx.x;  // 0..127
    int as = 0;
    uint32_t aph = 0;
    // Up to `depth` cp.async groups (= stages) stay in flight per thread; the oldest is
    // published (wait_group -> fence.proxy.async -> arrive) once `depth` newer ones exist.
    // depth = a_stages / 2 leaves the other half of the ring published ahead of the MMA
    // warp (depth = a_stages - 1 would run loader and MMA in lock-step).
    const int depth = p.a_stages / 2 > 1 ? p.a_stages / 2 : 1;
    int inflight = 0, oldest = 0;
    const uint32_t ring_addr = smem_u32(a_ring);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n_tile, X0, Y0, n;
      halo2_decode(p, t, n_tile, X0, Y0, n);
      if (MODE == 1) {
        const HaloSrc& sv = p.src[0];
        const uint16_t* img = sv.ptr + (int64_t)n * sv.Hs * sv.Ws;
        const int m = ptid, ox = X0 + (m & 7), oy = Y0 + (m >> 3);
        uint32_t vals[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) vals[i] = 0;
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
          const int iy = 2 * oy + ky - 3;
          const bool oky = iy >= 0 && iy < sv.Hs;
#pragma unroll
          for (int kx = 0; kx < 7; ++kx) {
            const int ix = 2 * ox + kx - 3;
            const int k = ky * 7 + kx;
            uint32_t v = 0;
            if (oky && ix >= 0 && ix < sv.Ws) v = __ldg(img + (int64_t)iy * sv.Ws + ix);
            vals[k >> 1] |= v << (16 * (k & 1));
          }
        }
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
      for (int s = 0; s < p.nslabs; ++s) {
        const HaloSrc& sv = p.src[p.slab_src[s]];
        const int c0 = p.slab_c0[s];
        const int ch8_log2 = p.slab_kc[s] == 64 ? 3 : (p.slab_kc[s] >= 32 ? 2 : 1);  // 16-B chunks per pixel
        const int ch8 = p.slab_kc[s] >> 3;
        mbar_wait(&ctl->a_empty[as], aph ^ 1);
        const uint32_t stage_addr = ring_addr + (uint32_t)as * p.a_stage_bytes;
        const uint16_t* img = sv.ptr + (int64_t)n * sv.Hs * sv.Ws * sv.C + c0;
        if (ch8 == (1 << ch8_log2)) {
          const int total = (HW * HH) << ch8_log2;
          for (int idx = ptid; idx < total; idx += 32 * HALO2_LOAD_WARPS) {
            const int j = idx & (ch8 - 1);
            const int pix = idx >> ch8_log2;
            const int hy = pix / HW, hx = pix - hy * HW;
            const int y = Y0 - 1 + hy, x = X0 - 1 + hx;
            const bool ok = y >= 0 && y < p.H && x >= 0 && x < p.W;
            const int sy = sv.up ? y >> 1 : y, sx = sv.up ? x >> 1 : x;
            const uint16_t* g = img + ((int64_t)sy * sv.Ws + sx) * sv.C + j * 8;
            cp_async_16(stage_addr + pix * 128 + ((j ^ (pix & 7)) << 4), ok ? g : sv.ptr, ok ? 16u : 0u);
          }
        } else {  // 48 channels: 6 chunks per pixel
          const int total = HW * HH * ch8;
          for (int idx = ptid; idx < total; idx += 32 * HALO2_LOAD_WARPS) {
            const int pix = idx / ch8, j = idx - pix * ch8;
            const int hy = pix / HW, hx = pix - hy * HW;
            const int y = Y0 - 1 + hy, x = X0 - 1 + hx;
            const bool ok = y >= 0 && y < p.H && x >= 0 && x < p.W;
            const int sy = sv.up ? y >> 1 : y, sx = sv.up ? x >> 1 : x;
            const uint16_t* g = img + ((int64_t)sy * sv.Ws + sx) * sv.C + j * 8;
            cp_async_16(stage_addr + pix * 128 + ((j ^ (pix & 7)) << 4), ok ? g : sv.ptr, ok ? 16u : 0u);
          }
        }
        cp_async_commit();
        if (++inflight > depth) {
          switch (depth) {  // all but the `depth` most recent groups have landed
            case 1: cp_async_wait<1>(); break;
            case 2: cp_async_wait<2>(); break;
            case 3: cp_async_wait<3>(); break;
            case 4: cp_async_wait<4>(); break;
            case 5: cp_async_wait<5>(); break;
            case 6: cp_async_wait<6>(); break;
            default: cp_async_wait<7>(); break;
          }
          fence_proxy_async_smem();
          mbar_arrive(&ctl->a_full[oldest]);
          if (++oldest == p.a_stages) oldest = 0;
          --inflight;
        }
        if (++as == p.a_stages) {
          as = 0;
          aph ^= 1;
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    while (inflight-- > 0) {
      mbar_arrive(&ctl->a_full[oldest]);
      if (++oldest == p.a_stages) oldest = 0;
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
  } else if (warp == HALO2_LOAD_WARPS) {
    // ===================== MMA issuer =====================
    int as = 0, bs = 0, acc = 0;
    uint32_t aph = 0, bph = 0, acc_phase = 0;
    const uint32_t idesc = umma_idesc_act(128, p.BN);
    const uint64_t a_desc0 = umma_smem_desc_sw128(smem_u32(a_ring), MODE == 0 ? HW * 128 : 1024);
    const uint64_t b_desc0 = umma_smem_desc_sw128(smem_u32(b_area), 1024);
    const uint32_t a_step = (uint32_t)p.a_stage_bytes >> 4;
    const uint32_t b_step = (uint32_t)p.b_bytes >> 4;
    if (resident) mbar_wait(&ctl->w_full, 0);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      mbar_wait(&ctl->acc_empty[acc], acc_phase ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
      for (int s = 0; s < p.nslabs; ++s) {
        const int ksteps = KS ? KS : (p.slab_kc[s] >> 4);
        mbar_wait(&ctl->a_full[as], aph);
        tc_fence_after_sync();
        const uint64_t a_stage_desc = desc_add(a_desc0, as * a_step);
        if (resident && KS != 0) {
          if (elect_one()) {
            issue_slab_resident<NTAPS, KS ? KS : 1>(d_tmem, a_stage_desc, desc_add(b_desc0, s * NTAPS * b_step),
                                                    b_step, HW * 8, 8, idesc, s != 0);
            umma_commit(&ctl->a_empty[as]);
            if (s == p.nslabs - 1) umma_commit(&ctl->acc_full[acc]);
          }
          __syncwarp();
        } else {
#pragma unroll
          for (int tap = 0; tap < NTAPS; ++tap) {
            uint64_t bt;
            if (resident) {
              bt = desc_add(b_desc0, (s * NTAPS + tap) * b_step);
            } else {
              mbar_wait(&ctl->b_full[bs], bph);
              tc_fence_after_sync();
              bt = desc_add(b_desc0, bs * b_step);
            }
            if (elect_one()) {
              const uint64_t at = desc_add(a_stage_desc, (tap / 3) * (HW * 8) + (tap % 3) * 8);
              if (KS != 0) {
#pragma unroll
                for (int k = 0; k < (KS ? KS : 1); ++k)
                  umma_bf16_ss(d_tmem, desc_add(at, 2 * k), desc_add(bt, 2 * k), idesc,
                               (tap | k) != 0 ? 1u : (s != 0 ? 1u : 0u));
              } else {
                for (int k = 0; k < ksteps; ++k)
                  umma_bf16_ss(d_tmem, desc_add(at, 2 * k), desc_add(bt, 2 * k), idesc, (s | tap | k) != 0);
              }
              if (!resident) umma_commit(&ctl->b_empty[bs]);
              if (tap == NTAPS - 1) {
                umma_commit(&ctl->a_empty[as]);
                if (s == p.nslabs - 1) umma_commit(&ctl->acc_full[acc]);
              }
            }
            __syncwarp();
            if (!resident && ++bs == p.b_stages) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
        if (++as == p.a_stages) {
          as = 0;
          aph ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue =====================
    int acc = 0;
    uint32_t acc_phase = 0;
    const int quarter = warp & 3;
    const int half = (warp - (HALO2_LOAD_WARPS + 2)) >> 2;
    const int row = quarter * 32 + lane;
    const int xi = row & 7, yi = row >> 3;
    const EpiOut eo{p.out, p.residual, p.out_f32, p.relu, p.cout};
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n_tile, X0, Y0, n;
      halo2_decode(p, t, n_tile, X0, Y0, n);
      const int ox = X0 + xi, oy = Y0 + yi;
      const bool valid = ox < p.W && oy < p.H;
      const int64_t pix = ((int64_t)n * p.H + oy) * p.W + ox;
      const int ch0 = n_tile * p.BN;
      mbar_wait(&ctl->acc_full[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(quarter * 32) << 16);
      for (int c = half * 32; c < p.BN; c += 64) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c, v);
        tmem_ld_wait();
        if (valid) epilogue_chunk32(eo, v, bias_s, pix, ch0 + c);
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
  return (size_t)p.a_stages * p.a_stage_bytes + b + kHaloCtlBytes + kHaloBiasBytes + 1024;
}

cudaError_t launch_conv_halo2(const ConvHalo2Params& p, int num_sms, cudaStream_t st) {
  const int total_tiles = p.n_tiles * p.tiles_x * p.tiles_y * p.NB;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  const size_t smem = conv_halo2_smem_bytes(p);
  int ks = p.slab_kc[0] >> 4;  // uniform K-steps per slab, else 0 (runtime)
  for (int s = 1; s < p.nslabs; ++s)
    if ((p.slab_kc[s] >> 4) != ks) ks = 0;
  if (p.stem) conv_halo2_kernel<1, 4><<<grid, HALO2_THREADS, smem, st>>>(p);
  else if (ks == 4) conv_halo2_kernel<0, 4><<<grid, HALO2_THREADS, smem, st>>>(p);
  else if (ks == 2) conv_halo2_kernel<0, 2><<<grid, HALO2_THREADS, smem, st>>>(p);
  else if (ks == 1) conv_halo2_kernel<0, 1><<<grid, HALO2_THREADS, smem, st>>>(p);
  else conv_halo2_kernel<0, 0><<<grid, HALO2_THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

size_t conv_halo_smem_bytes(const ConvHaloParams& p) {
  const size_t b = p.b_stages ? (size_t)p.b_stages * p.b_bytes : (size_t)p.ncs * 9 * p.b_bytes;
  return (size_t)p.a_stages * p.a_stage_bytes + b + kHaloCtlBytes + kHaloBiasBytes + 1024;
}

cudaError_t conv_halo_configure() {
  cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_halo2_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_halo2_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_halo2_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_halo2_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_halo2_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  return e;
}

cudaError_t launch_conv_halo(const ConvHaloParams& p, int num_sms, cudaStream_t st) {
  const int total_tiles = p.n_tiles * p.tiles_x * p.tiles_y * p.NB;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  conv_halo_kernel<<<grid, HALO_THREADS, conv_halo_smem_bytes(p), st>>>(p);
  return cudaGetLastError();
}

}  // namespace vsb
