// Stem of the ResNet encoders: conv 7x7 / stride 2 / pad 3, 1 -> 64 channels, folded BN + ReLU,
// FUSED with the 3x3 / stride 2 / pad 1 max-pool that follows it (smp ResNetEncoder conv1 + bn1 +
// relu + maxpool [ext]; reached from vol_seg_2d_predictor.py:44).  HBM-bound: per conv output
// pixel 4 x 2 B in, 128 B out (the /2 skip tensor), 32 B pooled.
//
// A CTA owns a block of 8 x 7 POOLED pixels.  The pooling windows of that block cover conv rows
// [2*py0 - 1, 2*py0 + 16) and conv columns [2*px0 - 1, 2*px0 + 14): 17 x 15 = 255 conv pixels = the
// 256 rows of two M = 128 tcgen05 tiles (one dummy row).  The one conv row / column above / left of
// the 16 x 14 pixels this CTA OWNS is recomputed (14 % more MMA rows, which are not the limit) so the
// pool needs no exchange between CTAs, no atomics and no zero-initialised output:
//   warps 0..3   loaders: raw input window (39 rows x 48 px, 16-byte cp.async, zero-filled outside
//                the image, two tiles ahead) -> im2col rows in the SW128 K-major A layout:
//                K chunk ky (16 bytes) = input columns 2*cx-4 .. 2*cx+3 of row 2*cy+ky-3; column
//                2*cx-4 lies outside the 7x7 window and meets a zero weight; K padded 49 -> 64
//   warp 4       TMEM allocator + MMA issuer: 2 tiles x 4 K-steps, N = 64, accumulators in 4 stages
//   warp 5       weights (8 KB, once), then TMA-store issuer: the owned 16 x 14 conv pixels and the
//                8 x 7 pooled pixels leave as one cp.async.bulk.tensor store each (clipped at the
//                image edge by TMA)
//   warps 6..13  epilogue: tcgen05.ld (thread = conv pixel, 64 channels), + bias, ReLU, 16-bit pack
//                -> swizzled staging tile; then the 3x3 max over the staged tile -> pooled staging.
// Pixels outside the image (the halo row / column of border tiles, the dummy row) are staged as 0,
// the identity of max over ReLU outputs -- every pooling window holds at least one real pixel.
#include "conv_halo.cuh"

#include "conv_epilogue.cuh"
#include "kernels.h"

namespace vsb {

namespace {

constexpr int ST_PH = 8, ST_PW = 7;                  // pooled block
constexpr int ST_RH = 2 * ST_PH + 1, ST_RW = 2 * ST_PW + 1;  // conv region 17 x 15
constexpr int ST_OH = 2 * ST_PH, ST_OW = 2 * ST_PW;  // owned conv block 16 x 14
constexpr int ST_RAW_ROWS = 2 * (ST_RH - 1) + 7;     // 39 input rows
constexpr int ST_RAW_PITCH = 96;                     // bytes per raw row: 48 fp16 pixels
constexpr int ST_RAW_BYTES = 3840;                   // >= 39 * 96, multiple of 128
constexpr int ST_RAW_RING = 4;
constexpr int ST_A_STAGE = 2 * 128 * 128;            // two M tiles of 128 rows x 128 B
constexpr int ST_OWNED_BYTES = ST_OH * ST_OW * 128;  // 28672
constexpr int ST_HALO_BYTES = 4096;                  // 31 halo pixels x 128 B
constexpr int ST_POOL_BYTES = ST_PH * ST_PW * 128;   // 7168
constexpr int ST_OUT_BUF = ST_OWNED_BYTES + ST_HALO_BYTES + ST_POOL_BYTES;  // 39936 (multiple of 1024)
constexpr int ST_LOAD_WARPS = 4, ST_EPI_WARPS = 8;
constexpr int ST_THREADS = 32 * (ST_LOAD_WARPS + 2 + ST_EPI_WARPS);
constexpr int ST_MAX_A = 4;

struct StemCtl {
  uint64_t a_full[ST_MAX_A];
  uint64_t a_empty[ST_MAX_A];
  uint64_t w_full;
  uint64_t acc_full[4];
  uint64_t acc_empty[4];
  uint64_t out_full[2];
  uint64_t out_empty[2];
  uint64_t staged_full[2];  // v3: staging warps -> pooling warps
  uint32_t tmem_base;
  uint32_t pad[3];
};

__device__ __forceinline__ void st_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint4 st_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t st_max2(uint32_t a, uint32_t b) {
  uint32_t r;
#if VSB_ACT_F16
  asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
#else
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
#endif
  return r;
}
__device__ __forceinline__ uint64_t st_desc(uint32_t addr) {  // SW128 K-major, 8-row groups 1024 B apart
  return (uint64_t)((addr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// byte offset of chunk `j` (16 B) of conv-region pixel (ry, rx) inside one staging buffer
__device__ __forceinline__ uint32_t st_px_off(int ry, int rx) {
  if (ry >= 1 && rx >= 1) return (uint32_t)((ry - 1) * ST_OW + (rx - 1)) * 128u;              // owned block, TMA box order
  return (uint32_t)ST_OWNED_BYTES + (uint32_t)(ry == 0 ? rx : ST_RW - 1 + ry) * 128u;          // halo row 0, then halo column 0
}
__device__ __forceinline__ uint32_t st_chunk(uint32_t px_off, int j) {  // Swizzle<3,4,3> on the address bits
  return px_off + ((uint32_t)(j ^ (int)((px_off >> 7) & 7u)) << 4);
}

}  // namespace

__global__ void __launch_bounds__(ST_THREADS, 1) stem_pool_kernel(const __grid_constant__ ConvStemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_ring = smem;                                                   // a_stages x 32 KB
  uint8_t* out_stage = a_ring + (size_t)p.a_stages * ST_A_STAGE;            // 2 x 39 KB
  uint8_t* b_area = out_stage + 2 * ST_OUT_BUF;                             // 8 KB weights
  uint8_t* raw_ring = b_area + 64 * 128;                                    // 4 x 3840 B
  StemCtl* ctl = reinterpret_cast<StemCtl*>(raw_ring + ST_RAW_RING * ST_RAW_BYTES);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctl) + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_x * p.tiles_y * p.NB;
  const int Hc = p.H, Wc = p.W;            // conv output (= stem tensor) size
  const int Hin = 2 * Hc, Win = 2 * Wc;    // network input size

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(&ctl->a_full[i], 32 * ST_LOAD_WARPS);
      mbar_init(&ctl->a_empty[i], 1);
    }
    mbar_init(&ctl->w_full, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&ctl->acc_full[i], 1);
      mbar_init(&ctl->acc_empty[i], ST_EPI_WARPS);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->out_full[i], ST_EPI_WARPS);
      mbar_init(&ctl->out_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (threadIdx.x < 64) bias_s[threadIdx.x] = p.bias[threadIdx.x];
  if (warp == ST_LOAD_WARPS) tmem_alloc<512>(&ctl->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  auto decode = [&](int t, int& px0, int& py0, int& n) {
    auto fdiv = [](uint32_t v, const FastDiv& f) { return f.m ? __umulhi(v, f.m) : v; };
    const uint32_t q = fdiv((uint32_t)t, p.div_tx);
    const int tx = (int)((uint32_t)t - q * p.div_tx.d);
    const uint32_t q2 = fdiv(q, p.div_ty);
    const int ty = (int)(q - q2 * p.div_ty.d);
    n = (int)q2 + p.n_base;
    px0 = tx * ST_PW;
    py0 = ty * ST_PH;
  };

  if (warp < ST_LOAD_WARPS) {
    // ===================== loaders =====================
    const int ptid = threadIdx.x;  // 0..127
    const uint32_t raw0 = smem_u32(raw_ring), ring_addr = smem_u32(a_ring);
    auto fetch_raw = [&](int tt, int slot) {
      if (tt < total_tiles) {
        int px2, py2, n2;
        decode(tt, px2, py2, n2);
        const uint16_t* img = p.in + (int64_t)n2 * Hin * Win;
        const int iy0 = 4 * py2 - 5, ix0 = (4 * px2 - 8) & ~7;  // 16-byte aligned window origin
        const uint32_t dst0 = raw0 + (uint32_t)slot * ST_RAW_BYTES;
        for (int i = ptid; i < ST_RAW_ROWS * 6; i += 32 * ST_LOAD_WARPS) {
          const int rr = i / 6, cc = i - rr * 6;
          const int iy = iy0 + rr, ix = ix0 + 8 * cc;
          const bool ok = iy >= 0 && iy < Hin && ix >= 0 && ix < Win;  // Win % 8 == 0: a chunk is all in or all out
          const uint16_t* g = ok ? img + (int64_t)iy * Win + ix : p.in;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + rr * ST_RAW_PITCH + cc * 16), "l"(g),
                       "r"(ok ? 16u : 0u)
                       : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // this thread builds row `ptid` of both M tiles: conv-region pixels q = ptid and q = 128 + ptid
    int ryA, rxA, ryB, rxB;
    ryA = ptid / ST_RW;
    rxA = ptid - ryA * ST_RW;
    ryB = (128 + ptid) / ST_RW;
    rxB = (128 + ptid) - ryB * ST_RW;
    const bool dummyB = 128 + ptid >= ST_RH * ST_RW;
    int as = 0;
    uint32_t aph = 0;
    int k = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++k) {
      if (k == 0) {
        fetch_raw(t, 0);
        fetch_raw(t + (int)gridDim.x, 1);
      }
      fetch_raw(t + 2 * (int)gridDim.x, (k + 2) % ST_RAW_RING);
      asm volatile("cp.async.wait_group 2;" ::: "memory");  // this tile's window has landed
      st_bar_sync(5, 32 * ST_LOAD_WARPS);                    // ... for every loader thread
      int px0, py0, n;
      decode(t, px0, py0, n);
      const int xoff = (4 * px0 - 8) - ((4 * px0 - 8) & ~7);  // 0 or 4 pixels
      const uint32_t raw = raw0 + (uint32_t)(k % ST_RAW_RING) * ST_RAW_BYTES + (uint32_t)(xoff + 2) * 2u;
      uint32_t va[28], vb[28];
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
        const uint32_t ra = raw + (uint32_t)(2 * ryA + ky) * ST_RAW_PITCH + (uint32_t)rxA * 4u;
        const uint32_t rb = raw + (uint32_t)(2 * ryB + ky) * ST_RAW_PITCH + (uint32_t)rxB * 4u;
#pragma unroll
        for (int wq = 0; wq < 4; ++wq) {
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(va[ky * 4 + wq]) : "r"(ra + wq * 4));
          if (!dummyB) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(vb[ky * 4 + wq]) : "r"(rb + wq * 4));
          else vb[ky * 4 + wq] = 0u;
        }
      }
      mbar_wait(&ctl->a_empty[as], aph ^ 1);
      const uint32_t rowA = ring_addr + (uint32_t)as * ST_A_STAGE + (uint32_t)ptid * 128u;
      const uint32_t rowB = rowA + 128u * 128u;
      const int sw = ptid & 7;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        st_sts128(rowA + ((uint32_t)(j ^ sw) << 4), make_uint4(va[4 * j], va[4 * j + 1], va[4 * j + 2], va[4 * j + 3]));
        st_sts128(rowB + ((uint32_t)(j ^ sw) << 4), make_uint4(vb[4 * j], vb[4 * j + 1], vb[4 * j + 2], vb[4 * j + 3]));
      }
      st_sts128(rowA + ((uint32_t)(7 ^ sw) << 4), make_uint4(0u, 0u, 0u, 0u));
      st_sts128(rowB + ((uint32_t)(7 ^ sw) << 4), make_uint4(0u, 0u, 0u, 0u));
      fence_proxy_async_smem();
      mbar_arrive(&ctl->a_full[as]);
      if (++as == p.a_stages) {
        as = 0;
        aph ^= 1;
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp == ST_LOAD_WARPS) {
    // ===================== MMA issuer =====================
    int as = 0, acc = 0;
    uint32_t aph = 0, acc_phase = 0;
    const uint32_t idesc = umma_idesc_act(128, 64);
    const uint64_t a_desc0 = st_desc(smem_u32(a_ring));
    const uint64_t b_desc0 = st_desc(smem_u32(b_area));
    mbar_wait(&ctl->w_full, 0);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      mbar_wait(&ctl->acc_empty[acc], acc_phase ^ 1);
      mbar_wait(&ctl->a_full[as], aph);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t d0 = tmem_base + (uint32_t)acc * 128u;
        const uint32_t a_units = (uint32_t)(as * ST_A_STAGE) >> 4;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          const uint64_t ad = (a_desc0 & 0xffffffff00000000ull) | (uint32_t)((uint32_t)a_desc0 + a_units + m * (16384u >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16_ss(d0 + (uint32_t)m * 64u, (ad & 0xffffffff00000000ull) | (uint32_t)((uint32_t)ad + 2 * ks),
                         (b_desc0 & 0xffffffff00000000ull) | (uint32_t)((uint32_t)b_desc0 + 2 * ks), idesc, ks != 0 ? 1u : 0u);
        }
        umma_commit(&ctl->a_empty[as]);
        umma_commit(&ctl->acc_full[acc]);
      }
      __syncwarp();
      if (++as == p.a_stages) {
        as = 0;
        aph ^= 1;
      }
      if (++acc == 4) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp == ST_LOAD_WARPS + 1) {
    // ===================== weights, then TMA-store issuer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(&ctl->w_full, 64 * 128);
      bulk_load_1d(b_area, p.wpacked, 64 * 128, &ctl->w_full);
      int ob = 0;
      uint32_t oph = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int px0, py0, n;
        decode(t, px0, py0, n);
        mbar_wait(&ctl->out_full[ob], oph);
        uint8_t* buf = out_stage + (size_t)ob * ST_OUT_BUF;
        tma_store_5d(p.out_map, buf, 0, 2 * px0, 0, 2 * py0, n);
        tma_store_5d(p.pool_map, buf + ST_OWNED_BYTES + ST_HALO_BYTES, 0, px0, 0, py0, n);
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(&ctl->out_empty[ob]);
        if (++ob == 2) {
          ob = 0;
          oph ^= 1;
        }
      }
      tma_store_wait_all<0>();
    }
  } else {
    // ===================== epilogue + pool =====================
    const int ew = warp - (ST_LOAD_WARPS + 2);     // 0..7
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may read
    const int mtile = ew >> 2;                     // warps 6..9 -> tile 0, 10..13 -> tile 1 (6 & 3 = 2: quarters differ per warp, all four covered)
    const int q = mtile * 128 + quarter * 32 + lane;  // conv-region pixel of this thread
    const int ry = q / ST_RW, rx = q - ry * ST_RW;
    const bool dummy = q >= ST_RH * ST_RW;
    const uint32_t my_off = dummy ? 0u : st_px_off(ry, rx);
    const int et = ew * 32 + lane;  // 0..255
    const uint32_t stage0 = smem_u32(out_stage);
    int acc = 0, ob = 0;
    uint32_t acc_phase = 0, oph = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int px0, py0, n;
      decode(t, px0, py0, n);
      const int cy = 2 * py0 - 1 + ry, cx = 2 * px0 - 1 + rx;
      const bool valid = !dummy && cy >= 0 && cy < Hc && cx >= 0 && cx < Wc;
      mbar_wait(&ctl->acc_full[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (uint32_t)acc * 128u + (uint32_t)mtile * 64u + ((uint32_t)(quarter * 32) << 16);
      uint32_t v[2][32];
      tmem_ld_32x32b_x32(taddr, v[0]);
      tmem_ld_32x32b_x32(taddr + 32, v[1]);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->acc_empty[acc]);
      mbar_wait(&ctl->out_empty[ob], oph ^ 1);  // TMA has read the previous tile out of this buffer
      const uint32_t stage = stage0 + (uint32_t)ob * ST_OUT_BUF;
      if (!dummy) {
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            uint4 pk = make_uint4(0u, 0u, 0u, 0u);
            if (valid) {
              const float4 b0 = *reinterpret_cast<const float4*>(bias_s + h2 * 32 + qq * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(bias_s + h2 * 32 + qq * 8 + 4);
              pk.x = pack2<true>(__uint_as_float(v[h2][qq * 8 + 0]) + b0.x, __uint_as_float(v[h2][qq * 8 + 1]) + b0.y);
              pk.y = pack2<true>(__uint_as_float(v[h2][qq * 8 + 2]) + b0.z, __uint_as_float(v[h2][qq * 8 + 3]) + b0.w);
              pk.z = pack2<true>(__uint_as_float(v[h2][qq * 8 + 4]) + b1.x, __uint_as_float(v[h2][qq * 8 + 5]) + b1.y);
              pk.w = pack2<true>(__uint_as_float(v[h2][qq * 8 + 6]) + b1.z, __uint_as_float(v[h2][qq * 8 + 7]) + b1.w);
            }
            st_sts128(stage + st_chunk(my_off, h2 * 4 + qq), pk);
          }
      }
      st_bar_sync(1, 32 * ST_EPI_WARPS);  // the whole 17 x 15 region is staged
      // pooled block: 56 pooled pixels x 8 channel chunks
      for (int i = et; i < ST_PH * ST_PW * 8; i += 32 * ST_EPI_WARPS) {
        const int ch = i & 7, pp = i >> 3;
        const int pyl = pp / ST_PW, pxl = pp - pyl * ST_PW;
        uint4 m = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const uint4 val = st_lds128(stage + st_chunk(st_px_off(2 * pyl + dy, 2 * pxl + dx), ch));
            m.x = st_max2(m.x, val.x);
            m.y = st_max2(m.y, val.y);
            m.z = st_max2(m.z, val.z);
            m.w = st_max2(m.w, val.w);
          }
        st_sts128(stage + (uint32_t)(ST_OWNED_BYTES + ST_HALO_BYTES) + st_chunk((uint32_t)pp * 128u, ch), m);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->out_full[ob]);  // the buffer is restaged only after TMA has read it (out_empty)
      if (++acc == 4) {
        acc = 0;
        acc_phase ^= 1;
      }
      if (++ob == 2) {
        ob = 0;
        oph ^= 1;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == ST_LOAD_WARPS) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// =====================================================================================
// Stem v3: the same fused stem + pool WITHOUT an im2col pass.
//
// The A operand of tcgen05.mma may be any "K-major, no swizzle" arrangement of 8-row x 16-byte
// core matrices: rows of a core matrix 16 bytes apart, core matrices LBO apart along K and SBO
// apart along M.  For a stride-2, single-channel convolution the K chunk (filter row ky) of
// output pixel cx is the 16 bytes  in[2*cy + ky - 3][2*cx - 4 .. 2*cx + 3]  -- so the chunks of the
// output pixels cx, cx + 4, cx + 8, ... lie 16 bytes apart IN THE RAW IMAGE ROW.  Eight of them
// are one core matrix = 128 contiguous bytes of one input row; the next filter row is the next
// input row (LBO = row pitch) and the next output row two input rows further (SBO = 2 pitches).
// One MMA (M = 128) therefore covers 16 output rows x 8 output pixels of one "phase"
// r = cx mod 4, reading the raw window in place; the four phases need the window shifted by
// 0 / 2 / 4 / 6 pixels so that every core matrix starts on a 16-byte boundary -- four copies of
// the same 38 x 64-pixel window at shifted x positions.  TMA cannot make them: a tensor-map
// coordinate whose byte offset in dimension 0 is not a multiple of 16 raises an illegal-instruction
// fault (tests/micro/tma_plain_test.cu, measured on B200), so four loader warps issue cp.async
// copies of 16 / 8 / 4 bytes -- whatever the phase's alignment allows -- with zero fill outside
// the image (= the convolution's zero padding).  104 warp-level copy instructions per tile and no
// shared-memory reads replace v2's im2col (700 LSU wavefronts per 224 pixels; ncu: l1tex 75 %).
//
// A CTA tile is a 16 x 32 block of conv pixels (four phases x 128 rows, 256 TMEM columns):
// conv rows [2*py0 - 1, 2*py0 + 15), columns [2*px0 - 1, 2*px0 + 31).  It yields the 7 x 15 pooled
// pixels whose windows it contains and stores conv rows 1..14; columns are stored per phase (the
// staging order [row][i] of a phase makes the epilogue's shared-memory stores conflict-free), all
// 32 of them -- columns 0 and 31 duplicate the neighbour tiles' identical values.
//   warps 0..1  loaders (cp.async, two tiles ahead)             warp 2   MMA issuer (16 MMAs per tile)
//   warp 3      weights, then TMA-store issuer (4 + 1 stores per tile; TMA stores must not start at a
//               negative coordinate either, see the shifted staging of the leftmost tiles)
//   warps 4..11 staging: thread = (row, i) of all four phases x half of the channels; bias, ReLU, pack, stage
//   warps 12..15 pooling: 3x3 max over the staged tile, one pipeline step behind the staging warps
// =====================================================================================
namespace {
constexpr int S3_PH = 7, S3_PW = 15;                 // pooled block
constexpr int S3_ROWS = 16;                          // conv region: 16 rows x 32 columns
constexpr int S3_RAW_ROWS = 2 * (S3_ROWS - 1) + 8;   // 38 input rows (7 filter rows + the zero-weight 8th)
constexpr int S3_COPY_BYTES = S3_RAW_ROWS * 128;     // 4864: one phase-shifted window, 64 px x 2 B per row
constexpr int S3_A_STAGE = 20 * 1024;                // 4 copies (19456 B), 1024-aligned
constexpr int S3_PHASE_STAGE = S3_ROWS * 8 * 128;    // 16 KB: staged conv pixels of one phase
constexpr int S3_POOL_BYTES = 14 * 1024;             // 105 pooled pixels x 128 B (13440), 1024-aligned
constexpr int S3_OUT_BUF = 4 * S3_PHASE_STAGE + S3_POOL_BYTES;  // 78 KB
constexpr int S3_LOAD_WARPS = 2;
constexpr int S3_EPI_WARPS = 8;    // TMEM -> bias / ReLU / pack -> staging
constexpr int S3_POOL_WARPS = 4;   // 3x3 max over the staged tile of the PREVIOUS step of the pipeline
constexpr int S3_THREADS = 32 * (S3_LOAD_WARPS + 2 + S3_EPI_WARPS + S3_POOL_WARPS);

__device__ __forceinline__ uint64_t s3_desc_a(uint32_t addr, bool swap) {  // K-major, no swizzle: LBO = 128 B, SBO = 256 B
  const uint64_t lbo = swap ? 256u : 128u, sbo = swap ? 128u : 256u;
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1ull << 46);
}
// Phase 0 of the leftmost tiles (conv column -1 does not exist): pixels i = 1..7 of rows 1..14 are staged as a
// dense 14 x 7 box at the start of the phase buffer (what its TMA store reads), rows 0 and 15 behind it.
__device__ __forceinline__ int s3_row_shifted(int ry, int i) {
  return (ry >= 1 && ry <= 14) ? (ry - 1) * 7 + (i - 1) : 98 + (ry == 0 ? 0 : 7) + (i - 1);
}
}  // namespace

__global__ void __launch_bounds__(S3_THREADS, 1) stem_pool_v3_kernel(const __grid_constant__ ConvStemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_ring = smem;                                          // a_stages x 20 KB
  uint8_t* out_stage = a_ring + (size_t)p.a_stages * S3_A_STAGE;   // 2 x 78 KB
  uint8_t* b_area = out_stage + 2 * S3_OUT_BUF;                    // 8 KB weights
  StemCtl* ctl = reinterpret_cast<StemCtl*>(b_area + 64 * 128);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctl) + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_x * p.tiles_y * p.NB;
  const int Hc = p.H, Wc = p.W;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(&ctl->a_full[i], 32 * S3_LOAD_WARPS);
      mbar_init(&ctl->a_empty[i], 1);
    }
    mbar_init(&ctl->w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->acc_full[i], 1);
      mbar_init(&ctl->acc_empty[i], S3_EPI_WARPS);
      mbar_init(&ctl->staged_full[i], S3_EPI_WARPS);
      mbar_init(&ctl->out_full[i], S3_POOL_WARPS);
      mbar_init(&ctl->out_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (threadIdx.x < 64) bias_s[threadIdx.x] = p.bias[threadIdx.x];
  if (warp == S3_LOAD_WARPS) tmem_alloc<512>(&ctl->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  auto decode = [&](int t, int& px0, int& py0, int& n) {
    auto fdiv = [](uint32_t v, const FastDiv& f) { return f.m ? __umulhi(v, f.m) : v; };
    const uint32_t q = fdiv((uint32_t)t, p.div_tx);
    const int tx = (int)((uint32_t)t - q * p.div_tx.d);
    const uint32_t q2 = fdiv(q, p.div_ty);
    const int ty = (int)(q - q2 * p.div_ty.d);
    n = (int)q2 + p.n_base;
    px0 = tx * S3_PW;
    py0 = ty * S3_PH;
  };

  if (warp < S3_LOAD_WARPS) {
    // ===================== loaders: four phase-shifted copies of the raw window =====================
    const int ptid = threadIdx.x;  // 0..127
    const int Hin = 2 * Hc, Win = 2 * Wc;
    const uint32_t ring = smem_u32(a_ring);
    auto fetch = [&](int tt, int stage) {
      if (tt < total_tiles && !(p.dbg & 2)) {
        int px2, py2, n2;
        decode(tt, px2, py2, n2);
        const uint16_t* img = p.in + (int64_t)n2 * Hin * Win;
        const int iy0 = 4 * py2 - 5;
        const uint32_t stage_base = ring + (uint32_t)stage * S3_A_STAGE;
        if (iy0 >= 0 && iy0 + S3_RAW_ROWS <= Hin && 4 * px2 - 6 >= 0 && 4 * px2 + 64 <= Win) {
          // Interior tile (all but the image border): no bounds checks, addresses advance by constants.  A loader
          // warp is a single instruction stream, so the ~20 instructions per copy of the checked path below, not
          // bandwidth, bounded the whole kernel (knock-out experiment: 0.94 -> 0.64 ms per 64 slices without loads).
          // Phases 0 and 2 start on 4-byte boundaries only (4-byte copies, a warp covers one 128-byte row);
          // phases 1 and 3 on 8-byte boundaries (8-byte copies, a warp covers two rows).
          const uint8_t* win = reinterpret_cast<const uint8_t*>(img + (int64_t)iy0 * Win + (4 * px2 - 6));
          const uint32_t pitch = (uint32_t)Win * 2u;
          constexpr int NT = 32 * S3_LOAD_WARPS;
          {
            const int rr0 = ptid >> 5, cc = ptid & 31;
            const uint8_t* s0 = win + (size_t)rr0 * pitch + cc * 4;
            const uint32_t d0 = stage_base + (uint32_t)rr0 * 128u + (uint32_t)cc * 4u;
#pragma unroll
            for (int q = 0; q < (S3_RAW_ROWS * 32 + NT - 1) / NT; ++q) {
              if (rr0 + q * (NT / 32) < S3_RAW_ROWS) {
                const uint8_t* sp = s0 + (size_t)q * (NT / 32) * pitch;
                const uint32_t dp = d0 + (uint32_t)q * (NT / 32) * 128u;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dp), "l"(sp) : "memory");                                // phase 0
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dp + 2 * S3_COPY_BYTES), "l"(sp + 8) : "memory");      // phase 2: 4 pixels on
              }
            }
          }
          {
            const int rr0 = ptid >> 4, cc = ptid & 15;
            const uint8_t* s0 = win + (size_t)rr0 * pitch + cc * 8 + 4;  // phase 1: 2 pixels on
            const uint32_t d0 = stage_base + S3_COPY_BYTES + (uint32_t)rr0 * 128u + (uint32_t)cc * 8u;
#pragma unroll
            for (int q = 0; q < (S3_RAW_ROWS * 16 + NT - 1) / NT; ++q) {
              if (rr0 + q * (NT / 16) < S3_RAW_ROWS) {
                const uint8_t* sp = s0 + (size_t)q * (NT / 16) * pitch;
                const uint32_t dp = d0 + (uint32_t)q * (NT / 16) * 128u;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dp), "l"(sp) : "memory");                                // phase 1
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dp + 2 * S3_COPY_BYTES), "l"(sp + 8) : "memory");      // phase 3
              }
            }
          }
        } else
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          // copy r: window of output pixel rx = r starts at input column 2 * (2 * px2 - 1 + r) - 4
          const int ix0 = 4 * px2 - 6 + 2 * r;
          const uint32_t dst0 = ring + (uint32_t)stage * S3_A_STAGE + (uint32_t)r * S3_COPY_BYTES;
          const int al = (2 * ix0) & 15;  // byte alignment of the window start (rows are 16-byte aligned: Win % 8 == 0)
          if (al == 0) {
            for (int i = ptid; i < S3_RAW_ROWS * 8; i += 32 * S3_LOAD_WARPS) {
              const int rr = i >> 3, cc = i & 7;
              const int iy = iy0 + rr, ix = ix0 + 8 * cc;
              const bool ok = iy >= 0 && iy < Hin && ix >= 0 && ix < Win;
              const uint16_t* g = ok ? img + (int64_t)iy * Win + ix : p.in;
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + rr * 128 + cc * 16), "l"(g), "r"(ok ? 16u : 0u) : "memory");
            }
          } else if (al == 8) {
            for (int i = ptid; i < S3_RAW_ROWS * 16; i += 32 * S3_LOAD_WARPS) {
              const int rr = i >> 4, cc = i & 15;
              const int iy = iy0 + rr, ix = ix0 + 4 * cc;
              const bool ok = iy >= 0 && iy < Hin && ix >= 0 && ix < Win;
              const uint16_t* g = ok ? img + (int64_t)iy * Win + ix : p.in;
              asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst0 + rr * 128 + cc * 8), "l"(g), "r"(ok ? 8u : 0u) : "memory");
            }
          } else {
            for (int i = ptid; i < S3_RAW_ROWS * 32; i += 32 * S3_LOAD_WARPS) {
              const int rr = i >> 5, cc = i & 31;
              const int iy = iy0 + rr, ix = ix0 + 2 * cc;
              const bool ok = iy >= 0 && iy < Hin && ix >= 0 && ix < Win;
              const uint16_t* g = ok ? img + (int64_t)iy * Win + ix : p.in;
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst0 + rr * 128 + cc * 4), "l"(g), "r"(ok ? 4u : 0u) : "memory");
            }
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // stage of the k-th tile of this CTA = k % a_stages; copies run two tiles ahead (a_stages >= 3)
    int k = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++k) {
      if (k == 0) {
        fetch(t, 0);
        fetch(t + (int)gridDim.x, 1 % p.a_stages);
      }
      {
        const int k2 = k + 2, st2 = k2 % p.a_stages;
        // the MMAs of the tile that used this stage last (k2 - a_stages) have completed
        mbar_wait(&ctl->a_empty[st2], ((uint32_t)(k2 / p.a_stages) & 1u) ^ 1u);
        fetch(t + 2 * (int)gridDim.x, st2);
      }
      asm volatile("cp.async.wait_group 2;" ::: "memory");  // this thread's copies of tile k have landed
      fence_proxy_async_smem();
      mbar_arrive(&ctl->a_full[k % p.a_stages]);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp == S3_LOAD_WARPS) {
    // ===================== MMA issuer =====================
    int as = 0, acc = 0;
    uint32_t aph = 0, acc_phase = 0;
    const uint32_t idesc = umma_idesc_act(128, 64);
    const uint64_t a_desc0 = s3_desc_a(smem_u32(a_ring), (p.dbg & 8) != 0);
    const uint64_t b_desc0 = st_desc(smem_u32(b_area));
    mbar_wait(&ctl->w_full, 0);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      mbar_wait(&ctl->acc_empty[acc], acc_phase ^ 1);
      mbar_wait(&ctl->a_full[as], aph);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t d0 = tmem_base + (uint32_t)acc * 256u;
        const uint32_t a_units = (uint32_t)(as * S3_A_STAGE) >> 4;
#pragma unroll
        for (int r = 0; r < ((p.dbg & 1) ? 0 : 4); ++r) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            // filter rows 2ks, 2ks + 1 = input rows +2ks, +2ks+1 of the window: 256 bytes further per K-step
            const uint32_t a_off = a_units + (uint32_t)((r * S3_COPY_BYTES + ks * 256) >> 4);
            umma_bf16_ss(d0 + (uint32_t)r * 64u, (a_desc0 & 0xffffffff00000000ull) | (uint32_t)((uint32_t)a_desc0 + a_off),
                         (b_desc0 & 0xffffffff00000000ull) | (uint32_t)((uint32_t)b_desc0 + 2 * ks), idesc, ks != 0 ? 1u : 0u);
          }
        }
        umma_commit(&ctl->a_empty[as]);
        umma_commit(&ctl->acc_full[acc]);
      }
      __syncwarp();
      if (++as == p.a_stages) {
        as = 0;
        aph ^= 1;
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp == S3_LOAD_WARPS + 1) {
    // ===================== weights, then TMA-store issuer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(&ctl->w_full, 64 * 128);
      bulk_load_1d(b_area, p.wpacked, 64 * 128, &ctl->w_full);
      int ob = 0;
      uint32_t oph = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int px0, py0, n;
        decode(t, px0, py0, n);
        // the conv pixels leave as soon as they are staged, while the pooling warps still work on the same tile
        mbar_wait(&ctl->staged_full[ob], oph);
        uint8_t* buf = out_stage + (size_t)ob * S3_OUT_BUF;
        const int cx0 = 2 * px0 - 1, cy0 = 2 * py0 - 1;
        if (!(p.dbg & 4)) {
#pragma unroll
        for (int r = 0; r < ((p.dbg & 16) ? 0 : 4); ++r) {
          const int x = cx0 + r;  // conv column of this phase's first pixel; -1 for phase 0 of the leftmost tiles
          // rows 1..14 of the phase: 14 x 8 pixels, starting one staged row (8 x 128 B) into the phase.  TMA stores
          // must not start at a negative coordinate (measured: the store never completes): the leftmost tiles stage
          // phase 0 shifted by one pixel (column -1 does not exist) and store 7 pixels per row through a second map.
          if (x < 0) tma_store_5d(p.out_map7, buf + r * S3_PHASE_STAGE, 0, 3, 0, cy0 + 1, n);
          else tma_store_5d(p.out_map, buf + r * S3_PHASE_STAGE + 8 * 128, 0, x & 3, x >> 2, cy0 + 1, n);  // dims (c, x % 4, x / 4, y, n)
        }
        tma_store_commit();
        }
        mbar_wait(&ctl->out_full[ob], oph);  // pooled block staged
        if (!(p.dbg & 4)) {
        if (!(p.dbg & 32)) tma_store_5d(p.pool_map, buf + 4 * S3_PHASE_STAGE, 0, px0, 0, py0, n);
        }
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(&ctl->out_empty[ob]);
        if (++ob == 2) {
          ob = 0;
          oph ^= 1;
        }
      }
      tma_store_wait_all<0>();
    }
  } else if (warp < S3_LOAD_WARPS + 2 + S3_EPI_WARPS) {
    // ===================== staging warps: TMEM -> bias, ReLU, 16-bit -> swizzled tile =====================
    // A warp owns one TMEM lane quarter (32 conv pixels per phase) and ONE HALF of the channels for all four phases,
    // so its 32 bias values live in registers for the whole kernel (a float4 broadcast from shared memory costs four
    // LSU wavefronts; ncu showed the bias reads as a third of the first version's shared-memory traffic).
    const int ew = warp - (S3_LOAD_WARPS + 2);  // 0..7
    const int quarter = warp & 3;     // TMEM lane quarter
    const int chalf = ew >> 2;        // channels 32 * chalf .. 32 * chalf + 31
    const int m = quarter * 32 + lane;
    const int ry = m >> 3, i = m & 7;
    const uint32_t stage0 = smem_u32(out_stage);
    float bias_r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias_r[j] = bias_s[chalf * 32 + j];
    int acc = 0, ob = 0;
    uint32_t acc_phase = 0, oph = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int px0, py0, n;
      decode(t, px0, py0, n);
      const int cy = 2 * py0 - 1 + ry;
      const bool row_ok = cy >= 0 && cy < Hc;
      const bool shift0 = px0 == 0;
      mbar_wait(&ctl->acc_full[acc], acc_phase);
      tc_fence_after_sync();
      mbar_wait(&ctl->out_empty[ob], oph ^ 1);  // TMA has read the previous tile out of this buffer
      const uint32_t stage = stage0 + (uint32_t)ob * S3_OUT_BUF;
      const uint32_t tbase = tmem_base + (uint32_t)acc * 256u + (uint32_t)chalf * 32u + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
      for (int rp = 0; rp < 2; ++rp) {  // two phases per round: both TMEM loads in flight before the first is consumed
        uint32_t v[2][32];
        tmem_ld_32x32b_x32(tbase + (uint32_t)(2 * rp) * 64u, v[0]);
        tmem_ld_32x32b_x32(tbase + (uint32_t)(2 * rp + 1) * 64u, v[1]);
        tmem_ld_wait();
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int r = 2 * rp + rr;
          const int cx = 2 * px0 - 1 + r + 4 * i;
          const bool valid = row_ok && cx >= 0 && cx < Wc;
          // staged row of this pixel inside its phase: ry * 8 + i; phase 0 of the leftmost tiles is staged as the
          // 7-pixel-wide box its store uses (s3_row_shifted)
          const bool sh = r == 0 && shift0;
          if (sh && i == 0) continue;  // conv column -1: outside the image
          const int srow = sh ? s3_row_shifted(ry, i) : ry * 8 + i;
          const uint32_t row_addr = stage + (uint32_t)r * S3_PHASE_STAGE + (uint32_t)srow * 128u;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            uint4 pk = make_uint4(0u, 0u, 0u, 0u);
            if (valid) {
              pk.x = pack2<true>(__uint_as_float(v[rr][qq * 8 + 0]) + bias_r[qq * 8 + 0], __uint_as_float(v[rr][qq * 8 + 1]) + bias_r[qq * 8 + 1]);
              pk.y = pack2<true>(__uint_as_float(v[rr][qq * 8 + 2]) + bias_r[qq * 8 + 2], __uint_as_float(v[rr][qq * 8 + 3]) + bias_r[qq * 8 + 3]);
              pk.z = pack2<true>(__uint_as_float(v[rr][qq * 8 + 4]) + bias_r[qq * 8 + 4], __uint_as_float(v[rr][qq * 8 + 5]) + bias_r[qq * 8 + 5]);
              pk.w = pack2<true>(__uint_as_float(v[rr][qq * 8 + 6]) + bias_r[qq * 8 + 6], __uint_as_float(v[rr][qq * 8 + 7]) + bias_r[qq * 8 + 7]);
            }
            st_sts128(row_addr + ((uint32_t)((chalf * 4 + qq) ^ (srow & 7)) << 4), pk);  // SW128: chunk ^ (staged row & 7)
          }
        }
      }
      tc_fence_before_sync();
      fence_proxy_async_smem();  // the staged pixels are read by TMA (store) as well as by the pooling warps
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&ctl->acc_empty[acc]);
        mbar_arrive(&ctl->staged_full[ob]);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
      if (++ob == 2) {
        ob = 0;
        oph ^= 1;
      }
    }
  } else {
    // ===================== pooling warps: 3x3 / stride 2 max over the staged tile =====================
    // One tile behind the staging warps (they already fill the other buffer).  Thread = (pooled column, 8-channel
    // chunk): the 3-wide row maxima of the 15 conv rows are reduced in registers, 45 loads for 7 pooled pixels.
    const int et = (warp - (S3_LOAD_WARPS + 2 + S3_EPI_WARPS)) * 32 + lane;  // 0..127
    const uint32_t stage0 = smem_u32(out_stage);
    const int pxl = et >> 3, pch = et & 7;  // et < 120 active
    uint32_t pcol[3];
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int wx = 2 * pxl + dx, slot = wx >> 2;
      pcol[dx] = (uint32_t)(wx & 3) * S3_PHASE_STAGE + (uint32_t)slot * 128u + ((uint32_t)(pch ^ slot) << 4);
    }
    int ob = 0;
    uint32_t sph = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int px0, py0, n;
      decode(t, px0, py0, n);
      const bool shift0 = px0 == 0;
      mbar_wait(&ctl->staged_full[ob], sph);
      const uint32_t stage = stage0 + (uint32_t)ob * S3_OUT_BUF;
      const uint32_t pool_stage = stage + 4u * S3_PHASE_STAGE;
      if (!shift0) {
        if (et < S3_PW * 8) {
          uint4 h0, h1, h2;  // row maxima of conv rows 2j, 2j + 1, 2j + 2
          auto rowmax = [&](int q) {
            const uint32_t ra = stage + (uint32_t)q * 1024u;
            const uint4 a0 = st_lds128(ra + pcol[0]), a1 = st_lds128(ra + pcol[1]), a2 = st_lds128(ra + pcol[2]);
            uint4 h;
            h.x = st_max2(st_max2(a0.x, a1.x), a2.x);
            h.y = st_max2(st_max2(a0.y, a1.y), a2.y);
            h.z = st_max2(st_max2(a0.z, a1.z), a2.z);
            h.w = st_max2(st_max2(a0.w, a1.w), a2.w);
            return h;
          };
          h0 = rowmax(0);
#pragma unroll
          for (int j = 0; j < S3_PH; ++j) {
            h1 = rowmax(2 * j + 1);
            h2 = rowmax(2 * j + 2);
            uint4 mx;
            mx.x = st_max2(st_max2(h0.x, h1.x), h2.x);
            mx.y = st_max2(st_max2(h0.y, h1.y), h2.y);
            mx.z = st_max2(st_max2(h0.z, h1.z), h2.z);
            mx.w = st_max2(st_max2(h0.w, h1.w), h2.w);
            const int pp = j * S3_PW + pxl;
            st_sts128(pool_stage + (uint32_t)pp * 128u + ((uint32_t)(pch ^ (pp & 7)) << 4), mx);
            h0 = h2;
          }
        }
      } else {
        // leftmost tiles: phase 0 is staged in the shifted 7-wide order (generic item loop)
        for (int it = et; it < S3_PH * S3_PW * 8; it += 32 * S3_POOL_WARPS) {
          const int ch = it & 7, pp = it >> 3;
          const int pyl = pp / S3_PW, px_ = pp - pyl * S3_PW;
          uint4 mx = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const int wx = 2 * px_ + dx;
              int srow = (2 * pyl + dy) * 8 + (wx >> 2);
              if ((wx & 3) == 0) {
                if (wx == 0) continue;  // conv column -1: outside the image
                srow = s3_row_shifted(2 * pyl + dy, wx >> 2);
              }
              const uint4 val = st_lds128(stage + (uint32_t)(wx & 3) * S3_PHASE_STAGE + (uint32_t)srow * 128u +
                                          ((uint32_t)(ch ^ (srow & 7)) << 4));
              mx.x = st_max2(mx.x, val.x);
              mx.y = st_max2(mx.y, val.y);
              mx.z = st_max2(mx.z, val.z);
              mx.w = st_max2(mx.w, val.w);
            }
          st_sts128(pool_stage + (uint32_t)pp * 128u + ((uint32_t)(ch ^ (pp & 7)) << 4), mx);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->out_full[ob]);
      if (++ob == 2) {
        ob = 0;
        sph ^= 1;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == S3_LOAD_WARPS) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

size_t conv_stem_smem_bytes(const ConvStemParams& p) {
  if (p.version == 3) return (size_t)p.a_stages * S3_A_STAGE + 2 * S3_OUT_BUF + 64 * 128 + 1024 + 1024;
  return (size_t)p.a_stages * ST_A_STAGE + 2 * ST_OUT_BUF + 64 * 128 + ST_RAW_RING * ST_RAW_BYTES + 1024 + 1024;
}

cudaError_t conv_stem_configure() {
  cudaError_t e = cudaFuncSetAttribute(stem_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_pool_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  return e;
}

void conv_stem_tiles(int version, int Hc, int Wc, int* tiles_x, int* tiles_y) {
  const int pw = version == 3 ? S3_PW : ST_PW, ph = version == 3 ? S3_PH : ST_PH;
  *tiles_x = (Wc / 2 + pw - 1) / pw;
  *tiles_y = (Hc / 2 + ph - 1) / ph;
}

cudaError_t launch_conv_stem(const ConvStemParams& p0, int num_sms, cudaStream_t st) {
  ConvStemParams p = p0;
  const int64_t total = (int64_t)p.tiles_x * p.tiles_y * p.NB;
  const int64_t dmax = p.tiles_x > p.tiles_y ? p.tiles_x : p.tiles_y;
  if (total * dmax >= (1ll << 32)) return cudaErrorInvalidValue;  // FastDiv exactness bound
  p.div_tx = make_fastdiv((uint32_t)p.tiles_x);
  p.div_ty = make_fastdiv((uint32_t)p.tiles_y);
  const int grid = total < num_sms ? (int)total : num_sms;
  if (p.version == 3) stem_pool_v3_kernel<<<grid, S3_THREADS, conv_stem_smem_bytes(p), st>>>(p);
  else stem_pool_kernel<<<grid, ST_THREADS, conv_stem_smem_bytes(p), st>>>(p);
  return cudaGetLastError();
}

}  // namespace vsb
