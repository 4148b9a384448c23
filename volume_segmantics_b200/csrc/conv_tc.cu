// tcgen05 / TMEM implicit-GEMM convolution, persistent, warp specialised.
// See conv_tc.cuh for the GEMM view and the tensor-map conventions.
//
//   warp 0      TMA producer: per K-slab, one 5-D box load per pixel class for
//               the A operand + one linear bulk copy of the pre-swizzled
//               weights; `full[stage]` mbarrier counts the bytes.
//   warp 1      TMEM allocator and MMA issuer: an elected lane issues
//               tcgen05.mma (M=128, N=BN, K=16) RB/32 times per slab,
//               tcgen05.commit releases the smem stage (`empty[stage]`) and,
//               after the last slab, publishes the accumulator (`acc_full`).
//   warps 2..9  epilogue: tcgen05.ld the fp32 accumulator (thread = output
//               pixel; two warps share a TMEM lane quarter and split the
//               columns), + folded-BN bias, + residual, ReLU, 16-bit (or f32)
//               NHWC store; `acc_empty` hands the TMEM stage back.  Two
//               accumulator stages (2 x 256 columns) overlap the epilogue of
//               tile i with the main loop of tile i+1.
// The producer and MMA loops run warp-convergent (every lane executes the
// waits, one elected lane issues), so addresses and descriptors stay in the
// uniform datapath; the K loop needs no table look-up on the MMA side because
// the slab width RB is a template parameter.
//
// Replaces every cuDNN conv2d + batch_norm + relu_ + add_ + nearest
// upsample + cat the reference reaches through `self.model(...)`
// (vol_seg_2d_predictor.py:44).
#include "conv_tc.cuh"
#include "conv_epilogue.cuh"

namespace vsb {

namespace {

struct SmemCtl {
  uint64_t full[TC_MAX_STAGES];
  uint64_t empty[TC_MAX_STAGES];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t res_full[2];
  uint32_t tmem_base;
  uint32_t pad[3];
};

constexpr int kCtlBytes = 1024;       // >= sizeof(SmemCtl)
constexpr int kBiasBytes = 2048 * 4;  // bias for up to 2048 output channels
constexpr int kRunBytes = TC_MAX_RUNS * (int)sizeof(TcRun);
constexpr int kEpiBuf = 16384;  // one staged 64-channel group: 128 rows x 128 B (SW128)

__device__ __forceinline__ void tc_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint4 tc_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tc_sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void decode_tile(const ConvTcParams& p, int t, int& n_tile, int& X0,
                                            int& Y0, int& N0) {
  n_tile = t % p.n_tiles;
  int sp = t / p.n_tiles;
  const int tx = sp % p.tiles_x;
  sp /= p.tiles_x;
  const int ty = sp % p.tiles_y;
  const int tn = sp / p.tiles_y;
  X0 = tx << p.bw_log2;
  Y0 = ty << p.bh_log2;
  N0 = (tn << p.nt_log2);
}

}  // namespace

template <int RB>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: required by the 128B swizzle atoms (TMA and UMMA).
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* stages = smem;
  const bool smem_epi = p.out_map != nullptr;
  // staging of the shared-memory epilogue: 2 output + 2 residual buffers, 1024-aligned behind the ring
  uint8_t* epi_stage = smem + (((size_t)p.num_stages * p.stage_bytes + 1023) & ~(size_t)1023);
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem_epi ? epi_stage + (p.res_map ? 4 : 2) * kEpiBuf
                                                     : smem + (size_t)p.num_stages * p.stage_bytes);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctl) + kCtlBytes);
  TcRun* runs_s = reinterpret_cast<TcRun*>(reinterpret_cast<uint8_t*>(bias_s) + kBiasBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.n_tiles * p.tiles_x * p.tiles_y * p.tiles_n;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&ctl->full[i], 1);
      mbar_init(&ctl->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->acc_full[i], 1);
      mbar_init(&ctl->acc_empty[i], 32 * TC_EPI_WARPS);
      mbar_init(&ctl->res_full[i], 1);
    }
    fence_mbar_init();
  }
  bias_s -= p.cout_off;  // indexed by absolute output channel
  for (int i = threadIdx.x; i < p.n_tiles * p.BN; i += TC_THREADS) bias_s[p.cout_off + i] = p.bias[p.cout_off + i];
  {
    const int4* src = reinterpret_cast<const int4*>(p.runs);
    int4* dst = reinterpret_cast<int4*>(runs_s);
    for (int i = threadIdx.x; i < p.num_runs * (int)(sizeof(TcRun) / 16); i += TC_THREADS)
      dst[i] = src[i];
  }
  if (warp == 1) tmem_alloc<512>(&ctl->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ============================ TMA producer ============================
    int stage = 0;
    uint32_t phase = 0;
    const int ncls = 1 << p.ncls_log2;
    const uint32_t cls_bytes = (uint32_t)((128 >> p.ncls_log2) * RB);
    const uint32_t tx_bytes = (uint32_t)((128 + p.BN) * RB);
    const uint32_t b_bytes = (uint32_t)(p.BN * RB);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n_tile, X0, Y0, N0;
      decode_tile(p, t, n_tile, X0, Y0, N0);
      const uint8_t* wtile = p.wpacked + (size_t)n_tile * b_bytes;
      for (int r = 0; r < p.num_runs; ++r) {
        const TcRun& run = runs_s[r];
        const void* map = &p.maps[run.map];
        const int nblk = run.nblk;
        const uint8_t* wsrc = wtile + (size_t)run.w_off16 * 16;
        const size_t wstep = (size_t)run.w_step16 * 16;
        for (int b = 0; b < nblk; ++b) {
          mbar_wait(&ctl->empty[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* a_dst = stages + (size_t)stage * p.stage_bytes;
            mbar_arrive_expect_tx(&ctl->full[stage], tx_bytes);
            for (int q = 0; q < ncls; ++q)
              tma_load_5d(map, &ctl->full[stage], a_dst + q * cls_bytes,
                          p.cin_off + run.cls[q][0] + b * (RB / 2), X0 + run.cls[q][1], run.cls[q][2],
                          Y0 + run.cls[q][3], N0 + p.n_base);
            bulk_load_1d(a_dst + p.a_bytes, wsrc + b * wstep, b_bytes, &ctl->full[stage]);
          }
          __syncwarp();
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ==============================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t idesc = umma_idesc_act(128, p.BN);
    const uint64_t a_desc0 = umma_smem_desc(smem_u32(stages), RB);
    const uint64_t b_desc0 = umma_smem_desc(smem_u32(stages) + p.a_bytes, RB);
    const uint32_t stage_step = (uint32_t)p.stage_bytes >> 4;  // descriptor address units (16 B)
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      mbar_wait(&ctl->acc_empty[acc], acc_phase ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
      for (int s = 0; s < p.num_slabs; ++s) {
        mbar_wait(&ctl->full[stage], phase);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint64_t a_desc = a_desc0 + (uint64_t)(stage * stage_step);
          const uint64_t b_desc = b_desc0 + (uint64_t)(stage * stage_step);
#pragma unroll
          for (int k = 0; k < RB / 32; ++k)
            umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (s | k) != 0);
          umma_commit(&ctl->empty[stage]);
          if (s == p.num_slabs - 1) umma_commit(&ctl->acc_full[acc]);
        }
        __syncwarp();
        if (++stage == p.num_stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (smem_epi) {
    // ================= epilogue through shared memory + TMA store =================
    // Work unit j = (tile, 64-channel group g); buffers alternate with j.  One barrier per unit:
    //   TMEM load -> (residual group landed) -> bias / residual / ReLU / pack -> swizzled staging
    //   -> [thread 0: the store that last used the OTHER buffer has been read] -> barrier
    //   -> [thread 0: TMA store of unit j; residual load of unit j + 2 into the buffer just consumed].
    int acc = 0;
    uint32_t acc_phase = 0;
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const int et = (warp - 2) * 32 + lane;
    const bool has_res = p.res_map != nullptr;
    const int groups = p.BN >> 6;
    const uint32_t out0 = smem_u32(epi_stage), res0 = out0 + 2 * kEpiBuf;
    // residual prefetch cursor (thread 0 only)
    int pf_t = blockIdx.x, pf_g = 0, pf_j = 0;
    auto prefetch_res = [&]() {
      if (pf_t >= total_tiles) return;
      int n_tile, X0, Y0, N0;
      decode_tile(p, pf_t, n_tile, X0, Y0, N0);
      uint64_t* bar = &ctl->res_full[pf_j & 1];
      mbar_arrive_expect_tx(bar, kEpiBuf);
      tma_load_5d(p.res_map, bar, epi_stage + (2 + (pf_j & 1)) * kEpiBuf, p.cout_off + n_tile * p.BN + pf_g * 64, X0, 0, Y0,
                  N0 + p.n_base);
      ++pf_j;
      if (++pf_g == groups) {
        pf_g = 0;
        pf_t += gridDim.x;
      }
    };
    if (has_res && et == 0) {
      prefetch_res();
      prefetch_res();
    }
    int j = 0;
    uint32_t rph[2] = {0, 0};
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n_tile, X0, Y0, N0;
      decode_tile(p, t, n_tile, X0, Y0, N0);
      const int ch0 = p.cout_off + n_tile * p.BN;
      const float* bt = bias_s + ch0;
      mbar_wait(&ctl->acc_full[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(quarter * 32) << 16);
      for (int g = 0; g < groups; ++g, ++j) {
        const int ob = j & 1;
        const int c = g * 64 + half * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c, v);
        tmem_ld_wait();
        if (g == groups - 1) {  // the accumulator stage is free as soon as every thread has its last columns
          tc_fence_before_sync();
          mbar_arrive(&ctl->acc_empty[acc]);
        }
        if (has_res) {
          mbar_wait(&ctl->res_full[ob], rph[ob]);
          rph[ob] ^= 1;
        }
        const uint32_t ost = out0 + (uint32_t)ob * kEpiBuf + (uint32_t)row * 128u;
        const uint32_t rst = res0 + (uint32_t)ob * kEpiBuf + (uint32_t)row * 128u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t off = (uint32_t)((half * 4 + q) ^ (row & 7)) << 4;
          const float4 b0 = *reinterpret_cast<const float4*>(bt + c + q * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(bt + c + q * 8 + 4);
          float f0 = __uint_as_float(v[q * 8 + 0]) + b0.x, f1 = __uint_as_float(v[q * 8 + 1]) + b0.y;
          float f2 = __uint_as_float(v[q * 8 + 2]) + b0.z, f3 = __uint_as_float(v[q * 8 + 3]) + b0.w;
          float f4 = __uint_as_float(v[q * 8 + 4]) + b1.x, f5 = __uint_as_float(v[q * 8 + 5]) + b1.y;
          float f6 = __uint_as_float(v[q * 8 + 6]) + b1.z, f7 = __uint_as_float(v[q * 8 + 7]) + b1.w;
          if (has_res) {
            const uint4 rv = tc_lds128(rst + off);
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
          tc_sts128(ost + off, pk);
        }
        fence_proxy_async_smem();
        if (et == 0) tma_store_wait_read<0>();  // the store of unit j - 1 (other buffer) has been read out
        tc_bar_sync(1, 32 * TC_EPI_WARPS);
        if (et == 0) {
          tma_store_5d(p.out_map, epi_stage + ob * kEpiBuf, ch0 + g * 64, X0, 0, Y0, N0 + p.n_base);
          tma_store_commit();
          if (has_res) prefetch_res();  // unit j + 2 reuses the residual buffer every thread has just finished with
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (et == 0) tma_store_wait_all<0>();
  } else {
    // ============================ epilogue ================================
    int acc = 0;
    uint32_t acc_phase = 0;
    const int quarter = warp & 3;         // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;     // which half of the column chunks
    const int row = quarter * 32 + lane;  // accumulator row == pixel within the tile
    const int rpc_log2 = 7 - p.ncls_log2;
    const int r = row & ((1 << rpc_log2) - 1);
    const int q = row >> rpc_log2;
    const int xi = r & ((1 << p.bw_log2) - 1);
    const int yi = (r >> p.bw_log2) & ((1 << p.bh_log2) - 1);
    const int ni = r >> (p.bw_log2 + p.bh_log2);
    const EpiOut eo{p.out, p.residual, p.out_f32, p.relu, p.cout};
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n_tile, X0, Y0, N0;
      decode_tile(p, t, n_tile, X0, Y0, N0);
      int oy, ox;
      if (p.ncls_log2) {
        oy = 2 * (Y0 + yi) + (q >> 1);
        ox = 2 * (X0 + xi) + (q & 1);
      } else {
        oy = Y0 + yi;
        ox = X0 + xi;
      }
      const int on = N0 + ni;
      const bool valid = on < p.NB && oy < p.H && ox < p.W;
      const int64_t pix = ((int64_t)(on + p.n_base) * p.H + oy) * p.W + ox;
      const int ch0 = p.cout_off + n_tile * p.BN;

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

size_t conv_tc_epilogue_bytes(bool with_residual) { return (size_t)(with_residual ? 4 : 2) * kEpiBuf + 1024; }

size_t conv_tc_smem_bytes(const ConvTcParams& p) {
  return (size_t)p.num_stages * p.stage_bytes + (p.out_map ? conv_tc_epilogue_bytes(p.res_map != nullptr) : 0) + kCtlBytes +
         kBiasBytes + kRunBytes + 1024;
}

cudaError_t conv_tc_configure() {
  cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<32>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024);
  return e;
}

cudaError_t launch_conv_tc(const ConvTcParams& p, int num_sms, cudaStream_t st) {
  const int total_tiles = p.n_tiles * p.tiles_x * p.tiles_y * p.tiles_n;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  const size_t smem = conv_tc_smem_bytes(p);
  if (p.row_bytes == 128) conv_tc_kernel<128><<<grid, TC_THREADS, smem, st>>>(p);
  else if (p.row_bytes == 64) conv_tc_kernel<64><<<grid, TC_THREADS, smem, st>>>(p);
  else conv_tc_kernel<32><<<grid, TC_THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace vsb
