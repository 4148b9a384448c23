// tcgen05 / TMEM implicit-GEMM convolution, persistent, warp specialised.
// See conv_tc.cuh for the GEMM view and the tensor-map conventions.
//
//   warp 0      TMA producer: per K-slab, one 5-D box load per pixel class for
//               the A operand + one linear bulk copy of the pre-swizzled
//               weights; `full[stage]` mbarrier counts the bytes.
//   warp 1      TMEM allocator and MMA issuer: one elected lane issues
//               tcgen05.mma (M=128, N=BN, K=16) row_bytes/32 times per slab,
//               tcgen05.commit releases the smem stage (`empty[stage]`) and,
//               after the last slab, publishes the accumulator (`acc_full`).
//   warps 2..5  epilogue: tcgen05.ld the fp32 accumulator (thread = output
//               pixel), + folded-BN bias, + residual, ReLU, bf16 (or f32) NHWC
//               store; `acc_empty` hands the TMEM stage back.  Two accumulator
//               stages (2 x 256 columns) overlap the epilogue of tile i with
//               the main loop of tile i+1.
//
// Replaces every cuDNN conv2d + batch_norm + relu_ + add_ + nearest
// upsample + cat the reference reaches through `self.model(...)`
// (vol_seg_2d_predictor.py:44).
#include "conv_tc.cuh"

namespace vsb {

namespace {

struct SmemCtl {
  uint64_t full[TC_MAX_STAGES];
  uint64_t empty[TC_MAX_STAGES];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  uint32_t pad[3];
};

constexpr int kCtlBytes = 1024;      // >= sizeof(SmemCtl)
constexpr int kBiasBytes = 2048 * 4; // bias for up to 2048 output channels

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
  N0 = tn << p.nt_log2;
}

}  // namespace

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: required by the 128B swizzle atoms (TMA and UMMA).
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* stages = smem;
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem + (size_t)p.num_stages * p.stage_bytes);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctl) + kCtlBytes);

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
      mbar_init(&ctl->acc_empty[i], 128);
    }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < p.n_tiles * p.BN; i += TC_THREADS) bias_s[i] = p.bias[i];
  if (warp == 1) tmem_alloc<512>(&ctl->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int ncls = 1 << p.ncls_log2;
      const int rows_per_cls = 128 >> p.ncls_log2;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int n_tile, X0, Y0, N0;
        decode_tile(p, t, n_tile, X0, Y0, N0);
        for (int s = 0; s < p.num_slabs; ++s) {
          const TcSlab sl = p.slabs[s];
          mbar_wait(&ctl->empty[stage], phase ^ 1);
          uint8_t* a_dst = stages + (size_t)stage * p.stage_bytes;
          uint8_t* b_dst = a_dst + p.a_bytes;
          mbar_arrive_expect_tx(&ctl->full[stage], (uint32_t)((128 + p.BN) * sl.row_bytes));
          const void* map = &p.maps[sl.map];
          for (int q = 0; q < ncls; ++q) {
            const int ty = (q >> 1) + sl.dy, tx = (q & 1) + sl.dx;
            int cy, cx, pary = 0, parx = 0;
            if (sl.flags & TC_HALVE) {
              cy = Y0 + (ty >> 1);
              cx = X0 + (tx >> 1);
              pary = ty & 1;
              parx = tx & 1;
            } else {
              cy = Y0 + ty;
              cx = X0 + tx;
            }
            const bool folded = (sl.flags & TC_FOLDED) != 0;
            const int cc = sl.c0 + (folded ? parx * sl.cfold : 0);
            const int pp = folded ? pary : 0;
            tma_load_5d(map, &ctl->full[stage], a_dst + (size_t)q * rows_per_cls * sl.row_bytes,
                        cc, cx, pp, cy, N0);
          }
          bulk_load_1d(b_dst,
                       p.wpacked + (size_t)sl.w_off16 * 16 + (size_t)n_tile * p.BN * sl.row_bytes,
                       (uint32_t)(p.BN * sl.row_bytes), &ctl->full[stage]);
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = umma_idesc_act(128, p.BN);
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&ctl->acc_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        for (int s = 0; s < p.num_slabs; ++s) {
          const int row_bytes = p.slabs[s].row_bytes;
          mbar_wait(&ctl->full[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(stages + (size_t)stage * p.stage_bytes);
          const uint32_t b_addr = a_addr + p.a_bytes;
          const int ksteps = row_bytes >> 5;
          for (int k = 0; k < ksteps; ++k) {
            umma_bf16_ss(d_tmem, umma_smem_desc(a_addr + k * 32, row_bytes),
                         umma_smem_desc(b_addr + k * 32, row_bytes), idesc, (s | k) != 0);
          }
          umma_commit(&ctl->empty[stage]);
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&ctl->acc_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ============================ epilogue ================================
    int acc = 0;
    uint32_t acc_phase = 0;
    const int quarter = warp & 3;           // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;    // accumulator row == pixel within the tile
    const int rpc_log2 = 7 - p.ncls_log2;
    const int r = row & ((1 << rpc_log2) - 1);
    const int q = row >> rpc_log2;
    const int xi = r & ((1 << p.bw_log2) - 1);
    const int yi = (r >> p.bw_log2) & ((1 << p.bh_log2) - 1);
    const int ni = r >> (p.bw_log2 + p.bh_log2);
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
      const int64_t pix = ((int64_t)on * p.H + oy) * p.W + ox;
      const int ch0 = n_tile * p.BN;

      mbar_wait(&ctl->acc_full[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(quarter * 32) << 16);
      for (int c = 0; c < p.BN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const int ch = ch0 + c + g8 * 8;
            if (ch >= p.cout) break;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              f[j] = __uint_as_float(v[g8 * 8 + j]) + bias_s[ch + j];
            const bool full8 = ch + 8 <= p.cout;
            if (p.residual) {
              if (full8) {
                const uint4 rv =
                    __ldg(reinterpret_cast<const uint4*>(p.residual + pix * p.cout + ch));
                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 rf = unpack_act2(rw[j]);
                  f[2 * j] += rf.x;
                  f[2 * j + 1] += rf.y;
                }
              } else {
                for (int j = 0; j < 8 && ch + j < p.cout; ++j)
                  f[j] += act_to_float(p.residual[pix * p.cout + ch + j]);
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (p.out_f32) {
              float* o = reinterpret_cast<float*>(p.out) + pix * p.cout + ch;
              if (full8 && (p.cout & 3) == 0) {
                *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
              } else {
                for (int j = 0; j < 8 && ch + j < p.cout; ++j) o[j] = f[j];
              }
            } else {
              uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + pix * p.cout + ch;
              if (full8 && (p.cout & 7) == 0) {
                uint4 pk;
                pk.x = pack_act2(f[0], f[1]);
                pk.y = pack_act2(f[2], f[3]);
                pk.z = pack_act2(f[4], f[5]);
                pk.w = pack_act2(f[6], f[7]);
                *reinterpret_cast<uint4*>(o) = pk;
              } else {
                for (int j = 0; j < 8 && ch + j < p.cout; ++j) o[j] = float_to_act(f[j]);
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
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

size_t conv_tc_smem_bytes(const ConvTcParams& p) {
  return (size_t)p.num_stages * p.stage_bytes + kCtlBytes + kBiasBytes + 1024;
}

cudaError_t conv_tc_configure() {
  return cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              227 * 1024);
}

cudaError_t launch_conv_tc(const ConvTcParams& p, int num_sms, cudaStream_t st) {
  const int total_tiles = p.n_tiles * p.tiles_x * p.tiles_y * p.tiles_n;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  conv_tc_kernel<<<grid, TC_THREADS, conv_tc_smem_bytes(p), st>>>(p);
  return cudaGetLastError();
}

}  // namespace vsb
