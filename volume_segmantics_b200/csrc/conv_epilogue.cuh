// Epilogue helpers shared by the tcgen05 convolution kernels: fp32 accumulator row
// (+bias already added) -> optional residual / ReLU -> 16-bit or f32 NHWC store.
#pragma once
#include "common.cuh"

namespace vsb {

struct EpiOut {
  void* out;                 // 16-bit / f32 NHWC [.., cout]
  const uint16_t* residual;  // or null
  int32_t out_f32, relu, cout;
};

// two fp32 -> packed 16-bit pair (lo in the low half), optional ReLU, in one F2FP
template <bool RELU>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
#if VSB_ACT_F16
  if (RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
#else
  if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
#endif
  return r;
}

template <bool RELU>
__device__ __forceinline__ void store8(const EpiOut& p, const float (&f)[8], int64_t pix,
                                       int ch, bool full8) {
  if (p.out_f32) {
    float* o = reinterpret_cast<float*>(p.out) + pix * p.cout + ch;
    if (full8 && (p.cout & 3) == 0) {
      *reinterpret_cast<float4*>(o) =
          RELU ? make_float4(fmaxf(f[0], 0.f), fmaxf(f[1], 0.f), fmaxf(f[2], 0.f), fmaxf(f[3], 0.f))
               : make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(o + 4) =
          RELU ? make_float4(fmaxf(f[4], 0.f), fmaxf(f[5], 0.f), fmaxf(f[6], 0.f), fmaxf(f[7], 0.f))
               : make_float4(f[4], f[5], f[6], f[7]);
    } else {
      for (int j = 0; j < 8 && ch + j < p.cout; ++j) o[j] = RELU ? fmaxf(f[j], 0.f) : f[j];
    }
  } else {
    uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + pix * p.cout + ch;
    if (full8 && (p.cout & 7) == 0) {
      uint4 pk;
      pk.x = pack2<RELU>(f[0], f[1]);
      pk.y = pack2<RELU>(f[2], f[3]);
      pk.z = pack2<RELU>(f[4], f[5]);
      pk.w = pack2<RELU>(f[6], f[7]);
      *reinterpret_cast<uint4*>(o) = pk;
    } else {
      for (int j = 0; j < 8 && ch + j < p.cout; ++j)
        o[j] = float_to_act(RELU ? fmaxf(f[j], 0.f) : f[j]);
    }
  }
}


// 8 accumulator columns of one pixel (channels ch..ch+7): + bias, + residual, ReLU, store.
__device__ __forceinline__ void epilogue_group8(const EpiOut& p, const uint32_t* v8, const float* bias_s,
                                                int64_t pix, int ch, const uint4* res_pre = nullptr) {
  if (ch >= p.cout) return;
  float f[8];
  const float4 b0 = *reinterpret_cast<const float4*>(bias_s + ch);
  const float4 b1 = *reinterpret_cast<const float4*>(bias_s + ch + 4);
  f[0] = __uint_as_float(v8[0]) + b0.x;
  f[1] = __uint_as_float(v8[1]) + b0.y;
  f[2] = __uint_as_float(v8[2]) + b0.z;
  f[3] = __uint_as_float(v8[3]) + b0.w;
  f[4] = __uint_as_float(v8[4]) + b1.x;
  f[5] = __uint_as_float(v8[5]) + b1.y;
  f[6] = __uint_as_float(v8[6]) + b1.z;
  f[7] = __uint_as_float(v8[7]) + b1.w;
  const bool full8 = ch + 8 <= p.cout;
  if (p.residual) {
    if (full8) {
      const uint4 rv = res_pre ? *res_pre : __ldg(reinterpret_cast<const uint4*>(p.residual + pix * p.cout + ch));
      const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 rf = unpack_act2(rw[j]);
        f[2 * j] += rf.x;
        f[2 * j + 1] += rf.y;
      }
    } else {
      for (int j = 0; j < 8 && ch + j < p.cout; ++j) f[j] += act_to_float(p.residual[pix * p.cout + ch + j]);
    }
  }
  if (p.relu) store8<true>(p, f, pix, ch, full8);
  else store8<false>(p, f, pix, ch, full8);
}

// 32 accumulator columns [c, c+32) of one pixel.
__device__ __forceinline__ void epilogue_chunk32(const EpiOut& p, const uint32_t (&v)[32], const float* bias_s,
                                                 int64_t pix, int ch_base, const uint4* res_pre = nullptr) {
#pragma unroll
  for (int g8 = 0; g8 < 4; ++g8)
    epilogue_group8(p, &v[g8 * 8], bias_s, pix, ch_base + g8 * 8, res_pre ? res_pre + g8 : nullptr);
}

// Issue the residual loads of one 32-channel chunk early (before the accumulator is ready)
// so their latency hides behind the main loop.  Only whole 8-channel groups are prefetched.
__device__ __forceinline__ bool prefetch_residual32(const EpiOut& p, int64_t pix, int ch_base, uint4 (&r)[4]) {
  if (!p.residual || ch_base + 32 > p.cout) return false;
#pragma unroll
  for (int g8 = 0; g8 < 4; ++g8)
    r[g8] = __ldg(reinterpret_cast<const uint4*>(p.residual + pix * p.cout + ch_base + g8 * 8));
  return true;
}

}  // namespace vsb
