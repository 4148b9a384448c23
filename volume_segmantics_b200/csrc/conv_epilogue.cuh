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


// Generic (slow-path) 8-channel group: f32 output, partial groups, odd channel counts.
// Kept out of line so the hot epilogue loop stays small (the fully inlined version
// suffered instruction-cache misses -- ncu "no_inst" stalls).
// Values are passed in registers (not by pointer) so the caller's accumulator array is
// never forced into local memory.
static __device__ __noinline__ void epilogue_group8_generic(EpiOut p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                     uint32_t a4, uint32_t a5, uint32_t a6, uint32_t a7,
                                                     const float* bias_s, int64_t pix, int ch) {
  if (ch >= p.cout) return;
  const uint32_t v8[8] = {a0, a1, a2, a3, a4, a5, a6, a7};
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v8[j]) + bias_s[ch + j];
  const bool full8 = ch + 8 <= p.cout;
  if (p.residual) {
    for (int j = 0; j < 8 && ch + j < p.cout; ++j) f[j] += act_to_float(p.residual[pix * p.cout + ch + j]);
  }
  if (p.relu) store8<true>(p, f, pix, ch, full8);
  else store8<false>(p, f, pix, ch, full8);
}

// 8 accumulator columns of one pixel (channels ch..ch+7): + bias, + residual, ReLU, store.
// Fast path: 16-bit output, cout % 8 == 0 (whole group in range).
// `pre` / `use_pre`: residual chunk already in registers (prefetched before the accumulator was ready); passed
// by value so the caller's prefetch arrays are never address-taken (a pointer select put them in local memory).
__device__ __forceinline__ void epilogue_group8(const EpiOut& p, const uint32_t* v8, const float* bias_s,
                                                int64_t pix, int ch, uint4 pre = make_uint4(0u, 0u, 0u, 0u),
                                                bool use_pre = false) {
  if (p.out_f32 || (p.cout & 7)) {
    epilogue_group8_generic(p, v8[0], v8[1], v8[2], v8[3], v8[4], v8[5], v8[6], v8[7], bias_s, pix, ch);
    return;
  }
  if (ch >= p.cout) return;
  const float4 b0 = *reinterpret_cast<const float4*>(bias_s + ch);
  const float4 b1 = *reinterpret_cast<const float4*>(bias_s + ch + 4);
  float f0 = __uint_as_float(v8[0]) + b0.x, f1 = __uint_as_float(v8[1]) + b0.y;
  float f2 = __uint_as_float(v8[2]) + b0.z, f3 = __uint_as_float(v8[3]) + b0.w;
  float f4 = __uint_as_float(v8[4]) + b1.x, f5 = __uint_as_float(v8[5]) + b1.y;
  float f6 = __uint_as_float(v8[6]) + b1.z, f7 = __uint_as_float(v8[7]) + b1.w;
  const int64_t off = pix * p.cout + ch;
  if (p.residual) {
    const uint4 rv = use_pre ? pre : __ldg(reinterpret_cast<const uint4*>(p.residual + off));
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
  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + off) = pk;
}

// 32 accumulator columns [c, c+32) of one pixel.
__device__ __forceinline__ void epilogue_chunk32(const EpiOut& p, const uint32_t (&v)[32], const float* bias_s,
                                                 int64_t pix, int ch_base) {
#pragma unroll
  for (int g8 = 0; g8 < 4; ++g8) epilogue_group8(p, &v[g8 * 8], bias_s, pix, ch_base + g8 * 8);
}
// Same with the residual of the chunk prefetched into `pre` (valid when use_pre).
__device__ __forceinline__ void epilogue_chunk32_pre(const EpiOut& p, const uint32_t (&v)[32], const float* bias_s,
                                                     int64_t pix, int ch_base, const uint4 (&pre)[4], bool use_pre) {
#pragma unroll
  for (int g8 = 0; g8 < 4; ++g8) epilogue_group8(p, &v[g8 * 8], bias_s, pix, ch_base + g8 * 8, pre[g8], use_pre);
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
