// Head kernels for space-to-depth logits (plan.py "S2D tail"): logits f32
// [nb, Hp/2, Wp/2, 4*C], channel (2a+b)*C + k = class k of padded pixel (2i+a, 2j+b).
// Same per-voxel arithmetic as head_kernel (kernels_simple.cu): fp32 softmax, first-max
// label, fp16 RNE probability, torchvision centre crop, inverse axis/rot90 mapping,
// packed-key atomicMax (vol_seg_2d_predictor.py:45-64, 90-98) -- but one thread owns a
// 2x2 voxel block, the class count is a template parameter (everything stays in
// registers) and logits move as 16-byte vectors, so the kernel runs at HBM speed:
// 4*C B logits + 8 B key read + 8 B key written per voxel.
#include "common.cuh"
#include "kernels.h"

namespace vsb {
namespace {

// softmax over C logits -> probability and index of the FIRST maximum probability, with
// exactly the values head_pixel() produces: e_k = expf(l_k - max), sum in ascending k,
// p_k = e_k / sum (IEEE), first k attaining max p.  p is monotonic in e, so max p belongs
// to max e; an earlier class can only tie after rounding when its e is within an ulp or
// two of the maximum -- only those are divided.
template <int C>
__device__ __forceinline__ void softmax_first_max(const float (&l)[C], float& best, int& lab) {
  float m = l[0];
#pragma unroll
  for (int k = 1; k < C; ++k) m = fmaxf(m, l[k]);
  float e[C];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < C; ++k) {
    e[k] = expf(l[k] - m);
    sum += e[k];
  }
  float em = e[0];
  int km = 0;
#pragma unroll
  for (int k = 1; k < C; ++k)
    if (e[k] > em) {
      em = e[k];
      km = k;
    }
  best = __fdiv_rn(em, sum);
  lab = km;
  const float near = em * 0.99999f;
#pragma unroll
  for (int k = C - 2; k >= 0; --k)
    if (k < km && e[k] >= near && __fdiv_rn(e[k], sum) == best) lab = k;
}

// 4*C consecutive floats of one S2D pixel -> l[sub-pixel][class]
template <int C>
__device__ __forceinline__ void load_s2d_pixel(const float* __restrict__ p, float (&l)[4][C]) {
  float t[4 * C];
#pragma unroll
  for (int v = 0; v < C; ++v) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(p) + v);
    t[4 * v] = f.x;
    t[4 * v + 1] = f.y;
    t[4 * v + 2] = f.z;
    t[4 * v + 3] = f.w;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int k = 0; k < C; ++k) l[q][k] = t[q * C + k];
}

// Directions whose image columns are contiguous voxels (or anything when the batch is too
// small to transpose): one thread = one S2D pixel, lanes along the image row.
template <int C>
__global__ void __launch_bounds__(256) head_s2d_rows_kernel(HeadArgs a, int n0, uint32_t total) {
  const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
  if (idx >= total) return;
  const vsb_direction& g = a.g;
  const uint32_t Hh = (uint32_t)(g.Hp >> 1), Wh = (uint32_t)(g.Wp >> 1);
  const uint32_t rowid = idx / Wh, j = idx - rowid * Wh;
  const uint32_t nl = rowid / Hh, i = rowid - nl * Hh;
  const int64_t n = (int64_t)n0 + nl;
  float l[4][C];
  load_s2d_pixel<C>(a.logits + (((int64_t)n * Hh + i) * Wh + j) * (4 * C), l);
  const int r0 = 2 * (int)i - (int)g.crop_top, c0 = 2 * (int)j - (int)g.crop_left;
  const int64_t vox0 = g.base + (a.s0 + n) * g.stride_s;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int r = r0 + (q >> 1), c = c0 + (q & 1);
    if ((unsigned)r >= (unsigned)g.H || (unsigned)c >= (unsigned)g.W) continue;
    float best;
    int lab;
    softmax_first_max<C>(l[q], best, lab);
    const int64_t vox = vox0 + (int64_t)r * g.stride_r + (int64_t)c * g.stride_c;
    if (a.votes) {
      const int64_t vi = (int64_t)lab * a.nvox + vox;
      unsigned int* word = reinterpret_cast<unsigned int*>(a.votes + (vi & ~3ll));
      atomicAdd(word, 1u << (8 * (vi & 3)));
    } else {
      atomicMax(a.keys + vox, pack_key(__half_as_ushort(__float2half_rn(best)), a.d, (uint32_t)lab,
                                       __float_as_uint(best)));
    }
  }
}

// x-plane directions (the slice index runs along x, stride_s == 1): a block takes 32 slices
// x one S2D row x 16 S2D columns, reads the logits with the column fastest, transposes the
// 32 x 2 x 32 keys through shared memory and merges them with the slice index fastest, so
// each warp-wide atomic covers 256 contiguous bytes of the key volume.
template <int C>
__global__ void __launch_bounds__(256) head_s2d_xplane_kernel(HeadArgs a) {
  __shared__ unsigned long long tile[64][33];
  const vsb_direction& g = a.g;
  const int Hh = (int)(g.Hp >> 1), Wh = (int)(g.Wp >> 1);
  const int jt = (Wh + 15) >> 4;
  const int jl = threadIdx.x & 15, ns = threadIdx.x >> 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t stiles = (a.nb + 31) >> 5;
  const int64_t total = stiles * Hh * jt;
  for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
    const int jtile = (int)(t % jt);
    const int i = (int)((t / jt) % Hh);
    const int64_t stile = t / ((int64_t)jt * Hh);
    const int j = jtile * 16 + jl;
    const int r0 = 2 * i - (int)g.crop_top, c0 = 2 * j - (int)g.crop_left;
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int sl = ns + 16 * h;
      const int64_t n = stile * 32 + sl;
      unsigned long long key[4] = {0ull, 0ull, 0ull, 0ull};
      if (n < a.nb && j < Wh) {
        float l[4][C];
        load_s2d_pixel<C>(a.logits + ((n * Hh + i) * (int64_t)Wh + j) * (4 * C), l);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = r0 + (q >> 1), c = c0 + (q & 1);
          if ((unsigned)r < (unsigned)g.H && (unsigned)c < (unsigned)g.W) {
            float best;
            int lab;
            softmax_first_max<C>(l[q], best, lab);
            key[q] = pack_key(__half_as_ushort(__float2half_rn(best)), a.d, (uint32_t)lab, __float_as_uint(best));
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) tile[(q >> 1) * 32 + 2 * jl + (q & 1)][sl] = key[q];
    }
    __syncthreads();
    const int64_t n = stile * 32 + lane;
    const int64_t vox0 = g.base + (a.s0 + n) * g.stride_s;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rc = warp + 8 * it;
      const unsigned long long key = tile[rc][lane];
      if (key) {
        const int r = r0 + (rc >> 5), c = 2 * jtile * 16 - (int)g.crop_left + (rc & 31);
        atomicMax(a.keys + (vox0 + (int64_t)r * g.stride_r + (int64_t)c * g.stride_c), key);
      }
    }
  }
}

template <int C>
void launch_head_s2d_c(const HeadArgs& a, cudaStream_t st) {
  const int64_t Hh = a.g.Hp >> 1, Wh = a.g.Wp >> 1;
  if (!a.votes && a.g.stride_s == 1 && a.g.stride_c != 1 && a.nb >= 8) {
    const int64_t tiles = ((a.nb + 31) / 32) * Hh * ((Wh + 15) / 16);
    head_s2d_xplane_kernel<C><<<(int)(tiles < 148 * 32 ? tiles : 148 * 32), 256, 0, st>>>(a);
    return;
  }
  // 32-bit indexing inside the kernel: split the batch so a launch stays below 2^30 S2D pixels
  const int64_t per_img = Hh * Wh;
  const int64_t chunk = per_img >= (1ll << 30) ? 1 : (1ll << 30) / per_img;
  for (int64_t n0 = 0; n0 < a.nb; n0 += chunk) {
    const int64_t cnt = (a.nb - n0 < chunk ? a.nb - n0 : chunk) * per_img;
    head_s2d_rows_kernel<C><<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(a, (int)n0, (uint32_t)cnt);
  }
}

}  // namespace

bool launch_head_s2d(const HeadArgs& a, cudaStream_t st) {
  switch (a.C) {
    case 1: launch_head_s2d_c<1>(a, st); return true;
    case 2: launch_head_s2d_c<2>(a, st); return true;
    case 3: launch_head_s2d_c<3>(a, st); return true;
    case 4: launch_head_s2d_c<4>(a, st); return true;
    case 5: launch_head_s2d_c<5>(a, st); return true;
    case 6: launch_head_s2d_c<6>(a, st); return true;
    case 7: launch_head_s2d_c<7>(a, st); return true;
    case 8: launch_head_s2d_c<8>(a, st); return true;
    default: return false;
  }
}

}  // namespace vsb
