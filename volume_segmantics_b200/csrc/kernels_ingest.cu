// Ingest-side kernels of the volseg-b200 engine:
//   * typed slicer: volumes that are not uint8 are sliced, padded (reflect-101) and normalised
//     exactly as datasets.py:129-135 does -- integers of any depth become float32 and are
//     divided by 255, floats are taken as they are, then (x - 0.449) / 0.226 in fp32;
//   * volume moments: nanmean / nanstd (base_data_manager.py:33, base_data_utils.py:255) as a
//     two-pass float64 reduction with a FIXED reduction order (reproducible run to run);
//   * clip_to_uint8 (base_data_utils.py:243-287): NaN -> mean, clip, rescale, truncate in the
//     arithmetic numpy uses for the input dtype (float32 for float32 data, float64 otherwise),
//     with the clipped-voxel counts of the reference's log lines gathered in the same pass.
#include "common.cuh"
#include "kernels.h"

namespace vsb {

namespace {

__host__ __device__ inline int64_t reflect101_i(int64_t p, int64_t len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * len - 2 - p;
  }
  return p;
}

// dtype codes shared with vsb_clip_to_uint8: 0 f32, 1 f64, 2 u8, 3 i8, 4 u16, 5 i16, 6 u32, 7 i32, 8 i64
template <typename T>
struct IsFloat { static constexpr bool value = false; };
template <>
struct IsFloat<float> { static constexpr bool value = true; };

// datasets.py:129-135 on one voxel, fp32 throughout (numpy: image.astype(np.float32) / 255,
// then - 0.449, then / 0.226, each a correctly rounded float32 operation)
template <typename T>
__device__ __forceinline__ float normalise_voxel(T v) {
  float f;
  if (IsFloat<T>::value) {
    f = (float)v;
  } else {
    f = __fdiv_rn((float)v, 255.0f);  // int -> float32 is round-to-nearest-even, as astype
  }
  f = __fsub_rn(f, 0.449f);
  return __fdiv_rn(f, 0.226f);
}

// Row directions (image columns contiguous in the volume): one thread = 8 consecutive padded
// pixels of one image row, one 16-byte store.
template <typename T>
__global__ void __launch_bounds__(256) slicer_typed_rows_kernel(const T* __restrict__ vol, vsb_direction g,
                                                                int64_t s0, int nb, uint16_t* __restrict__ out) {
  const int64_t chunks = g.Wp >> 3;
  const int64_t total = (int64_t)nb * g.Hp * chunks;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ch = i % chunks;
    const int64_t pr = (i / chunks) % g.Hp;
    const int64_t s = i / (chunks * g.Hp);
    const int64_t r = reflect101_i(pr - g.pad_top, g.H);
    const T* rowp = vol + g.base + (s0 + s) * g.stride_s + r * g.stride_r;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t c = reflect101_i(ch * 8 + j - g.pad_left, g.W);
      f[j] = normalise_voxel<T>(rowp[c * g.stride_c]);
    }
    uint4 o;
    o.x = pack_act2(f[0], f[1]);
    o.y = pack_act2(f[2], f[3]);
    o.z = pack_act2(f[4], f[5]);
    o.w = pack_act2(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + ((s * g.Hp + pr) * g.Wp + ch * 8)) = o;
  }
}

// X-plane directions (slice index contiguous in the volume): a 32 (slices) x 32 (columns) tile of
// one padded image row is read with the slice index fastest and written with the column fastest.
template <typename T>
__global__ void __launch_bounds__(256) slicer_typed_xplane_kernel(const T* __restrict__ vol, vsb_direction g,
                                                                  int64_t s0, int nb, uint16_t* __restrict__ out) {
  __shared__ uint16_t tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t ctiles = g.Wp >> 5, stiles = (nb + 31) >> 5;
  const int64_t total = stiles * g.Hp * ctiles;
  for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
    const int64_t ct = t % ctiles;
    const int64_t pr = (t / ctiles) % g.Hp;
    const int64_t st = t / (ctiles * g.Hp);
    const int64_t r = reflect101_i(pr - g.pad_top, g.H);
    __syncthreads();
    for (int ci = ty; ci < 32; ci += 8) {
      const int64_t s = st * 32 + tx;
      const int64_t c = reflect101_i(ct * 32 + ci - g.pad_left, g.W);
      uint16_t v = 0;
      if (s < nb) v = float_to_act(normalise_voxel<T>(vol[g.base + (s0 + s) * g.stride_s + r * g.stride_r + c * g.stride_c]));
      tile[ci][tx] = v;
    }
    __syncthreads();
    for (int si = ty; si < 32; si += 8) {
      const int64_t s = st * 32 + si;
      if (s < nb) out[(s * g.Hp + pr) * g.Wp + ct * 32 + tx] = tile[tx][si];
    }
  }
}

template <typename T>
void launch_slicer_typed_t(const void* vol, const vsb_direction& g, int64_t s0, int nb, uint16_t* out,
                           cudaStream_t st) {
  if (g.stride_s == 1 && nb >= 8) {
    const int64_t total = ((nb + 31) / 32) * g.Hp * (g.Wp / 32);
    slicer_typed_xplane_kernel<T><<<(int)(total < 148 * 16 ? total : 148 * 16), 256, 0, st>>>((const T*)vol, g, s0, nb, out);
    return;
  }
  const int64_t total = (int64_t)nb * g.Hp * (g.Wp / 8);
  const int64_t blocks = (total + 255) / 256;
  slicer_typed_rows_kernel<T><<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>((const T*)vol, g, s0, nb, out);
}

// ---- moments -----------------------------------------------------------------------------
constexpr int kMomBlocks = 148 * 4, kMomThreads = 256;

template <typename T>
__device__ __forceinline__ double as_double(T v) { return (double)v; }

// Fixed-order block reduction of (a, b): warp shuffles (xor tree), then warp 0 over the 8 warp sums.
__device__ __forceinline__ void block_reduce2(double& a, double& b) {
  __shared__ double sa[8], sb[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (w == 0) {
    a = l < 8 ? sa[l] : 0.0;
    b = l < 8 ? sb[l] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
  }
}

// pass 1: partial[2*b] = sum of the non-NaN values, partial[2*b+1] = their count (block b)
// pass 2: partial[2*b] = sum of (x - mean)^2 over the non-NaN values, partial[2*b+1] = number of NaNs
template <typename T, int PASS>
__global__ void __launch_bounds__(kMomThreads) moments_kernel(const T* __restrict__ x, int64_t n, double mean,
                                                              double* __restrict__ partial) {
  double a = 0.0, b = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * kMomThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kMomThreads) {
    const double v = as_double<T>(x[i]);
    const bool nan = v != v;
    if (PASS == 1) {
      if (!nan) { a += v; b += 1.0; }
    } else {
      if (!nan) { const double d = v - mean; a += d * d; }
      else b += 1.0;
    }
  }
  block_reduce2(a, b);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = a;
    partial[2 * blockIdx.x + 1] = b;
  }
}

template <typename T>
void launch_moments_t(const void* x, int64_t n, int pass, double mean, double* partial, cudaStream_t st) {
  if (pass == 1) moments_kernel<T, 1><<<kMomBlocks, kMomThreads, 0, st>>>((const T*)x, n, mean, partial);
  else moments_kernel<T, 2><<<kMomBlocks, kMomThreads, 0, st>>>((const T*)x, n, mean, partial);
}

// ---- clip + quantise + clipped-voxel counts --------------------------------------------------
// F = float: float32 data, every step a float32 operation (numpy keeps the array dtype);
// F = double: float64 data and integers (clip_to_uint8 casts integers with astype(float)).
// The comparisons for the two counts (`data > upper_bound`, `data < lower_bound`) are done in
// the precision numpy promotes to (array dtype vs the bound's scalar type), which for every
// supported dtype equals comparing the exactly converted values.
template <typename T, typename F>
__global__ void __launch_bounds__(256) clip_count_kernel(const T* __restrict__ in, int64_t n, F mean, F lower, F upper,
                                                         uint8_t* __restrict__ out, unsigned long long* __restrict__ counts) {
  const F range = upper - lower;
  unsigned long long gt = 0, lt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    F x = (F)in[i];
    gt += x > upper ? 1u : 0u;  // NaN compares false, as numpy under errstate(invalid="ignore")
    lt += x < lower ? 1u : 0u;
    if (x != x) x = mean;  // np.nan_to_num(nan=data_mean)
    if (sizeof(F) == 4) {
      float y = fminf(fmaxf((float)x, (float)lower), (float)upper);
      y = __fsub_rn(y, (float)lower);
      y = __fdiv_rn(y, (float)range);
      y = fminf(fmaxf(y, 0.0f), 1.0f);
      y = __fmul_rn(y, 255.0f);
      out[i] = (uint8_t)(long long)y;
    } else {
      double y = fmin(fmax((double)x, (double)lower), (double)upper);
      y = __dsub_rn(y, (double)lower);
      y = __ddiv_rn(y, (double)range);
      y = fmin(fmax(y, 0.0), 1.0);
      y = __dmul_rn(y, 255.0);
      out[i] = (uint8_t)(long long)y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gt += __shfl_xor_sync(0xffffffffu, gt, o);
    lt += __shfl_xor_sync(0xffffffffu, lt, o);
  }
  if ((threadIdx.x & 31) == 0) {  // integer counts: the order of the atomics does not matter
    if (gt) atomicAdd(counts, gt);
    if (lt) atomicAdd(counts + 1, lt);
  }
}

template <typename T>
void launch_clip_count_t(const void* in, int64_t n, double mean, double lower, double upper, uint8_t* out,
                         unsigned long long* counts, cudaStream_t st) {
  const int64_t blocks = (n + 255) / 256;
  const int grid = (int)(blocks < 148 * 32 ? blocks : 148 * 32);
  if (IsFloat<T>::value)
    clip_count_kernel<T, float><<<grid, 256, 0, st>>>((const T*)in, n, (float)mean, (float)lower, (float)upper, out, counts);
  else
    clip_count_kernel<T, double><<<grid, 256, 0, st>>>((const T*)in, n, mean, lower, upper, out, counts);
}

}  // namespace

#define VSB_DISPATCH_DTYPE(dtype, FN, ...)                 \
  switch (dtype) {                                         \
    case 0: FN<float>(__VA_ARGS__); break;                 \
    case 1: FN<double>(__VA_ARGS__); break;                \
    case 2: FN<uint8_t>(__VA_ARGS__); break;               \
    case 3: FN<int8_t>(__VA_ARGS__); break;                \
    case 4: FN<uint16_t>(__VA_ARGS__); break;              \
    case 5: FN<int16_t>(__VA_ARGS__); break;               \
    case 6: FN<uint32_t>(__VA_ARGS__); break;              \
    case 7: FN<int32_t>(__VA_ARGS__); break;               \
    default: FN<long long>(__VA_ARGS__); break;            \
  }

bool slicer_typed_supported(int dtype) { return dtype == 0 || dtype == 2 || dtype == 3 || dtype == 4 || dtype == 5 || dtype == 7; }

void launch_slicer_typed(const void* vol, int dtype, const vsb_direction& g, int64_t s0, int nb, uint16_t* out,
                         cudaStream_t st) {
  switch (dtype) {
    case 0: launch_slicer_typed_t<float>(vol, g, s0, nb, out, st); break;
    case 2: launch_slicer_typed_t<uint8_t>(vol, g, s0, nb, out, st); break;
    case 3: launch_slicer_typed_t<int8_t>(vol, g, s0, nb, out, st); break;
    case 4: launch_slicer_typed_t<uint16_t>(vol, g, s0, nb, out, st); break;
    case 5: launch_slicer_typed_t<int16_t>(vol, g, s0, nb, out, st); break;
    default: launch_slicer_typed_t<int32_t>(vol, g, s0, nb, out, st); break;
  }
}

int moments_partials() { return 2 * kMomBlocks; }

void launch_moments(const void* x, int dtype, int64_t n, int pass, double mean, double* partial, cudaStream_t st) {
  VSB_DISPATCH_DTYPE(dtype, launch_moments_t, x, n, pass, mean, partial, st);
}

void launch_clip_count(const void* in, int dtype, int64_t n, double mean, double lower, double upper, uint8_t* out,
                       unsigned long long* counts, cudaStream_t st) {
  VSB_DISPATCH_DTYPE(dtype, launch_clip_count_t, in, n, mean, lower, upper, out, counts, st);
}

}  // namespace vsb
