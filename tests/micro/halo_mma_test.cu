// Micro-tests (standalone, run by hand under gpurun):
//  E3: does a UMMA K-major SW128 descriptor whose start address is shifted by whole
//      128-byte rows (a 3x3 tap window over a halo tile) read the rows TMA-style
//      swizzling wrote?  Which base_offset does it need?  SBO = halo row pitch.
//  E1: TMA 5-D box load throughput for 128-row boxes of 32 / 64 / 128-byte rows.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../volume_segmantics_b200/csrc
//        halo_mma_test.cu -o halo_mma_test
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define VSB_ACT_F16 1
#include "common.cuh"

using namespace vsb;

#define CHECK(x)                                                                  \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      printf("CUDA error %s at line %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                    \
    }                                                                             \
  } while (0)

// ---------------------------------------------------------------------------------- E3
constexpr int HW_ = 10, HH_ = 18;  // halo tile: 10 wide x 18 tall pixels, 64 channels (128 B)
constexpr int NOUT = 64;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
  return (uint64_t)((addr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
}

// in : halo [18][10][64] half, w: [9][64 n][64 c] half, out: [128 px][64 n] float; mode: base_offset policy
__global__ void __launch_bounds__(128, 1) halo_kernel(const __half* in, const __half* w, float* out, int mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* halo = smem;                       // 180 rows x 128 B = 23040 B
  uint8_t* wts = smem + 24 * 1024;            // 9 x 64 rows x 128 B
  uint64_t* bar = (uint64_t*)(smem + 24 * 1024 + 9 * 8192);
  uint32_t* tmem_slot = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x;
  // write with the swizzle TMA would apply: 16-byte chunk c of the 128-byte row at absolute
  // smem address A goes to chunk c ^ ((A >> 7) & 7)
  for (int i = tid; i < HH_ * HW_ * 8; i += 128) {
    const int row = i >> 3, c = i & 7;
    const uint32_t rowaddr = smem_u32(halo) + row * 128;
    const int pc = c ^ ((rowaddr >> 7) & 7);
    *(uint4*)(halo + row * 128 + pc * 16) = *(const uint4*)(in + row * 64 + c * 8);
  }
  for (int i = tid; i < 9 * 64 * 8; i += 128) {
    const int row = i >> 3, c = i & 7;
    const uint32_t rowaddr = smem_u32(wts) + row * 128;
    const int pc = c ^ ((rowaddr >> 7) & 7);
    *(uint4*)(wts + row * 128 + pc * 16) = *(const uint4*)(w + row * 64 + c * 8);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (tid < 32) tmem_alloc<64>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_act(128, NOUT);
    bool first = true;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const uint32_t a0 = smem_u32(halo) + (ky * HW_ + kx) * 128;  // window origin for this tap
        const uint32_t b0 = smem_u32(wts) + (ky * 3 + kx) * 8192;
        const uint32_t boff = mode == 0 ? 0 : ((a0 >> 7) & 7);
        for (int k = 0; k < 4; ++k) {
          umma_bf16_ss(tmem, desc_sw128(a0 + k * 32, HW_ * 128, boff), desc_sw128(b0 + k * 32, 1024, 0), idesc,
                       first ? 0u : 1u);
          first = false;
        }
      }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after_sync();
  const int warp = tid >> 5, lane = tid & 31;
  for (int c = 0; c < NOUT; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * NOUT + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc<64>(tmem);
}

static void run_e3() {
  std::vector<__half> in(HH_ * HW_ * 64), w(9 * 64 * 64);
  std::vector<float> inf(in.size()), wf(w.size());
  srand(1);
  for (size_t i = 0; i < in.size(); ++i) { inf[i] = (rand() % 17 - 8) / 8.f; in[i] = __float2half(inf[i]); }
  for (size_t i = 0; i < w.size(); ++i) { wf[i] = (rand() % 15 - 7) / 16.f; w[i] = __float2half(wf[i]); }
  // reference: out[m = y*8 + x][n] = sum_{ky,kx,c} in[(y+ky)*10 + (x+kx)][c] * w[ky*3+kx][n][c]
  std::vector<float> ref(128 * 64, 0.f);
  for (int y = 0; y < 16; ++y)
    for (int x = 0; x < 8; ++x)
      for (int n = 0; n < 64; ++n) {
        float s = 0;
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx)
            for (int c = 0; c < 64; ++c)
              s += inf[((y + ky) * HW_ + x + kx) * 64 + c] * wf[((ky * 3 + kx) * 64 + n) * 64 + c];
        ref[(y * 8 + x) * 64 + n] = s;
      }
  __half *din, *dw;
  float* dout;
  CHECK(cudaMalloc(&din, in.size() * 2));
  CHECK(cudaMalloc(&dw, w.size() * 2));
  CHECK(cudaMalloc(&dout, ref.size() * 4));
  CHECK(cudaMemcpy(din, in.data(), in.size() * 2, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(dw, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
  CHECK(cudaFuncSetAttribute(halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
  for (int mode = 0; mode < 2; ++mode) {
    CHECK(cudaMemset(dout, 0, ref.size() * 4));
    halo_kernel<<<1, 128, 110 * 1024>>>(din, dw, dout, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("E3 mode %d: launch failed: %s\n", mode, cudaGetErrorString(e)); exit(1); }
    std::vector<float> got(ref.size());
    CHECK(cudaMemcpy(got.data(), dout, ref.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    int bad = 0;
    for (size_t i = 0; i < ref.size(); ++i) {
      const double d = fabs(got[i] - ref[i]);
      if (d > maxerr) maxerr = d;
      if (d > 1e-2) ++bad;
    }
    printf("E3 halo-window MMA, base_offset %s: max err %.5f, bad %d / %zu -> %s\n",
           mode == 0 ? "= 0" : "= (start>>7)&7", maxerr, bad, ref.size(), bad ? "MISMATCH" : "OK");
  }
}

// ---------------------------------------------------------------------------------- E1
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// one producer lane streams `iters` boxes through a ring of `stages`; one consumer lane frees them.
__global__ void __launch_bounds__(64, 1) tma_bw_kernel(const __grid_constant__ CUtensorMap map, int iters, int stages,
                                                        int box_bytes, int tiles_x, int tiles_y, int bw, int bh,
                                                        long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + 200 * 1024);
  uint64_t* empty = full + 16;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    int stage = 0; uint32_t phase = 0;
    int t = blockIdx.x;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_arrive_expect_tx(&full[stage], box_bytes);
      const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, tn = t / (tiles_x * tiles_y);
      tma_load_5d(&map, &full[stage], smem + (size_t)stage * box_bytes, 0, tx * bw - 1, 0, ty * bh - 1, tn);
      t += gridDim.x;
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&full[stage], phase);
      mbar_arrive(&empty[stage]);
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

static void run_e1() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  const int NB = 8, H = 1024, W = 1024;
  CHECK(cudaFuncSetAttribute(tma_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  long long* dcyc;
  CHECK(cudaMalloc(&dcyc, 148 * 8));
  struct Cfg { int C, bw, bh; const char* name; };
  const Cfg cfgs[] = {{64, 16, 8, "C=64 box 16x8 (128 rows x 128 B)"}, {32, 16, 8, "C=32 box 16x8 (128 rows x 64 B)"},
                      {16, 16, 8, "C=16 box 16x8 (128 rows x 32 B)"}, {64, 10, 18, "C=64 halo box 10x18 (180 rows x 128 B)"},
                      {16, 128, 1, "C=16 box 128x1 (contiguous 4 KB)"}, {64, 8, 16, "C=64 box 8x16"}};
  for (const Cfg& c : cfgs) {
    void* dten;
    const size_t bytes = (size_t)NB * H * W * c.C * 2;
    CHECK(cudaMalloc(&dten, bytes));
    CHECK(cudaMemset(dten, 1, bytes));
    CUtensorMap m;
    cuuint64_t dims[5] = {(cuuint64_t)c.C, (cuuint64_t)W, 1, (cuuint64_t)H, (cuuint64_t)NB};
    cuuint64_t strides[4] = {(cuuint64_t)c.C * 2, (cuuint64_t)W * c.C * 2, (cuuint64_t)W * c.C * 2, (cuuint64_t)H * W * c.C * 2};
    cuuint32_t box[5] = {(cuuint32_t)c.C, (cuuint32_t)c.bw, 1, (cuuint32_t)c.bh, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    const CUtensorMapSwizzle sw = c.C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (c.C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, dten, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d for %s\n", (int)r, c.name); continue; }
    const int box_bytes = c.C * 2 * c.bw * c.bh;
    const int tiles_x = W / c.bw, tiles_y = H / c.bh;
    const int total = tiles_x * tiles_y * NB;
    const int iters = total / 148 < 4000 ? total / 148 : 4000;
    for (int stages : {2, 8}) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      tma_bw_kernel<<<148, 64, 210 * 1024>>>(m, iters, stages, box_bytes, tiles_x, tiles_y, c.bw, c.bh, dcyc);
      CHECK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      tma_bw_kernel<<<148, 64, 210 * 1024>>>(m, iters, stages, box_bytes, tiles_x, tiles_y, c.bw, c.bh, dcyc);
      cudaEventRecord(e1);
      CHECK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      long long cyc[148];
      CHECK(cudaMemcpy(cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost));
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += cyc[i];
      avg /= 148;
      printf("E1 %-42s stages %d: %.1f cycles/box, %.0f GB/s aggregate (%d boxes/SM, %.3f ms)\n", c.name, stages,
             avg / iters, (double)box_bytes * iters * 148 / (ms * 1e6), iters, ms);
    }
    cudaFree(dten);
  }
}

// ---------------------------------------------------------------------------------- E4
// Compact pitch: C = 16 (32-byte pixels, SW32) and C = 32 (64-byte pixels, SW64) halo tiles,
// 34 x 18 pixels (four 8-wide M tiles side by side), absolute-address swizzle on write.
template <int C>
__global__ void __launch_bounds__(128, 1) halo_compact_kernel(const __half* in, const __half* w, float* out) {
  constexpr int P = C * 2;            // pixel pitch in bytes
  constexpr int HWc = 34, HHc = 18, MT = 4, N = 16;
  constexpr int LAYOUT = C == 16 ? 6 : 4;  // SW32 : SW64
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* halo = smem;                 // 612 px * P
  uint8_t* wts = smem + 48 * 1024;      // 9 taps x 16 rows x P
  uint64_t* bar = (uint64_t*)(smem + 60 * 1024);
  uint32_t* tmem_slot = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x;
  auto swz = [](uint32_t addr) -> uint32_t {   // Swizzle<B,4,3> on the absolute address
    return C == 16 ? (addr ^ (((addr >> 7) & 1) << 4)) : (addr ^ (((addr >> 7) & 3) << 4));
  };
  for (int i = tid; i < HWc * HHc * (P / 16); i += 128) {
    const int pix = i / (P / 16), c = i % (P / 16);
    const uint32_t a = smem_u32(halo) + pix * P + c * 16;
    *(uint4*)(halo + (swz(a) - smem_u32(halo))) = *(const uint4*)(in + pix * C + c * 8);
  }
  for (int i = tid; i < 9 * N * (P / 16); i += 128) {
    const int row = i / (P / 16), c = i % (P / 16);
    const uint32_t a = smem_u32(wts) + row * P + c * 16;
    *(uint4*)(wts + (swz(a) - smem_u32(wts))) = *(const uint4*)(w + row * C + c * 8);
  }
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (tid < 32) tmem_alloc<64>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  auto desc = [](uint32_t addr, uint32_t sbo) -> uint64_t {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
           ((uint64_t)LAYOUT << 61);
  };
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_act(128, N);
    for (int mt = 0; mt < MT; ++mt) {
      bool first = true;
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          const uint32_t a0 = smem_u32(halo) + (ky * HWc + kx + mt * 8) * P;
          const uint32_t b0 = smem_u32(wts) + (ky * 3 + kx) * N * P;
          for (int k = 0; k < C / 16; ++k) {
            umma_bf16_ss(tmem + mt * N, desc(a0 + k * 32, HWc * P), desc(b0 + k * 32, 8 * P), idesc, first ? 0u : 1u);
            first = false;
          }
        }
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after_sync();
  const int warp = tid >> 5, lane = tid & 31;
  for (int c = 0; c < MT * N; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * (MT * N) + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc<64>(tmem);
}

template <int C>
static void run_e4() {
  constexpr int HWc = 34, HHc = 18, MT = 4, N = 16;
  std::vector<__half> in(HHc * HWc * C), w(9 * N * C);
  std::vector<float> inf(in.size()), wf(w.size());
  srand(7);
  for (size_t i = 0; i < in.size(); ++i) { inf[i] = (rand() % 17 - 8) / 8.f; in[i] = __float2half(inf[i]); }
  for (size_t i = 0; i < w.size(); ++i) { wf[i] = (rand() % 15 - 7) / 16.f; w[i] = __float2half(wf[i]); }
  std::vector<float> ref(128 * MT * N, 0.f);
  for (int mt = 0; mt < MT; ++mt)
    for (int y = 0; y < 16; ++y)
      for (int x = 0; x < 8; ++x)
        for (int n = 0; n < N; ++n) {
          float s = 0;
          for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx)
              for (int c = 0; c < C; ++c)
                s += inf[((y + ky) * HWc + mt * 8 + x + kx) * C + c] * wf[((ky * 3 + kx) * N + n) * C + c];
          ref[(y * 8 + x) * (MT * N) + mt * N + n] = s;
        }
  __half *din, *dw;
  float* dout;
  CHECK(cudaMalloc(&din, in.size() * 2));
  CHECK(cudaMalloc(&dw, w.size() * 2));
  CHECK(cudaMalloc(&dout, ref.size() * 4));
  CHECK(cudaMemcpy(din, in.data(), in.size() * 2, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(dw, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
  CHECK(cudaMemset(dout, 0, ref.size() * 4));
  CHECK(cudaFuncSetAttribute(halo_compact_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
  halo_compact_kernel<C><<<1, 128, 70 * 1024>>>(din, dw, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("E4 C=%d: launch failed: %s\n", C, cudaGetErrorString(e)); exit(1); }
  std::vector<float> got(ref.size());
  CHECK(cudaMemcpy(got.data(), dout, ref.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0; int bad = 0;
  for (size_t i = 0; i < ref.size(); ++i) { const double d = fabs(got[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > 1e-2) ++bad; }
  printf("E4 compact pitch C=%d (%s), 4 M-tiles, shifted windows: max err %.5f, bad %d / %zu -> %s\n", C,
         C == 16 ? "SW32" : "SW64", maxerr, bad, ref.size(), bad ? "MISMATCH" : "OK");
}

int main() {
  run_e3();
  run_e4<16>();
  run_e4<32>();
  if (getenv("RUN_E1")) run_e1();
  return 0;
}
