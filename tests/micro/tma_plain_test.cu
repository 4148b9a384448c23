// Bring-up micro-test (built and run on the GPU box): plain (no swizzle) TMA box loads of a 16-bit
// image at arbitrary x / y coordinates, and TMA stores at negative coordinates.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tma_plain_test tests/micro/tma_plain_test.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

struct alignas(64) Desc { uint8_t b[128]; };

__device__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void load_kernel(const __grid_constant__ Desc map, int rank, int x, int y, int n, uint32_t expect, uint16_t* out, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(expect) : "memory");
    if (rank == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem)),
                   "l"(&map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(n) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem)),
                   "l"(&map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(n), "r"(0), "r"(0) : "memory");
    int ok = 0;
    for (int it = 0; it < 2000000 && !ok; ++it) {
      uint32_t p;
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(p) : "r"(smem_u32(bar)) : "memory");
      ok = p;
    }
    *status = ok;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 38; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

__global__ void store_kernel(const __grid_constant__ Desc map, int c1, int c2, int c3, int c4, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < 16384 / 2; i += blockDim.x) reinterpret_cast<uint16_t*>(smem)[i] = (uint16_t)(i & 0x7fff);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(&map), "r"(smem_u32(smem)),
                 "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    *status = 1;
  }
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int W = 128, H = 96, NB = 2;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  Enc enc = (Enc)fn;
  std::vector<uint16_t> img((size_t)NB * H * W);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (uint16_t)(i % 30000);
  uint16_t* dimg; CK(cudaMalloc(&dimg, img.size() * 2)); CK(cudaMemcpy(dimg, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  uint16_t* dout; CK(cudaMalloc(&dout, 64 * 38 * 2));
  int* dstat; CK(cudaMalloc(&dstat, 4));
  CK(cudaFuncSetAttribute(load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  struct Case { int rank, x, y; } cases[] = {{3, 0, 0}, {3, 8, 3}, {3, 6, 3}, {3, -6, -5}, {5, 0, 0}, {5, 54, 7}, {5, -6, -5}, {3, 100, 80}};
  for (auto& c : cases) {
    CUtensorMap m;
    cuuint64_t d3[3] = {W, H, NB}, s3[2] = {W * 2, (cuuint64_t)H * W * 2};
    cuuint64_t d5[5] = {W, H, NB, 1, 1}, s5[4] = {W * 2, (cuuint64_t)H * W * 2, (cuuint64_t)NB * H * W * 2, (cuuint64_t)NB * H * W * 2};
    cuuint32_t b3[3] = {64, 38, 1}, b5[5] = {64, 38, 1, 1, 1};
    CUresult r = c.rank == 3 ? enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dimg, d3, s3, b3, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                             : enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, dimg, d5, s5, b5, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    Desc d; memcpy(&d, &m, 128);
    CK(cudaMemset(dstat, 0, 4));
    load_kernel<<<1, 128, 40960>>>(d, c.rank, c.x, c.y, 1, 64 * 38 * 2, dout, dstat);
    cudaError_t e = cudaDeviceSynchronize();
    int st = -1; std::vector<uint16_t> o(64 * 38);
    if (e == cudaSuccess) { cudaMemcpy(&st, dstat, 4, cudaMemcpyDeviceToHost); cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost); }
    int bad = 0;
    for (int yy = 0; yy < 38; ++yy) for (int xx = 0; xx < 64; ++xx) {
      const int gx = c.x + xx, gy = c.y + yy;
      const uint16_t want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? img[((size_t)1 * H + gy) * W + gx] : 0;
      bad += o[yy * 64 + xx] != want;
    }
    printf("load rank %d at (%d,%d): encode %d, sync %s, barrier completed %d, mismatches %d\n", c.rank, c.x, c.y, (int)r, cudaGetErrorString(e), st, bad);
    if (e != cudaSuccess) return 2;
  }
  // stores at negative coordinates through a (c, x%4, x/4, y, n) view
  {
    const int Wc = 64, Hc = 48;
    uint16_t* dt; CK(cudaMalloc(&dt, (size_t)NB * Hc * Wc * 64 * 2)); CK(cudaMemset(dt, 0, (size_t)NB * Hc * Wc * 64 * 2));
    cuuint64_t od[5] = {64, 4, Wc / 4, Hc, NB}, os[4] = {128, 512, (cuuint64_t)Wc * 128, (cuuint64_t)Hc * Wc * 128};
    cuuint32_t ob[5] = {64, 1, 8, 14, 1};
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, dt, od, os, ob, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    Desc d; memcpy(&d, &m, 128);
    int coords[][4] = {{0, 0, 0, 0}, {3, 1, 2, 1}, {3, -1, 0, 0}};
    for (auto& c : coords) {
      CK(cudaMemset(dstat, 0, 4));
      store_kernel<<<1, 128, 40960>>>(d, c[0], c[1], c[2], c[3], dstat);
      cudaError_t e = cudaDeviceSynchronize();
      int st = -1; if (e == cudaSuccess) cudaMemcpy(&st, dstat, 4, cudaMemcpyDeviceToHost);
      printf("store at (xr %d, xi %d, y %d, n %d): encode %d, sync %s, done %d\n", c[0], c[1], c[2], c[3], (int)r, cudaGetErrorString(e), st);
      fflush(stdout);
      if (e != cudaSuccess) return 3;
    }
  }
  return 0;
}
