// Bring-up micro-test (built and run on the GPU box): one tcgen05.mma.cta_group::2 tile.
// Two CTAs of a cluster hold 128 rows of A each and HALF of the rows of B each (same shared-memory
// offsets in both), the leader issues the M = 256 MMAs, tcgen05.commit multicasts the completion
// to a barrier in both CTAs, and each CTA reads its 128 accumulator lanes back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/cta2_mma_test tests/micro/cta2_mma_test.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

constexpr int K = 64;  // one SW128 slab: 64 fp16 = 128 bytes per row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {  // K-major, SWIZZLE_128B, SBO = 1024 B
  return (uint64_t)((addr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// N: GEMM columns (each CTA holds N / 2 rows of B)
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
cta2_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_s = smem;                     // 128 rows x 128 B
  uint8_t* b_s = smem + 16384;             // N / 2 rows x 128 B
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 16384 + 16384);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(done)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // operands, pre-swizzled (Swizzle<3,4,3>: 16-byte chunk c of row r at chunk c ^ (r & 7))
  for (int i = threadIdx.x; i < 128 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(a_s + r * 128 + ((c ^ (r & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(A + ((size_t)(rank * 128 + r) * K + c * 8));
  }
  for (int i = threadIdx.x; i < (N / 2) * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(b_s + r * 128 + ((c ^ (r & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(B + ((size_t)(rank * (N / 2) + r) * K + c * 8));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (rank == 0 && threadIdx.x == 0) {
    // kind::f16, D = f32, A = B = f16, K-major, M = 256, N
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t ad = desc_sw128(smem_u32(a_s)), bd = desc_sw128(smem_u32(b_s));
#pragma unroll
    for (int k = 0; k < K / 16; ++k) {
      const uint32_t acc = k ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(done)),
                 "h"((uint16_t)3)
                 : "memory");
  }
  // every thread waits for the completion in ITS CTA
  int ok = 0;
  for (int it = 0; it < 4000000 && !ok; ++it) {
    uint32_t p;
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(p)
                 : "r"(smem_u32(done))
                 : "memory");
    ok = p;
  }
  if (threadIdx.x == 0) status[rank] = ok;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (ok) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) D[(size_t)(rank * 128 + row) * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(N) : "memory");
}

template <int N>
int run() {
  std::vector<__half> a((size_t)256 * K), b((size_t)N * K);
  for (size_t i = 0; i < a.size(); ++i) a[i] = __float2half((float)((int)((i * 7 + i / K) % 13) - 6));
  for (size_t i = 0; i < b.size(); ++i) b[i] = __float2half((float)((int)((i * 5 + 3 * (i / K)) % 11) - 5));
  __half *da, *db;
  float* dd;
  int* ds;
  CK(cudaMalloc(&da, a.size() * 2));
  CK(cudaMalloc(&db, b.size() * 2));
  CK(cudaMalloc(&dd, (size_t)256 * N * 4));
  CK(cudaMalloc(&ds, 8));
  CK(cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xff, (size_t)256 * N * 4));
  CK(cudaMemset(ds, 0, 8));
  CK(cudaFuncSetAttribute(cta2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  cta2_kernel<N><<<2, 128, 40960>>>(da, db, dd, ds);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> d((size_t)256 * N);
  int st[2];
  CK(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(st, ds, 8, cudaMemcpyDeviceToHost));
  long bad = 0;
  double worst = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0;
      for (int k = 0; k < K; ++k) ref += __half2float(a[(size_t)m * K + k]) * __half2float(b[(size_t)n * K + k]);
      const double e = fabs((double)ref - (double)d[(size_t)m * N + n]);
      if (!(e <= 1e-3)) ++bad;
      if (e > worst) worst = e;
    }
  printf("N=%d: barrier seen by CTA0 %d CTA1 %d, mismatches %ld of %d, worst %.3g\n", N, st[0], st[1], bad, 256 * N, worst);
  return bad != 0;
}

int main() {
  int rc = run<128>();
  rc |= run<256>();
  rc |= run<64>();
  printf(rc ? "FAILED\n" : "OK\n");
  return rc;
}
