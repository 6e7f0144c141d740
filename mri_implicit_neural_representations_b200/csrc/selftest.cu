// Self-test of the UMMA building block: one tcgen05 GEMM per operand-layout family used by the
// engine, checked against a host fp32 loop.  Exposed through the C-ABI as inr_selftest_umma().
//   mode 0: A K-major  (activation image, rows x K)   B K-major  (packed weight stages)   -> fwd / dgrad
//   mode 1: A MN-major (dZ image, K = batch rows)     B MN-major (H image)                -> wgrad
//   mode 2: as mode 1 with N = 16                                                         -> last-layer wgrad / bias
// variant 1 swaps the LBO/SBO descriptor fields (diagnostic for the descriptor convention).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "inr_ptx.cuh"

namespace inr {

struct SelfTestArgs {
  const __half* a;
  const __half* b;
  float* d;
  uint32_t a_bytes, b_bytes;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_kstep, b_kstep;   // byte advance of the descriptor start address per K=16 step
  uint32_t idesc;
  int ksteps, n;
};

__global__ void __launch_bounds__(128, 1) selftest_umma_kernel(SelfTestArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((p.a_bytes + 1023) & ~1023u);
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_load, p.a_bytes + p.b_bytes);
    bulk_g2s(sa, p.a, p.a_bytes, &bar_load);
    bulk_g2s(sb, p.b, p.b_bytes, &bar_load);
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    for (int k = 0; k < p.ksteps; ++k) {
      uint64_t da = umma_smem_desc(smem_u32(sa) + k * p.a_kstep, p.a_lbo, p.a_sbo);
      uint64_t db = umma_smem_desc(smem_u32(sb) + k * p.b_kstep, p.b_lbo, p.b_sbo);
      umma_f16(tmem, da, db, p.idesc, k > 0);
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const int row = warp * 32 + (threadIdx.x & 31);
  for (int c0 = 0; c0 < p.n; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) p.d[static_cast<size_t>(row) * p.n + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace inr

static float lcg(uint32_t& s) {
  s = s * 1664525u + 1013904223u;
  return (static_cast<float>((s >> 8) & 0xFFFF) / 65536.0f) - 0.5f;
}

extern "C" int inr_selftest_umma(int mode, int variant, float* max_abs_err, float* ref_absmax) {
  using namespace inr;
  const int M = 128;
  int N, K;
  if (mode == 0) { N = 256; K = 64; }
  else if (mode == 1) { N = 128; K = 128; }
  else if (mode == 2) { N = 16; K = 128; }
  else return -1;
  std::vector<float> A(static_cast<size_t>(M) * K), B(static_cast<size_t>(N) * K);
  uint32_t seed = 12345u + mode;
  for (auto& x : A) x = __half2float(__float2half(lcg(seed)));
  for (auto& x : B) x = __half2float(__float2half(lcg(seed)));
  // logical A[m][k], B[n][k];  D[m][n] = sum_k A[m][k] B[n][k]
  SelfTestArgs p{};
  std::vector<__half> ia, ib;
  if (mode == 0) {
    // activation image: (k/8)*2048 + m*16 + (k%8)*2 ; weight stages of K=32: s*16384 + (kk/8)*4096 + n*16 + (kk%8)*2
    ia.assign(static_cast<size_t>(M) * K, __float2half(0.f));
    ib.assign(static_cast<size_t>(N) * K, __float2half(0.f));
    for (int m = 0; m < M; ++m)
      for (int k = 0; k < K; ++k) ia[((k / 8) * 2048 + m * 16 + (k % 8) * 2) / 2] = __float2half(A[m * K + k]);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) {
        int s = k / 32, kk = k % 32;
        ib[(s * 16384 + (kk / 8) * 4096 + n * 16 + (kk % 8) * 2) / 2] = __float2half(B[n * K + k]);
      }
    p.a_lbo = 2048; p.a_sbo = 128; p.b_lbo = 4096; p.b_sbo = 128;
    p.a_kstep = 2 * 2048; p.b_kstep = 2 * 4096;   // stage boundary (k=32) lands on 16384 = 4*4096: contiguous
    p.idesc = umma_idesc_f16(M, N, false, false);
  } else {
    // images indexed [feature/8][row][feature%8]: elem(row r, feature f) at (f/8)*2048 + r*16 + (f%8)*2 ; rows are K.
    ia.assign(static_cast<size_t>(M) * K, __float2half(0.f));
    ib.assign(static_cast<size_t>(N) * K, __float2half(0.f));
    for (int m = 0; m < M; ++m)
      for (int k = 0; k < K; ++k) ia[((m / 8) * 2048 + k * 16 + (m % 8) * 2) / 2] = __float2half(A[m * K + k]);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) ib[((n / 8) * 2048 + k * 16 + (n % 8) * 2) / 2] = __float2half(B[n * K + k]);
    p.a_lbo = 128; p.a_sbo = 2048; p.b_lbo = 128; p.b_sbo = 2048;
    p.a_kstep = 256; p.b_kstep = 256;             // 16 rows * 16 B
    p.idesc = umma_idesc_f16(M, N, true, true);
  }
  if (variant == 1) {
    std::swap(p.a_lbo, p.a_sbo);
    std::swap(p.b_lbo, p.b_sbo);
  }
  p.ksteps = K / 16;
  p.n = N;
  p.a_bytes = static_cast<uint32_t>(ia.size() * 2);
  p.b_bytes = static_cast<uint32_t>(ib.size() * 2);
  __half *da = nullptr, *db = nullptr;
  float* dd = nullptr;
  if (cudaMalloc(&da, p.a_bytes) != cudaSuccess) return -2;
  cudaMalloc(&db, p.b_bytes);
  cudaMalloc(&dd, sizeof(float) * M * N);
  cudaMemcpy(da, ia.data(), p.a_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(db, ib.data(), p.b_bytes, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0, sizeof(float) * M * N);
  p.a = da; p.b = db; p.d = dd;
  size_t smem = ((p.a_bytes + 1023) & ~1023u) + p.b_bytes + 1024;
  cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  selftest_umma_kernel<<<1, 128, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> D(static_cast<size_t>(M) * N);
  if (e == cudaSuccess) cudaMemcpy(D.data(), dd, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
  cudaFree(da); cudaFree(db); cudaFree(dd);
  if (e != cudaSuccess) {
    std::fprintf(stderr, "inr_selftest_umma: %s\n", cudaGetErrorString(e));
    return -3;
  }
  float maxerr = 0.f, refmax = 0.f;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc += A[m * K + k] * B[n * K + k];
      maxerr = std::fmax(maxerr, std::fabs(acc - D[static_cast<size_t>(m) * N + n]));
      refmax = std::fmax(refmax, std::fabs(acc));
    }
  if (max_abs_err) *max_abs_err = maxerr;
  if (ref_absmax) *ref_absmax = refmax;
  return 0;
}
