// Probe of tcgen05.mma tf32 operand layouts (not product code): D[M x N] = A[M x K] . B[N x K]^T, small integers.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../drone_rl_b200/csrc/ppo_update_tc.cuh"
using namespace dronecu;
using namespace dronecu::tcu;

struct Cfg { int M, N, K; int a_mode, b_mode; };   // mode: 0 = K-major plain, 1 = K-major SW128, 2 = MN-major SW128, 3 = MN-major plain
// element (mn, k) byte offset in the operand buffer for each mode.  MNdim = number of MN rows, K = total K
__host__ __device__ int off_bytes(int mode, int mn, int k, int MN, int K) {
  switch (mode) {
    case 0: return (mn >> 3) * (K * 32) + (k >> 2) * 128 + (mn & 7) * 16 + (k & 3) * 4;          // LBO 128, SBO K*32
    case 1: { int half = k >> 5, kk = k & 31; int c = kk >> 2; return half * (MN * 128) + mn * 128 + ((c ^ (mn & 7)) << 4) + (kk & 3) * 4; }
    case 2: { int half = mn >> 5, m2 = mn & 31; int c = m2 >> 2; return half * (K * 128) + k * 128 + ((c ^ (k & 7)) << 4) + (m2 & 3) * 4; }
    case 4: { int half = mn >> 5, m2 = mn & 31; int q = m2 >> 3; return half * (K * 128) + k * 128 + ((q ^ (k & 3)) << 5) + (m2 & 7) * 4; }
    case 5: return (mn >> 3) * ((K >> 2) * 144) + (k >> 2) * 144 + (mn & 7) * 16 + (k & 3) * 4;
    case 3: return (mn & 3) * 4 + (k & 7) * 16 + (mn >> 2) * 128 + (k >> 3) * (MN * 32);           // SBO 128 (MN chunks), LBO MN*32 (K groups)
  }
  return 0;
}
__global__ void probe(Cfg c, const float* A, const float* B, float* D) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* base = sm + ((1024u - (smem_addr(sm) & 1023u)) & 1023u);
  unsigned char* bufA = base; unsigned char* bufB = base + 65536;
  __shared__ unsigned long long bar; __shared__ uint32_t tb;
  const int tid = threadIdx.x;
  for (int i = tid; i < 65536 / 4; i += 128) { ((float*)bufA)[i] = 0.f; ((float*)bufB)[i] = 0.f; }
  __syncthreads();
  for (int i = tid; i < c.M * c.K; i += 128) { int m = i / c.K, k = i % c.K; *(float*)(bufA + off_bytes(c.a_mode, m, k, c.M, c.K)) = A[i]; }
  for (int i = tid; i < c.N * c.K; i += 128) { int n = i / c.K, k = i % c.K; *(float*)(bufB + off_bytes(c.b_mode, n, k, c.N, c.K)) = B[i]; }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_addr(&tb)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  proxy_fence(); fence_before(); __syncthreads(); fence_after();
  // zero D
  { float z[16]; for (int i = 0; i < 16; ++i) z[i] = -777.f;
    for (int cc = 0; cc < 16; ++cc) tc::tmem_st16(tb + ((uint32_t)((tid >> 5) * 32) << 16) + 16 * cc, z);
    tc::wait_st(); }
  fence_before(); __syncthreads();
  if (tid == 0) {
    fence_after();
    const uint32_t aA = smem_addr(bufA), aB = smem_addr(bufB);
    const uint32_t id = idesc(c.M, c.N, (c.a_mode >= 2 && c.a_mode != 5), (c.b_mode >= 2 && c.b_mode != 5));
    for (int s = 0; s < c.K / 8; ++s) {
      uint64_t da, db;
      auto mk = [&](int mode, uint32_t a, int MN) -> uint64_t {
        switch (mode) {
          case 0: return desc_plain(a + 256 * s, 128, c.K * 32);
          case 1: return desc_sw128(a + (s >> 2) * (MN * 128) + (s & 3) * 32, 16, 1024);
          case 2: return desc_sw128(a + 1024 * s, c.K * 128, 1024);
          case 5: return desc_plain(a + 288 * s, 144, (c.K >> 2) * 144);
          case 4: return (desc_plain(a + 1024 * s, c.K * 128, 512) | ((uint64_t)1 << 61));
          default: return desc_plain(a + (MN * 32) * s, MN * 32, 128);
        }
      };
      da = mk(c.a_mode, aA, c.M); db = mk(c.b_mode, aB, c.N);
      mma_ss(tb, da, db, id, s > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0); fence_after();
  for (int cc = 0; cc < 16; ++cc) {
    float v[16]; tmem_ld16(tb + ((uint32_t)((tid >> 5) * 32) << 16) + 16 * cc, v);
    for (int i = 0; i < 16; ++i) D[tid * 256 + 16 * cc + i] = v[i];
  }
  fence_before(); __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(256u) : "memory");
}
int main() {
  Cfg cfgs[] = {{64, 16, 128, 4, 5}, {64, 8, 128, 4, 5}, {64, 24, 128, 4, 5}, {64, 64, 128, 4, 0}, {64, 16, 128, 4, 0}, {64, 8, 128, 4, 0}, {64, 64, 128, 4, 4}, {128, 64, 64, 4, 0}, {64, 64, 8, 4, 0}, {64, 64, 128, 0, 4}};
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 * 2 + 1024);
  for (Cfg c : cfgs) {
    std::vector<float> A(c.M * c.K), B(c.N * c.K), D(128 * 256);
    for (int i = 0; i < c.M * c.K; ++i) A[i] = (float)((i * 7 + i / c.K) % 5 - 2);
    for (int i = 0; i < c.N * c.K; ++i) B[i] = (float)((i * 3 + i / c.K * 2) % 7 - 3);
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    probe<<<1, 128, 65536 * 2 + 1024>>>(c, dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, zero = 0, untouched = 0; double maxerr = 0;
    for (int m = 0; m < c.M; ++m) {
      const int lane = (c.M == 128) ? m : (m % 16 + 32 * (m / 16));
      for (int n = 0; n < c.N; ++n) {
        double ref = 0; for (int k = 0; k < c.K; ++k) ref += (double)A[m * c.K + k] * B[n * c.K + k];
        const float got = D[lane * 256 + n];
        if (got == -777.f) ++untouched;
        if (got == 0.f && ref != 0) ++zero;
        if (std::fabs(got - ref) > 1e-3) ++bad;
        maxerr = std::fmax(maxerr, std::fabs(got - ref));
      }
    }
    printf("M%d N%d K%d a_mode %d b_mode %d: %s  bad %d / %d  zero %d untouched %d maxerr %g  D[0][0..3] = %g %g %g %g\n", c.M, c.N, c.K, c.a_mode, c.b_mode,
           cudaGetErrorString(e), bad, c.M * c.N, zero, untouched, maxerr, D[0], D[1], D[2], D[3]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  return 0;
}
