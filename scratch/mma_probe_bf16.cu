// Probe of tcgen05.mma kind::f16 (bf16 operands, fp32 accumulate) shared-memory operand layouts (not product code):
// D[M x N] = A[M x K] . B[N x K]^T with small integers.  Modes: 0 = K-major no-swizzle, 2 = MN-major SWIZZLE_128B,
// 3 = MN-major no-swizzle (interleave).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline uint64_t mkdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
struct Cfg { int M, N, K; int a_mode, b_mode; int sw_layout; };
// byte offset of element (mn, k); MN = rows of this operand, K total
__host__ __device__ int off_bytes(int mode, int mn, int k, int MN, int K) {
  switch (mode) {
    case 0: return (mn >> 3) * (K * 16) + (k >> 3) * 128 + (mn & 7) * 16 + (k & 7) * 2;                 // K-major: LBO 128 (K chunks of 8), SBO K*16
    case 2: { int atom = mn >> 6, m2 = mn & 63; int c = m2 >> 3; return atom * (K * 128) + k * 128 + ((c ^ (k & 7)) << 4) + (m2 & 7) * 2; }  // MN SW128
    case 3: return (mn >> 3) * ((K >> 3) * 128) + (k >> 3) * 128 + (k & 7) * 16 + (mn & 7) * 2;           // MN interleave: LBO 128 (K groups), SBO K*16
  }
  return 0;
}
__global__ void probe(Cfg c, const float* A, const float* B, float* D) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* base = sm + ((1024u - (smem_addr(sm) & 1023u)) & 1023u);
  unsigned char* bufA = base; unsigned char* bufB = base + 32768;
  __shared__ unsigned long long bar; __shared__ uint32_t tb;
  const int tid = threadIdx.x;
  for (int i = tid; i < 65536 / 4; i += 128) ((float*)base)[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < c.M * c.K; i += 128) { int m = i / c.K, k = i % c.K; *(__nv_bfloat16*)(bufA + off_bytes(c.a_mode, m, k, c.M, c.K)) = __float2bfloat16(A[i]); }
  for (int i = tid; i < c.N * c.K; i += 128) { int n = i / c.K, k = i % c.K; *(__nv_bfloat16*)(bufB + off_bytes(c.b_mode, n, k, c.N, c.K)) = __float2bfloat16(B[i]); }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(&bar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_addr(&tb)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tL = tb + ((uint32_t)((tid >> 5) * 32) << 16);
  for (int cc = 0; cc < 16; ++cc) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" :: "r"(tL + 16 * cc), "f"(-777.f) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t aA = smem_addr(bufA), aB = smem_addr(bufB);
    const uint32_t id = idesc_bf16(c.M, c.N, c.a_mode >= 2, c.b_mode >= 2);
    for (int s = 0; s < c.K / 16; ++s) {           // K = 16 per instruction
      auto mk = [&](int mode, uint32_t a, int MN) -> uint64_t {
        switch (mode) {
          case 0: return mkdesc(a + 256 * s, 128, c.K * 16, 0);
          case 2: return mkdesc(a + 2048 * s, c.K * 128, 1024, c.sw_layout);
          default: return mkdesc(a + 256 * s, 128, c.K * 16, 0);
        }
      };
      const uint64_t da = mk(c.a_mode, aA, c.M), db = mk(c.b_mode, aB, c.N);
      const uint32_t acc = s > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                   :: "r"(tb), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_addr(&bar)) : "memory");
  }
  { uint32_t done; do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_addr(&bar)), "r"(0u) : "memory"); } while (!done); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int cc = 0; cc < 16; ++cc) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tL + 16 * cc) : "memory");
    for (int i = 0; i < 16; ++i) D[tid * 256 + 16 * cc + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(256u) : "memory");
}
int main() {
  // sw_layout candidates for MN-major 128B swizzle: 2 (SWIZZLE_128B), 1 (128B with 32B base), 6?, 4?
  Cfg cfgs[] = {{64, 72, 128, 3, 3, 2}, {64, 16, 128, 3, 3, 2}, {64, 8, 128, 3, 3, 2}, {64, 64, 128, 2, 2, 2}, {64, 64, 128, 2, 2, 1}, {64, 16, 128, 2, 3, 2}, {64, 8, 128, 2, 3, 2}, {64, 16, 128, 2, 0, 2},
                {64, 8, 128, 2, 0, 2}, {64, 24, 128, 2, 3, 2}, {64, 64, 128, 0, 0, 2}, {128, 64, 64, 0, 0, 2}};
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  for (Cfg c : cfgs) {
    std::vector<float> A(c.M * c.K), B(c.N * c.K), D(128 * 256);
    for (int i = 0; i < c.M * c.K; ++i) A[i] = (float)((i * 7 + i / c.K) % 5 - 2);
    for (int i = 0; i < c.N * c.K; ++i) B[i] = (float)((i * 3 + i / c.K * 2) % 7 - 3);
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    probe<<<1, 128, 65536 + 1024>>>(c, dA, dB, dD);
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
    printf("M%d N%d K%d a_mode %d b_mode %d sw %d: %s  bad %d / %d  zero %d untouched %d maxerr %g  D[0][0..3] = %g %g %g %g\n", c.M, c.N, c.K, c.a_mode, c.b_mode, c.sw_layout,
           cudaGetErrorString(e), bad, c.M * c.N, zero, untouched, maxerr, D[0], D[1], D[2], D[3]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    if (e != cudaSuccess) break;
  }
  return 0;
}
