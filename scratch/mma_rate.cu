// Issue-to-completion cost of back-to-back tcgen05.mma instructions of the shapes / layouts the PPO gradient kernels use
// (scratch measurement, operand contents irrelevant).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline uint64_t mkdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
struct Cfg { int kind; int M, N; int a_tmem; uint32_t a_lbo, a_sbo, a_layout, a_step; uint32_t b_lbo, b_sbo, b_layout, b_step; int a_mn, b_mn; int count; const char* name; };
__global__ void rate(Cfg c, long long* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* base = sm + ((1024u - (smem_addr(sm) & 1023u)) & 1023u);
  __shared__ unsigned long long bar; __shared__ uint32_t tb;
  const int tid = threadIdx.x;
  for (int i = tid; i < 131072 / 4; i += blockDim.x) ((float*)base)[i] = 1.0f;
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(&bar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_addr(&tb)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t el = 0;
  if (tid < 32) asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(el));
  if (tid < 32 && el) {
    const uint32_t aA = smem_addr(base), aB = smem_addr(base + 65536);
    const uint32_t fmt = c.kind == 0 ? 2u : 1u;      // tf32 : bf16
    const uint32_t id = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    uint32_t parity = 0;
    const uint64_t da0 = mkdesc(aA, c.a_lbo, c.a_sbo, c.a_layout), db0 = mkdesc(aB, c.b_lbo, c.b_sbo, c.b_layout);
    const uint64_t sa = c.a_step >> 4, sb = c.b_step >> 4;
    for (int rep = 0; rep < 4; ++rep) {
      const long long t0 = clock64();
      for (int s8 = 0; s8 < c.count; s8 += 8) {
#pragma unroll
        for (int sl = 0; sl < 8; ++sl) {
          const uint64_t da = da0 + sa * sl, db = db0 + sb * sl;
          const uint32_t acc = (s8 + sl) > 0;
          if (c.a_tmem) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" :: "r"(tb), "r"(tb + 256 + 8 * sl), "l"(db), "r"(id), "r"(acc) : "memory");
          } else if (c.kind == 0) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" :: "r"(tb), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
          } else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" :: "r"(tb), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
          }
        }
      }
      const long long t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_addr(&bar)) : "memory");
      { uint32_t done; do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_addr(&bar)), "r"(parity) : "memory"); } while (!done); }
      parity ^= 1;
      const long long t2 = clock64();
      out[2 * rep] = t1 - t0; out[2 * rep + 1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(512u) : "memory");
}
int main() {
  const uint32_t G = 2048;   // bf16 interleave group stride
  Cfg cfgs[] = {
    {0, 128, 64, 1, 0, 0, 0, 0,   128, 2048, 0, 256,  0, 0, 64, "tf32 TS 128x64x8   (fwd / dgrad)"},
    {0, 128, 16, 1, 0, 0, 0, 0,   128, 2048, 0, 256,  0, 0, 64, "tf32 TS 128x16x8   (head)"},
    {0, 64, 64, 0, 16384, 512, 1, 1024,  16384, 512, 1, 1024,  1, 1, 64, "tf32 SS 64x64x8  MN sw128_32B (v2 dW2)"},
    {0, 64, 8, 0, 16384, 512, 1, 1024,  144, 4608, 0, 288,  1, 0, 64, "tf32 SS 64x8x8   (v2 dW3 / db2)"},
    {0, 64, 16, 0, 16384, 512, 1, 1024,  144, 4608, 0, 288,  1, 0, 64, "tf32 SS 64x16x8  (v2 dW1)"},
    {1, 64, 72, 0, 128, G, 0, 256,   128, G, 0, 256,   1, 1, 64, "bf16 SS 64x72x16 MN interleave (v3 dW2|db2)"},
    {1, 64, 64, 0, 128, G, 0, 256,   128, G, 0, 256,   1, 1, 64, "bf16 SS 64x64x16 MN interleave"},
    {1, 64, 64, 0, 16384, 1024, 2, 2048,  16384, 1024, 2, 2048,  1, 1, 64, "bf16 SS 64x64x16 MN sw128"},
    {1, 64, 8, 0, 128, G, 0, 256,    128, G, 0, 256,   1, 1, 64, "bf16 SS 64x8x16  MN interleave (v3 dW3)"},
    {1, 64, 16, 0, 128, G, 0, 256,   128, G, 0, 256,   1, 1, 64, "bf16 SS 64x16x16 MN interleave (v3 dW1)"},
    {1, 64, 8, 0, 16384, 1024, 2, 2048,  128, G, 0, 256,  1, 1, 64, "bf16 SS 64x8x16  A sw128, B interleave"},
    {1, 128, 64, 0, 128, 1024, 0, 256,  128, 1024, 0, 256,  0, 0, 64, "bf16 SS 128x64x16 K-major nosw"},
    {0, 128, 64, 1, 0, 0, 0, 0,   128, 2048, 0, 256,  0, 0, 8, "tf32 TS 128x64x8   x8 only"},
    {0, 128, 64, 1, 0, 0, 0, 0,   128, 2048, 0, 256,  0, 0, 2, "tf32 TS 128x64x8   x2 only"},
  };
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
  long long* d; cudaMalloc(&d, 64);
  for (Cfg c : cfgs) {
    rate<<<1, 128, 131072 + 1024>>>(c, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("%-48s n=%2d  issue %6lld  complete %6lld cycles  (%5.1f / MMA)   %s\n", c.name, c.count, h[6], h[7], (double)h[7] / c.count, cudaGetErrorString(e));
    if (e != cudaSuccess) break;
  }
  return 0;
}
