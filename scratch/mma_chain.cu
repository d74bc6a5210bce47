// Is the per-instruction cost that scratch/mma_rate.cu measured (45 cycles TS, 80-95 cycles SS for the small shapes of the PPO
// gradient kernel) a THROUGHPUT or the latency of a dependent accumulate chain?  mma_rate.cu accumulates all 64 MMAs into ONE
// accumulator.  Here the same MMAs are spread round-robin over `chains` independent accumulators (different TMEM columns),
// and a mixed stream alternates the TS dgrad MMA and the SS wgrad MMA of step S5 (scratch measurement; operand contents
// irrelevant).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/mma_chain scratch/mma_chain.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline uint64_t mkdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
// mode 0: tf32 TS 128xNx8 (A in TMEM, B K-major no-swizzle) ; 1: bf16 SS MxNx16 MN-major interleave ; 2: alternate mode 0 (N=64) and mode 1
struct Cfg { int mode, M, N, chains, count; const char* name; };
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" :: "r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss16(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" :: "r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
template <int MODE, int CHAINS>
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
    const uint32_t id_ts = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((MODE == 2 ? 64 : c.N) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t id_ss = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)((MODE == 2 ? 72 : c.N) >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    const uint64_t dbw = mkdesc(aB, 128, 2048, 0);                       // K-major weights (B of the TS MMAs)
    const uint64_t dai = mkdesc(aA, 128, 2048, 0), dbi = mkdesc(aB, 128, 2048, 0);   // MN-major interleave operands
    uint32_t parity = 0;
    for (int rep = 0; rep < 4; ++rep) {
      const long long t0 = clock64();
      for (int s8 = 0; s8 < c.count; s8 += 8) {     // count: a multiple of 8 (and of CHAINS where CHAINS == 3: 8 slices x 3 per round)
#pragma unroll
        for (int sl = 0; sl < 8; ++sl) {
          const uint32_t acc = s8 > 0;                 // the first round of 8 initialises every accumulator it touches
          if (MODE == 0) mma_ts(tb + 128 * (sl % CHAINS), tb + 384 + 8 * sl, dbw + 16 * sl, id_ts, acc | (sl >= CHAINS));
          else if (MODE == 1) mma_ss16(tb + 128 * (sl % CHAINS), dai + 16 * sl, dbi + 16 * sl, id_ss, acc | (sl >= CHAINS));
          else { mma_ts(tb, tb + 384 + 8 * sl, dbw + 16 * sl, id_ts, acc | (sl > 0)); mma_ss16(tb + 128, dai + 16 * sl, dbi + 16 * sl, id_ss, acc | (sl > 0)); }
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
  Cfg cfgs[] = {
    {0, 128, 64, 1, 64, "tf32 TS 128x64x8, 1 accumulator"}, {0, 128, 64, 2, 64, "tf32 TS 128x64x8, 2 accumulators"}, {0, 128, 64, 3, 64, "tf32 TS 128x64x8, 3 accumulators"},
    {0, 128, 16, 1, 64, "tf32 TS 128x16x8, 1 accumulator"}, {0, 128, 16, 2, 64, "tf32 TS 128x16x8, 2 accumulators"},
    {0, 128, 128, 1, 64, "tf32 TS 128x128x8, 1 accumulator"}, {0, 128, 128, 2, 64, "tf32 TS 128x128x8, 2 accumulators"},
    {1, 64, 72, 1, 64, "bf16 SS 64x72x16, 1 accumulator"}, {1, 64, 72, 2, 64, "bf16 SS 64x72x16, 2 accumulators"}, {1, 64, 72, 3, 64, "bf16 SS 64x72x16, 3 accumulators"},
    {1, 64, 16, 1, 64, "bf16 SS 64x16x16, 1 accumulator"}, {1, 64, 16, 2, 64, "bf16 SS 64x16x16, 2 accumulators"},
    {1, 64, 8, 1, 64, "bf16 SS 64x8x16, 1 accumulator"}, {1, 64, 8, 2, 64, "bf16 SS 64x8x16, 2 accumulators"},
    {1, 128, 96, 1, 64, "bf16 SS 128x96x16, 1 accumulator (stacked dZ2|dZ1)"}, {1, 128, 96, 2, 64, "bf16 SS 128x96x16, 2 accumulators"},
    {1, 128, 80, 1, 64, "bf16 SS 128x80x16, 1 accumulator"}, {1, 128, 128, 1, 64, "bf16 SS 128x128x16, 1 accumulator"},
    {2, 64, 72, 1, 64, "S5 mix: TS 128x64x8 -> P alternating with SS 64x72x16 -> acc (64 of each)"},
  };
  long long* d; cudaMalloc(&d, 64);
  for (Cfg c : cfgs) {
    auto k = c.mode == 2 ? rate<2, 1> : c.mode == 0 ? (c.chains == 1 ? rate<0, 1> : c.chains == 2 ? rate<0, 2> : rate<0, 3>)
                                                      : (c.chains == 1 ? rate<1, 1> : c.chains == 2 ? rate<1, 2> : rate<1, 3>);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
    k<<<1, 128, 131072 + 1024>>>(c, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    const int n = c.mode == 2 ? 2 * c.count : c.count;
    printf("%-78s n=%3d  issue %6lld  complete %6lld cycles  (%5.1f / MMA)   %s\n", c.name, n, h[6], h[7], (double)h[7] / n, cudaGetErrorString(e));
    if (e != cudaSuccess) break;
  }
  return 0;
}
