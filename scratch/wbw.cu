// HBM write-only / copy bandwidth probe (scratch; evidence for the write-stream ceiling in DESIGN.md)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void wr(float4* __restrict__ p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
  float4 v = make_float4(i, 1.f, 2.f, 3.f);
  for (; i < n; i += s) p[i] = v;
}
__global__ void wr_cs(float4* __restrict__ p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
  float4 v = make_float4(i, 1.f, 2.f, 3.f);
  for (; i < n; i += s) __stcs(p + i, v);
}
// each thread writes 60 B rows (15 floats) like the obs record: contiguous per warp
__global__ void wr_contig(float4* __restrict__ p, size_t n) {   // one-shot grid, 1 quad per thread x 4 unrolled
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
  float4 v = make_float4(i, 1.f, 2.f, 3.f);
  size_t base = (size_t)blockIdx.x * blockDim.x * 4 + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) if (base + k * blockDim.x < n) p[base + k * blockDim.x] = v;
}
__global__ void cp(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += s) b[i] = a[i];
}
// the record pattern of rollout_kernel without any compute: per warp and step 1920 B of obs (128-bit stores), 512 B of
// actions, 128 B of rewards, 32 B of done flags, into four step-major [K, n, ...] arrays
__global__ void __launch_bounds__(256, 3) wr_record(float* obs, float4* act, float* rew, unsigned char* done, size_t n, int K) {
  const unsigned lane = threadIdx.x & 31;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, wb = i - lane;
  if (i >= n) return;
  const float4 v = make_float4((float)i, 1.f, 2.f, 3.f);
  size_t koff = 0;
  for (int k = 0; k < K; ++k, koff += n) {
    float4* g4 = reinterpret_cast<float4*>(obs + (koff + wb) * 15);
#pragma unroll
    for (int j = lane; j < 120; j += 32) __stcs(g4 + j, v);
    __stcs(act + koff + i, v);
    rew[koff + i] = v.x;
    done[koff + i] = (unsigned char)k;
  }
}

int main() {
  size_t bytes = 8ull << 30, n = bytes / 16;
  float4 *a, *b; cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
  cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto f, double moved) {
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(e0); for (int i = 0; i < 10; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s %8.1f GB/s  err=%s\n", name, moved * 10 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  for (int mult : {8, 16, 32}) {
    char nm[64];
    snprintf(nm, 64, "write grid=148x%d x256", mult);
    run(nm, [&] { wr<<<148 * mult, 256>>>(a, n); }, (double)bytes);
    snprintf(nm, 64, "write.cs grid=148x%d x256", mult);
    run(nm, [&] { wr_cs<<<148 * mult, 256>>>(a, n); }, (double)bytes);
  }
  run("write one-shot 4/thread", [&] { wr_contig<<<(unsigned)((n + 1023) / 1024), 256>>>(a, n); }, (double)bytes);
  {
    const size_t n = 8388608; const int K = 32;
    float* obs; float4* act; float* rew; unsigned char* done;
    cudaMalloc(&obs, n * K * 60); cudaMalloc(&act, n * K * 16); cudaMalloc(&rew, n * K * 4); cudaMalloc(&done, n * K);
    run("record pattern 8.4M x 32 (81 B/step)", [&] { wr_record<<<(unsigned)(n / 256), 256>>>(obs, act, rew, done, n, K); }, (double)n * K * 81);
    cudaFree(obs); cudaFree(act); cudaFree(rew); cudaFree(done);
  }
  run("cudaMemsetAsync", [&] { cudaMemsetAsync(a, 1, bytes); }, (double)bytes);
  run("copy kernel (r+w)", [&] { cp<<<148 * 16, 256>>>(a, b, n); }, 2.0 * bytes);
  run("cudaMemcpyAsync d2d (r+w)", [&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, 2.0 * bytes);
  return 0;
}
