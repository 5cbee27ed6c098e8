// Microbenchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) and FADD vs FADD2 on sm_100a.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scratch/micro/ffma2 scratch/micro/ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096, ILP = 8;
__global__ void k_ffma(float* out, float a, float b) {
  float acc[2 * ILP];
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
  unsigned long long acc[ILP];
  float2 av = make_float2(a, a), bv = make_float2(b, b);
  unsigned long long a2 = *reinterpret_cast<unsigned long long*>(&av), b2 = *reinterpret_cast<unsigned long long*>(&bv);
#pragma unroll
  for (int i = 0; i < ILP; ++i) { float2 t = make_float2(threadIdx.x + i, threadIdx.x - i); acc[i] = *reinterpret_cast<unsigned long long*>(&t); }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(a2), "l"(b2));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { float2 t = *reinterpret_cast<float2*>(&acc[i]); s += t.x + t.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fadd2(float* out, float a, float b) {
  unsigned long long acc[ILP];
  float2 av = make_float2(a, a);
  unsigned long long a2 = *reinterpret_cast<unsigned long long*>(&av);
#pragma unroll
  for (int i = 0; i < ILP; ++i) { float2 t = make_float2(threadIdx.x + i, threadIdx.x - i); acc[i] = *reinterpret_cast<unsigned long long*>(&t); }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc[i]) : "l"(a2));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { float2 t = *reinterpret_cast<float2*>(&acc[i]); s += t.x + t.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the manhattan inner step: t = v - q ; acc += |t|   (2 FADD per element)
__global__ void k_l1(float* out, float a, float b) {
  float acc[2 * ILP];
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) acc[i] = threadIdx.x + i;
  float v = a;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) acc[i] += fabsf(v - b * (float)i);
    v += 1.0f;
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename K> float run(K kern, float* out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<148 * 8, 256>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) kern<<<148 * 8, 256>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 10;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const double elems = 148.0 * 8 * 256 * ITERS * 2 * ILP;
  float t;
  t = run(k_ffma, out);  printf("FFMA  : %.3f ms  %.2f T elem-ops/s\n", t, elems / t / 1e9);
  t = run(k_ffma2, out); printf("FFMA2 : %.3f ms  %.2f T elem-ops/s\n", t, elems / t / 1e9);
  t = run(k_fadd2, out); printf("FADD2 : %.3f ms  %.2f T elem-ops/s\n", t, elems / t / 1e9);
  t = run(k_l1, out);    printf("L1(2 FADD/elem): %.3f ms  %.2f T elems/s\n", t, elems / t / 1e9);
  return 0;
}
