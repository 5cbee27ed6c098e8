// Microbenchmark of the multi-query sweep's inner step with ALL-REGISTER operands (acc = fma(v, q, acc)):
// scalar FFMA vs packed FFMA2 (fma.rn.f32x2), and the L1 step (t = v - q; acc += |t|) scalar vs packed.
// Operands come from shared memory each iteration (as in the kernel), R = 4 rows x NQ = 8 queries x 4 elements per step.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scratch/micro/inner scratch/micro/inner.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 2048, R = 4, NQ = 8;
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float2 f) { return *reinterpret_cast<u64*>(&f); }
__device__ __forceinline__ float2 up(u64 u) { return *reinterpret_cast<float2*>(&u); }

template <int MODE>   // 0 scalar dot, 1 packed dot, 2 scalar L1, 3 packed L1, 4 scalar L2, 5 packed L2
__global__ void __launch_bounds__(256, 2) k(float* out) {
  __shared__ float4 sv[R * 32 * 2], sq[NQ * 32];
  for (int i = threadIdx.x; i < R * 64; i += 256) sv[i] = make_float4(i * 1e-3f, 1.f, 2.f, -i * 1e-3f);
  for (int i = threadIdx.x; i < NQ * 32; i += 256) sq[i] = make_float4(1.f, i * 1e-3f, 0.5f, 2.f);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float acc[NQ * R];
  u64 acc2[NQ * R];
#pragma unroll
  for (int i = 0; i < NQ * R; ++i) { acc[i] = 0.f; acc2[i] = 0; }
  for (int it = 0; it < ITERS; ++it) {
    float4 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = sv[(r * 2 + (it & 1)) * 32 + lane];
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      const float4 q = sq[j * 32 + lane];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (MODE == 0) {
          float a = acc[j * R + r];
          a = fmaf(v[r].x, q.x, a); a = fmaf(v[r].y, q.y, a); a = fmaf(v[r].z, q.z, a); a = fmaf(v[r].w, q.w, a);
          acc[j * R + r] = a;
        } else if (MODE == 1) {
          u64 a = acc2[j * R + r];
          asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(pk(make_float2(v[r].x, v[r].y))), "l"(pk(make_float2(q.x, q.y))));
          asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(pk(make_float2(v[r].z, v[r].w))), "l"(pk(make_float2(q.z, q.w))));
          acc2[j * R + r] = a;
        } else if (MODE == 2) {
          float a = acc[j * R + r];
          a += fabsf(v[r].x - q.x); a += fabsf(v[r].y - q.y); a += fabsf(v[r].z - q.z); a += fabsf(v[r].w - q.w);
          acc[j * R + r] = a;
        } else if (MODE == 3) {
          u64 a = acc2[j * R + r], t0, t1;
          asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(t0) : "l"(pk(make_float2(v[r].x, v[r].y))), "l"(pk(make_float2(q.x, q.y))));
          asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(t1) : "l"(pk(make_float2(v[r].z, v[r].w))), "l"(pk(make_float2(q.z, q.w))));
          t0 &= 0x7fffffff7fffffffull; t1 &= 0x7fffffff7fffffffull;
          asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(t0));
          asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(t1));
          acc2[j * R + r] = a;
        } else if (MODE == 4) {
          float a = acc[j * R + r];
          float d0 = v[r].x - q.x, d1 = v[r].y - q.y, d2 = v[r].z - q.z, d3 = v[r].w - q.w;
          a = fmaf(d0, d0, a); a = fmaf(d1, d1, a); a = fmaf(d2, d2, a); a = fmaf(d3, d3, a);
          acc[j * R + r] = a;
        } else {
          u64 a = acc2[j * R + r], t0, t1;
          asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(t0) : "l"(pk(make_float2(v[r].x, v[r].y))), "l"(pk(make_float2(q.x, q.y))));
          asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(t1) : "l"(pk(make_float2(v[r].z, v[r].w))), "l"(pk(make_float2(q.z, q.w))));
          asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(a) : "l"(t0));
          asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(a) : "l"(t1));
          acc2[j * R + r] = a;
        }
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NQ * R; ++i) { s += acc[i]; float2 t = up(acc2[i]); s += t.x + t.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float* out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 2, 256>>>(out);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) k<MODE><<<148 * 2, 256>>>(out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
  const double elems = 148.0 * 2 * 256 * (double)ITERS * R * NQ * 4;
  printf("%-12s %.3f ms  %.2f T element-steps/s\n", name, ms, elems / ms / 1e9);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 2 * 256 * 4);
  run<0>("dot scalar", out); run<1>("dot packed", out);
  run<2>("L1 scalar", out);  run<3>("L1 packed", out);
  run<4>("L2 scalar", out);  run<5>("L2 packed", out);
  return 0;
}
