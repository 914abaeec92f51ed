// Microbenchmark: issue cost of the packed fp32 instructions of sm_100a (FADD2 / FMUL2 / FFMA2) against their
// scalar forms, with the occupancy the MLP kernels run at (512 threads per SM = 4 warps per scheduler).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float2 upk(uint64_t r) {
  float2 c;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(r));
  return c;
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, float a, float b, int iters, long long* cycles) {
  constexpr int NV = 16;            // independent results per thread per iteration
  float v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = threadIdx.x * 0.001f + i;
  uint64_t p[NV / 2];
#pragma unroll
  for (int i = 0; i < NV / 2; ++i) p[i] = pk(v[2 * i], v[2 * i + 1]);
  const uint64_t pa = pk(a, a), pb = pk(b, b);
  uint32_t sel = threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {                 // scalar FFMA
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = fmaf(v[i], a, b);
    } else if (MODE == 1) {          // FFMA2
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
    } else if (MODE == 2) {          // scalar FADD
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = v[i] + a;
    } else if (MODE == 3) {          // FADD2
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
    } else if (MODE == 4) {          // scalar FFMA + one ALU op (LOP3) per result: the two pipes side by side
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        v[i] = fmaf(v[i], a, b);
        sel = (sel ^ __float_as_uint(v[i])) & 0x7fffffffu;
      }
    } else if (MODE == 5) {          // FFMA2 + the same ALU work
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) {
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
        const float2 c = upk(p[i]);
        sel = (sel ^ __float_as_uint(c.x)) & 0x7fffffffu;
        sel = (sel ^ __float_as_uint(c.y)) & 0x7fffffffu;
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += v[i];
#pragma unroll
  for (int i = 0; i < NV / 2; ++i) s += upk(p[i]).x + upk(p[i]).y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(sel);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
static void run(const char* name, float* out, long long* cyc) {
  const int iters = 4096;
  k<MODE><<<148, 512>>>(out, 1.0001f, 0.5f, iters, cyc);
  k<MODE><<<148, 512>>>(out, 1.0001f, 0.5f, iters, cyc);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  // 4 warps per scheduler, 16 results per thread per iteration
  printf("%-34s %8.3f cycles per warp-result per scheduler (%.3f results/clk/SMSP)\n", name, (double)c / (iters * 16.0 * 4.0),
         iters * 16.0 * 4.0 / (double)c);
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  run<0>("FFMA", out, cyc);
  run<1>("FFMA2", out, cyc);
  run<2>("FADD", out, cyc);
  run<3>("FADD2", out, cyc);
  run<4>("FFMA + LOP3 per result", out, cyc);
  run<5>("FFMA2 + LOP3 per result", out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
