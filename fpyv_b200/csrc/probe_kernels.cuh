// FP32-pipe probe (diagnostic entry fpv_probe_fp32): the measured FP32 peak that bench.py reports next to the nominal
// SMs x 128 lanes x 2 x clock.  16 independent accumulator chains per thread, multiplicand and addend shared by all of
// them (operand-reuse cache hits: the register file is not the limit here, the FMA pipe is).
#pragma once
#include <cuda_runtime.h>

namespace fpv {

template <int PACKED>
__global__ void __launch_bounds__(256) fp32_probe_kernel(float* sink, int iters, float a, float b) {
  if (PACKED) {
    unsigned long long x[16], y, z;
    asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(a), "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(z) : "f"(b), "f"(b));
#pragma unroll
    for (int i = 0; i < 16; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(threadIdx.x * 1e-3f + i), "f"(threadIdx.x * 2e-3f + i));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(y), "l"(z));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s ^= x[i];
    if (s == 0x123456789abcdefull) sink[blockIdx.x * blockDim.x + threadIdx.x] = 1.f;   // never true: keeps the chains alive
  } else {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123456.789f) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}

}  // namespace fpv
