// FP32 pipe micro-benchmark for sm_100a: issue rate of scalar FFMA vs packed FFMA2/FMUL2/FADD2 (3-register forms),
// to place the dynamics kernel on the REAL FP32 roofline of the part.   nvcc -arch=sm_100a -O3 -o fma_mb fma_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, int iters, float a, float b) {
  // 8 independent accumulator chains per thread
  if (MODE == 6 || MODE == 7 || MODE == 8) {  // scalar ops with distinct rotating operands (no reuse-cache help)
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i + a;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 6) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(x[i]) : "f"(x[(i + 1) & 7]), "f"(x[(i + 2) & 7]));
        if (MODE == 7) asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(x[i]) : "f"(x[(i + 1) & 7]), "f"(x[(i + 2) & 7]));
        if (MODE == 8) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x[i]) : "f"(x[(i + 1) & 7]), "f"(x[(i + 2) & 7]), "f"(x[(i + 3) & 7]));
      }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else if (MODE == 9) {  // hybrid: one packed 2-operand FMUL2 + two scalar 3-operand FFMA per group
    unsigned long long x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(threadIdx.x * 0.001f + i), "f"(threadIdx.x * 0.002f + i + a));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float l0, h0, l1, h1, l2, h2;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(l0), "=f"(h0) : "l"(x[i]));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(l1), "=f"(h1) : "l"(x[(i + 1) & 7]));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(l2), "=f"(h2) : "l"(x[(i + 2) & 7]));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(l0) : "f"(l1), "f"(l2));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(h0) : "f"(h1), "f"(h2));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(l0), "f"(h0));
      }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
  } else if (MODE == 0) {  // scalar FFMA, 3 register operands
    float x[8], y = a, z = b;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(y), "f"(z));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {
    unsigned long long x[8], y, z;
    asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(a), "f"(a + 1.f));
    asm("mov.b64 %0, {%1, %2};" : "=l"(z) : "f"(b), "f"(b + 1.f));
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(threadIdx.x * 0.001f + i), "f"(threadIdx.x * 0.002f + i));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(y), "l"(z));
        if (MODE == 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(y));
        if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(y));
        if (MODE == 4) {  // FFMA2 interleaved with an ALU op (LOP3) per packed op
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(y), "l"(z));
          asm volatile("xor.b64 %0, %0, 0x8000000080000000;" : "+l"(x[i]));
        }
        if (MODE == 5) {  // fma with chained distinct operands x[i], x[(i+1)%8], x[(i+2)%8]
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(x[i]) : "l"(x[(i + 1) & 7]), "l"(x[(i + 2) & 7]));
        }
      }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
  }
}

template <int MODE> void run(const char* name, int ops_per_inst, float* d_out, int sms, int clock_khz) {
  const int iters = 4096, blocks = sms * 8, threads = 256;
  bench<MODE><<<blocks, threads>>>(d_out, 16, 1.0001f, 0.5f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    bench<MODE><<<blocks, threads>>>(d_out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double insts = (double)blocks * threads / 32 * iters * 8 * (MODE == 4 ? 1 : 1);  // warp-instructions of the op under test
  const double lane_ops = insts * 32 * ops_per_inst;
  const double cycles_at_max = best * 1e-3 * clock_khz * 1e3;
  printf("%-38s %8.3f ms  %7.2f Tlane-op/s  %6.1f lane-ops/clk/SM (at max clock %d MHz)  %5.2f warp-inst/clk/SMSP\n", name, best,
         lane_ops / (best * 1e-3) / 1e12, lane_ops / cycles_at_max / sms, clock_khz / 1000, insts / cycles_at_max / sms / 4);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%s: %d SMs, max clock %d MHz\n", p.name, p.multiProcessorCount, clk / 1000);
  float* d; cudaMalloc(&d, sizeof(float) * p.multiProcessorCount * 8 * 256);
  run<0>("FFMA   (scalar, 3 reg operands)", 1, d, p.multiProcessorCount, clk);
  run<1>("FFMA2  (packed, 3 reg-pair operands)", 2, d, p.multiProcessorCount, clk);
  run<2>("FMUL2  (packed, 2 reg-pair operands)", 2, d, p.multiProcessorCount, clk);
  run<3>("FADD2  (packed, 2 reg-pair operands)", 2, d, p.multiProcessorCount, clk);
  run<4>("FFMA2 + LOP3x2 interleaved", 2, d, p.multiProcessorCount, clk);
  run<5>("FFMA2  (3 distinct rotating operands)", 2, d, p.multiProcessorCount, clk);
  run<6>("FFMA   (scalar, 3 distinct rotating, acc)", 1, d, p.multiProcessorCount, clk);
  run<7>("FMUL   (scalar, 2 distinct rotating)", 1, d, p.multiProcessorCount, clk);
  run<8>("FFMA   (scalar, 3 distinct rotating, new dst)", 1, d, p.multiProcessorCount, clk);
  run<9>("FFMA x2 on halves of rotating pairs (2/grp)", 2, d, p.multiProcessorCount, clk);
  return 0;
}
