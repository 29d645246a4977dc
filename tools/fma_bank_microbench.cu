// Register-operand bandwidth probe for sm_100a: is the ~0.66 warp-inst/clk/SMSP of a 3-distinct-operand
// FFMA / the ~0.33 of FFMA2 a property of the register file's read bandwidth, or of bank conflicts in one particular
// register allocation?  x[0..N) live in N registers; instruction i computes x[i] += x[(i+A)%N] * x[(i+B)%N].  Different
// (A, B) give different bank patterns in the SASS that ptxas emits; tools/sass_bank_model.py parses THIS binary's SASS,
// predicts the issue rate of every variant under each register-file hypothesis, and the measured rates printed here pick
// the hypothesis.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_bank tools/fma_bank_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int N, int A, int B>
__global__ void __launch_bounds__(256) bank_scalar(float* io, int iters) {
  float x[N];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = io[threadIdx.x * N + i];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(x[i]) : "f"(x[(i + A) % N]), "f"(x[(i + B) % N]));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) io[threadIdx.x * N + i] = x[i];
}

template <int N, int A, int B>
__global__ void __launch_bounds__(256) bank_packed(unsigned long long* io, int iters) {
  unsigned long long x[N];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = io[threadIdx.x * N + i];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(x[i]) : "l"(x[(i + A) % N]), "l"(x[(i + B) % N]));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) io[threadIdx.x * N + i] = x[i];
}

// two-operand forms for the same register sets: x[i] = x[(i+A)%N] * x[(i+B)%N]
template <int N, int A, int B>
__global__ void __launch_bounds__(256) bank_mul(float* io, int iters) {
  float x[N];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = io[threadIdx.x * N + i];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(x[i]) : "f"(x[(i + A) % N]), "f"(x[(i + B) % N]));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) io[threadIdx.x * N + i] = x[i];
}

#define RUN_S(N, A, B)                                                                                                   \
  {                                                                                                                      \
    bank_scalar<N, A, B><<<blocks, 256>>>((float*)d, 4);                                                                 \
    float best = 1e9f;                                                                                                   \
    for (int r = 0; r < 5; ++r) {                                                                                        \
      cudaEventRecord(e0); bank_scalar<N, A, B><<<blocks, 256>>>((float*)d, iters); cudaEventRecord(e1);                 \
      cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;                   \
    }                                                                                                                    \
    printf("scalar N=%d A=%d B=%d  %.4f ms  ipc %.3f\n", N, A, B, best, (double)blocks * 8 * iters * N / (best * 1e-3 * clk * 1e3) / sms / 4); \
  }
#define RUN_P(N, A, B)                                                                                                   \
  {                                                                                                                      \
    bank_packed<N, A, B><<<blocks, 256>>>((unsigned long long*)d, 4);                                                    \
    float best = 1e9f;                                                                                                   \
    for (int r = 0; r < 5; ++r) {                                                                                        \
      cudaEventRecord(e0); bank_packed<N, A, B><<<blocks, 256>>>((unsigned long long*)d, iters); cudaEventRecord(e1);    \
      cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;                   \
    }                                                                                                                    \
    printf("packed N=%d A=%d B=%d  %.4f ms  ipc %.3f\n", N, A, B, best, (double)blocks * 8 * iters * N / (best * 1e-3 * clk * 1e3) / sms / 4); \
  }
#define RUN_M(N, A, B)                                                                                                   \
  {                                                                                                                      \
    bank_mul<N, A, B><<<blocks, 256>>>((float*)d, 4);                                                                    \
    float best = 1e9f;                                                                                                   \
    for (int r = 0; r < 5; ++r) {                                                                                        \
      cudaEventRecord(e0); bank_mul<N, A, B><<<blocks, 256>>>((float*)d, iters); cudaEventRecord(e1);                    \
      cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;                   \
    }                                                                                                                    \
    printf("mul    N=%d A=%d B=%d  %.4f ms  ipc %.3f\n", N, A, B, best, (double)blocks * 8 * iters * N / (best * 1e-3 * clk * 1e3) / sms / 4); \
  }

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int sms = p.multiProcessorCount, iters = 4096, blocks = sms * 8;
  printf("%s: %d SMs, max clock %d MHz (ipc = warp-instructions / clk / SM sub-partition at max clock)\n", p.name, sms, clk / 1000);
  void* d;
  cudaMalloc(&d, 8 * 16 * 256);
  cudaMemset(d, 0, 8 * 16 * 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  RUN_S(12, 1, 2) RUN_S(12, 4, 8) RUN_S(12, 4, 1) RUN_S(12, 2, 6) RUN_S(12, 1, 3) RUN_S(12, 2, 4) RUN_S(12, 3, 6) RUN_S(12, 5, 7)
  RUN_S(8, 1, 2) RUN_S(8, 2, 4) RUN_S(8, 1, 4) RUN_S(16, 1, 2) RUN_S(16, 4, 8) RUN_S(16, 5, 10) RUN_S(16, 8, 1)
  RUN_P(8, 1, 2) RUN_P(8, 2, 4) RUN_P(8, 1, 4) RUN_P(12, 1, 2) RUN_P(12, 4, 8) RUN_P(12, 2, 6) RUN_P(12, 3, 6) RUN_P(12, 5, 7)
  RUN_M(12, 1, 2) RUN_M(12, 4, 8) RUN_M(12, 2, 6)
  return 0;
}
