// FP64 pipe of one SM sub-partition: dependent-issue latency of DFMA / DADD /
// DMUL, and throughput against (warps per SM) x (independent chains per warp).
// nvcc -arch=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template<int kIlp, int kOp>
__global__ void Chains(double* out, int const iterations, long long* cycles) {
  double x[kIlp];
#pragma unroll
  for (int j = 0; j < kIlp; ++j) {
    x[j] = threadIdx.x * 1e-9 + j;
  }
  double const a = 0.999999999, b = 1e-9;
  long long const start = clock64();
  for (int k = 0; k < iterations; ++k) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int j = 0; j < kIlp; ++j) {
        if (kOp == 0) {
          x[j] = fma(x[j], a, b);
        } else if (kOp == 1) {
          x[j] = __dadd_rn(x[j], b);
        } else {
          x[j] = __dmul_rn(x[j], a);
        }
      }
    }
  }
  long long const stop = clock64();
  double sum = 0;
#pragma unroll
  for (int j = 0; j < kIlp; ++j) {
    sum += x[j];
  }
  if (sum == 12345.678) {
    out[0] = sum;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    cycles[0] = stop - start;
  }
}

template<int kIlp, int kOp>
void Run(int const warps_per_sm, char const* const name) {
  double* out;
  long long* cycles;
  cudaMalloc(&out, 8);
  cudaMalloc(&cycles, 8);
  int const iterations = 2000;
  int const sms = 148;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  Chains<kIlp, kOp><<<sms, 32 * warps_per_sm>>>(out, 10, cycles);
  cudaEventRecord(e0);
  Chains<kIlp, kOp><<<sms, 32 * warps_per_sm>>>(out, iterations, cycles);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c;
  cudaMemcpy(&c, cycles, 8, cudaMemcpyDeviceToHost);
  double const instr_per_warp = double(iterations) * 8 * kIlp;
  double const warp_instr_per_sm = instr_per_warp * warps_per_sm;
  printf("%s ilp %d warps/SM %2d: %7.2f cycles per dependent step, "
         "%.3f warp-instr/cycle/SM (%.1f%% of 2/cycle), %.2f ms\n",
         name, kIlp, warps_per_sm, double(c) / (iterations * 8.0),
         warp_instr_per_sm / double(c), 100.0 * warp_instr_per_sm / double(c) / 2.0, ms);
  cudaFree(out);
  cudaFree(cycles);
}

int main() {
  for (int w : {1, 4, 8, 16, 24, 32}) {
    Run<1, 0>(w, "DFMA");
  }
  for (int w : {4, 16}) {
    Run<2, 0>(w, "DFMA");
    Run<4, 0>(w, "DFMA");
    Run<8, 0>(w, "DFMA");
  }
  for (int w : {1, 16}) {
    Run<1, 1>(w, "DADD");
    Run<1, 2>(w, "DMUL");
    Run<4, 1>(w, "DADD");
    Run<4, 2>(w, "DMUL");
  }
  return 0;
}
