// Internal test hooks (not part of include/principia_b200.h): run the
// correctly-rounded elementary functions on the device so that the GPU tests
// can compare them with the oracle directly.
#include <cuda_runtime.h>

#include "pb_math.cuh"
#include "pb_runtime.cuh"

using namespace pb;

namespace {

__global__ void ElementaryFunctionsKernel(long long const n,
                                          double const* __restrict__ x,
                                          double* __restrict__ sin_x,
                                          double* __restrict__ cos_x,
                                          double* __restrict__ root4_x) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  double s, c;
  CrSinCos(x[i], s, c);
  sin_x[i] = s;
  cos_x[i] = c;
  root4_x[i] = CrRoot4(x[i]);
}

// Compares FastSqrt / FastDivide with the IEEE operators on pseudo-random
// operands spanning the fast range; counts the mismatches.
__global__ void FastMathCheckKernel(unsigned long long const seed,
                                    int const per_thread,
                                    unsigned long long* mismatches) {
  unsigned long long state =
      seed + 0x9E3779B97F4A7C15ull *
                 (static_cast<unsigned long long>(blockIdx.x) * blockDim.x +
                  threadIdx.x + 1);
  auto next = [&state]() {
    state ^= state << 13;
    state ^= state >> 7;
    state ^= state << 17;
    return state;
  };
  unsigned long long bad = 0;
  for (int k = 0; k < per_thread; ++k) {
    // Random mantissa, exponent in [-250, 250].
    unsigned long long const a_bits = next();
    unsigned long long const b_bits = next();
    int const ea = static_cast<int>(next() % 501) - 250;
    int const eb = static_cast<int>(next() % 501) - 250;
    double const a = __longlong_as_double(
        (a_bits & 0x000fffffffffffffull) |
        (static_cast<unsigned long long>(1023 + ea) << 52));
    double const b = __longlong_as_double(
        (b_bits & 0x000fffffffffffffull) |
        (static_cast<unsigned long long>(1023 + eb) << 52));
    if (FastSqrt(a) != sqrt(a)) {
      ++bad;
    }
    if (FastDivide(a, b) != a / b) {
      ++bad;
    }
    // The shape of the hot loop: d / (d2 * d2) with d = sqrt(d2).
    double const d2 = __longlong_as_double(
        (a_bits & 0x000fffffffffffffull) |
        (static_cast<unsigned long long>(1023 + ea / 4) << 52));
    double const d = sqrt(d2);
    if (FastDivide(d, d2 * d2) != d / (d2 * d2)) {
      ++bad;
    }
  }
  if (bad != 0) {
    atomicAdd(mismatches, bad);
  }
}

}  // namespace

extern "C" pb_status_t pb_internal_fast_math_check(int32_t cuda_device,
                                                   int64_t seed,
                                                   int64_t* n_checked,
                                                   int64_t* n_mismatches) {
  int device_count = 0;
  if (cudaGetDeviceCount(&device_count) != cudaSuccess || device_count == 0) {
    return Fail(PB_FAILED_PRECONDITION, "no CUDA device");
  }
  PB_CUDA(cudaSetDevice(cuda_device));
  DeviceBuffer<unsigned long long> counter;
  PB_RETURN_IF_ERROR(counter.Reserve(1));
  PB_CUDA(cudaMemset(counter.data, 0, sizeof(unsigned long long)));
  int const blocks = 148 * 16, threads = 256, per_thread = 600;
  FastMathCheckKernel<<<blocks, threads>>>(
      static_cast<unsigned long long>(seed), per_thread, counter.data);
  PB_LAUNCHED();
  unsigned long long bad = 0;
  PB_CUDA(cudaMemcpy(&bad, counter.data, sizeof(bad), cudaMemcpyDeviceToHost));
  *n_checked = 3ll * blocks * threads * per_thread;
  *n_mismatches = static_cast<int64_t>(bad);
  return PB_OK;
}

extern "C" pb_status_t pb_internal_elementary_functions(
    int32_t cuda_device, int64_t n, const double* x, double* sin_x,
    double* cos_x, double* root4_x) {
  int device_count = 0;
  if (cudaGetDeviceCount(&device_count) != cudaSuccess || device_count == 0) {
    return Fail(PB_FAILED_PRECONDITION, "no CUDA device");
  }
  PB_CUDA(cudaSetDevice(cuda_device));
  DeviceBuffer<double> in, s, c, r;
  PB_RETURN_IF_ERROR(in.Upload(x, n, nullptr));
  PB_RETURN_IF_ERROR(s.Reserve(n));
  PB_RETURN_IF_ERROR(c.Reserve(n));
  PB_RETURN_IF_ERROR(r.Reserve(n));
  ElementaryFunctionsKernel<<<static_cast<unsigned>((n + 127) / 128), 128>>>(
      n, in.data, s.data, c.data, r.data);
  PB_LAUNCHED();
  PB_CUDA(cudaMemcpy(sin_x, s.data, n * sizeof(double), cudaMemcpyDeviceToHost));
  PB_CUDA(cudaMemcpy(cos_x, c.data, n * sizeof(double), cudaMemcpyDeviceToHost));
  PB_CUDA(cudaMemcpy(root4_x, r.data, n * sizeof(double),
                     cudaMemcpyDeviceToHost));
  return PB_OK;
}
