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

}  // namespace

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
