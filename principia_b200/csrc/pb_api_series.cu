// C ABI: batched evaluation of series in the Чебышёв basis
// (include/principia_b200.h), the legacy representation of
// ContinuousTrajectory segments.
#include <cuda_runtime.h>

#include "../../include/principia_b200.h"
#include "pb_math.cuh"
#include "pb_runtime.cuh"

using namespace pb;

namespace {

// One thread per (evaluation, component).
__global__ void __launch_bounds__(128)
ChebyshevSeriesKernel(int const dimension, int32_t const* __restrict__ degree,
                      double const* __restrict__ lower_bound,
                      double const* __restrict__ upper_bound,
                      double const* __restrict__ coefficients,
                      long long const n_evaluations,
                      int32_t const* __restrict__ series_index,
                      double const* __restrict__ arguments,
                      double* __restrict__ values,
                      double* __restrict__ derivatives) {
  long long const index =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (index >= n_evaluations * dimension) {
    return;
  }
  long long const evaluation = index / dimension;
  int const component = static_cast<int>(index % dimension);
  long long const series = series_index[evaluation];
  double const lower = lower_bound[series];
  double const upper = upper_bound[series];
  // PolynomialInЧебышёвBasis's constructor.
  double const width = upper - lower;
  double const one_over_width = 1 / width;
  double const* const c =
      coefficients +
      series * (PB_MAX_CHEBYSHEV_DEGREE + 1) * dimension + component;
  int const d = degree[series];
  double const t = arguments[evaluation];
  // The stride is a run-time value: walk the coefficients through a gather.
  double gathered[PB_MAX_CHEBYSHEV_DEGREE + 1];
  for (int k = 0; k <= d; ++k) {
    gathered[k] = c[static_cast<long long>(k) * dimension];
  }
  if (values != nullptr) {
    values[index] =
        ChebyshevEvaluate<1>(gathered, d, lower, upper, one_over_width, t);
  }
  if (derivatives != nullptr) {
    derivatives[index] = ChebyshevEvaluateDerivative<1>(
        gathered, d, lower, upper, one_over_width, t);
  }
}

}  // namespace

extern "C" {

pb_status_t pb_chebyshev_series_evaluate(
    int32_t cuda_device, int64_t n_series, int32_t dimension,
    const int32_t* degree, const double* lower_bound,
    const double* upper_bound, const double* coefficients,
    int64_t n_evaluations, const int32_t* series_index,
    const double* arguments, double* values, double* derivatives) {
  if (n_series < 0 || n_evaluations < 0 || dimension <= 0 ||
      (n_series > 0 && (degree == nullptr || lower_bound == nullptr ||
                        upper_bound == nullptr || coefficients == nullptr)) ||
      (n_evaluations > 0 &&
       (series_index == nullptr || arguments == nullptr))) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  for (int64_t i = 0; i < n_series; ++i) {
    if (degree[i] < 0 || degree[i] > PB_MAX_CHEBYSHEV_DEGREE) {
      return Fail(PB_INVALID_ARGUMENT, "series degree out of range");
    }
  }
  for (int64_t j = 0; j < n_evaluations; ++j) {
    if (series_index[j] < 0 || series_index[j] >= n_series) {
      return Fail(PB_INVALID_ARGUMENT, "series index out of range");
    }
  }
  PB_CUDA(cudaSetDevice(cuda_device));
  if (n_evaluations == 0) {
    return PB_OK;
  }
  cudaStream_t const stream = nullptr;
  std::size_t const series = static_cast<std::size_t>(n_series);
  std::size_t const evaluations = static_cast<std::size_t>(n_evaluations);
  std::size_t const outputs = evaluations * dimension;
  DeviceBuffer<int32_t> device_degree, device_index;
  DeviceBuffer<double> device_lower, device_upper, device_coefficients,
      device_arguments, device_values, device_derivatives;
  PB_RETURN_IF_ERROR(device_degree.Upload(degree, series, stream));
  PB_RETURN_IF_ERROR(device_lower.Upload(lower_bound, series, stream));
  PB_RETURN_IF_ERROR(device_upper.Upload(upper_bound, series, stream));
  PB_RETURN_IF_ERROR(device_coefficients.Upload(
      coefficients, series * (PB_MAX_CHEBYSHEV_DEGREE + 1) * dimension,
      stream));
  PB_RETURN_IF_ERROR(device_index.Upload(series_index, evaluations, stream));
  PB_RETURN_IF_ERROR(device_arguments.Upload(arguments, evaluations, stream));
  if (values != nullptr) {
    PB_RETURN_IF_ERROR(device_values.Reserve(outputs));
  }
  if (derivatives != nullptr) {
    PB_RETURN_IF_ERROR(device_derivatives.Reserve(outputs));
  }
  ChebyshevSeriesKernel<<<static_cast<unsigned>((outputs + 127) / 128), 128, 0,
                          stream>>>(
      dimension, device_degree.data, device_lower.data, device_upper.data,
      device_coefficients.data, n_evaluations, device_index.data,
      device_arguments.data, values != nullptr ? device_values.data : nullptr,
      derivatives != nullptr ? device_derivatives.data : nullptr);
  PB_LAUNCHED();
  if (values != nullptr) {
    PB_CUDA(cudaMemcpyAsync(values, device_values.data,
                            outputs * sizeof(double), cudaMemcpyDeviceToHost,
                            stream));
  }
  if (derivatives != nullptr) {
    PB_CUDA(cudaMemcpyAsync(derivatives, device_derivatives.data,
                            outputs * sizeof(double), cudaMemcpyDeviceToHost,
                            stream));
  }
  PB_CUDA(cudaStreamSynchronize(stream));
  return PB_OK;
}

}  // extern "C"
