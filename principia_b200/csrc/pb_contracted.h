// Interface of the translation unit that holds the fixed-step flow kernels
// compiled WITH floating-point contraction (pb_contracted.cu, nvcc
// --fmad=true): PB_ARITHMETIC_CONTRACTED.  The same sources as the exact
// kernels, in a namespace of their own (their types have the layout of the
// pb:: ones, hence the untyped pointers here).  Everything else -- stage
// tables, sinks, host orchestration -- is shared with the exact arithmetic.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace pb_contracted_api {

// The constant bank of that translation unit (its own c_bodies / c_hot).
cudaError_t UploadConstants(void const* body_scalars,
                            std::size_t body_scalars_bytes,
                            void const* hot_body, std::size_t hot_body_bytes,
                            cudaStream_t stream);

struct SrknLaunch {
  void const* view;     // pb::EphemerisView.
  void const* tableau;  // pb::SrknTableau.
  double h;
  double const* stage_table;
  double const* stage_times;
  double const* burns;
  int n_steps;
  long long n, i_begin, i_end;
  double* state;
  long long steps_before, sample_every, max_samples;
  double* samples;
  std::int32_t* status;
  int threads;  // 256 (2 blocks per SM), 384 (1) or 128 (3).
  unsigned grid;
  cudaStream_t stream;
  // SrknFlowSpecializedKernel (256 threads; writes block_done) instead of
  // SrknFlowKernel (which skips the blocks of skip_blocks, if any).
  bool specialized = false;
  std::int32_t* block_done = nullptr;
  std::int32_t const* skip_blocks = nullptr;
};
cudaError_t LaunchSrkn(SrknLaunch const& launch);

struct MultistepLaunch {
  void const* view;     // pb::EphemerisView.
  void const* method;   // pb::MultistepMethod.
  double h;
  double const* stage_table;
  double const* stage_times;
  double const* burns;
  int n_steps;
  long long n, i_begin, i_end;
  double* state;
  void const* history;  // pb::MasslessHistory.
  int oldest;
  long long steps_before, sample_every, max_samples;
  double* samples;
  std::int32_t* status;
  int order;  // 8 or 12.
  unsigned grid;
  cudaStream_t stream;
};
cudaError_t LaunchMultistep(MultistepLaunch const& launch);

struct FillStepLaunch {
  void const* view;
  double const* stage;
  double time;
  double const* burns;
  long long n;
  double const* state;
  void const* history;
  int slot;
  std::int32_t* status;
  cudaStream_t stream;
};
cudaError_t LaunchFillStep(FillStepLaunch const& launch);

}  // namespace pb_contracted_api
