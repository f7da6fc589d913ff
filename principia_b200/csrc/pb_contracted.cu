// PB_ARITHMETIC_CONTRACTED: the fixed-step flow kernels of the massless path
// compiled with floating-point contraction (this file alone is built with
// nvcc --fmad=true; see the Makefile).  The kernel sources are the exact
// ones; they are instantiated here in a namespace of their own so that the
// two builds of every template and inline function keep distinct symbols.
//
// The reference forbids contraction (-ffp-contract=off, Makefile:123-169
// there), so these kernels are not bit-identical to it: a*b + c rounds once
// instead of twice.  They are within BASELINE.json's tolerance (relative
// position error <= 1e-12 after 10^4 steps, tests/test_gpu_contracted.py) and
// execute ~30 % fewer FP64 instructions.  Opt-in, per call.
#define pb pb_contracted
// Selects the cheaper formulations that only the contracted arithmetic may
// use (NormAndInverseCube, FlaggedScratch::Divide).
#define PB_CONTRACTED_BUILD 1
// Point-mass terms three at a time: measured 0.518 of the FP64 peak against
// 0.506 and 0.504 for two and four (with a third of the FP64 instructions of
// the exact build gone, the loop overhead per body weighs more).
#define PB_POINT_MASS_GROUP 3
#include "pb_kernels_massless.cuh"
#include "pb_kernels_massless_multistep.cuh"
#include "pb_kernels_massless_specialized.cuh"
#undef pb

#include "pb_contracted.h"

namespace pb_contracted_api {

namespace k = pb_contracted;

cudaError_t UploadConstants(void const* const body_scalars,
                            std::size_t const body_scalars_bytes,
                            void const* const hot_body,
                            std::size_t const hot_body_bytes,
                            cudaStream_t const stream) {
  if (body_scalars_bytes != sizeof(k::BodyScalars) ||
      hot_body_bytes != sizeof(k::HotBody)) {
    return cudaErrorInvalidValue;
  }
  cudaError_t error = cudaMemcpyToSymbolAsync(
      k::c_bodies, body_scalars, sizeof(k::BodyScalars), 0,
      cudaMemcpyHostToDevice, stream);
  if (error == cudaSuccess) {
    error = cudaMemcpyToSymbolAsync(k::c_hot, hot_body, sizeof(k::HotBody), 0,
                                    cudaMemcpyHostToDevice, stream);
  }
  if (error == cudaSuccess) {
    error = cudaStreamSynchronize(stream);
  }
  return error;
}

cudaError_t LaunchSrkn(SrknLaunch const& l) {
  auto const& view = *static_cast<k::EphemerisView const*>(l.view);
  auto const& tableau = *static_cast<k::SrknTableau const*>(l.tableau);
#define PB_LAUNCH(kThreads, kMinBlocks)                                       \
  {                                                                           \
    cudaError_t const attribute =                                             \
        k::SrknStageCacheAttribute<kThreads, kMinBlocks>();                   \
    if (attribute != cudaSuccess) {                                           \
      return attribute;                                                       \
    }                                                                         \
  }                                                                           \
  (l.burns != nullptr ? k::SrknFlowKernel<kThreads, kMinBlocks, true>         \
                      : k::SrknFlowKernel<kThreads, kMinBlocks, false>)       \
      <<<l.grid, kThreads,                                                    \
         k::SrknStageCacheBytes(kThreads, view.stage_stride), l.stream>>>(    \
          view, tableau, l.h, l.stage_table, l.stage_times, l.burns,          \
          l.n_steps, l.n, l.i_begin, l.i_end, l.state, l.steps_before,        \
          l.sample_every, l.max_samples, l.samples, l.status, l.skip_blocks)
  if (l.specialized) {
    static cudaError_t const attribute = cudaFuncSetAttribute(
        k::SrknFlowSpecializedKernel<0>,
        cudaFuncAttributeMaxDynamicSharedMemorySize,
        static_cast<int>(sizeof(k::SpecializedShared)));
    if (attribute != cudaSuccess) {
      return attribute;
    }
    k::SrknFlowSpecializedKernel<0>
        <<<l.grid, k::kSpecializedThreads, sizeof(k::SpecializedShared),
           l.stream>>>(view, tableau, l.h, l.stage_table, l.n_steps, l.n,
                       l.i_begin, l.i_end, l.state, l.sample_every,
                       l.sample_every > 0
                           ? l.sample_every - l.steps_before % l.sample_every
                           : 0,
                       l.sample_every > 0 ? l.steps_before / l.sample_every : 0,
                       l.max_samples, l.samples, l.status, l.block_done);
    return cudaGetLastError();
  }
  switch (l.threads) {
    case 256: PB_LAUNCH(256, 2); break;
    case 384: PB_LAUNCH(384, 1); break;
    case 128: PB_LAUNCH(128, 3); break;
    default: return cudaErrorInvalidValue;
  }
#undef PB_LAUNCH
  return cudaGetLastError();
}

cudaError_t LaunchMultistep(MultistepLaunch const& l) {
  auto const& view = *static_cast<k::EphemerisView const*>(l.view);
  auto const& method = *static_cast<k::MultistepMethod const*>(l.method);
  auto const& history = *static_cast<k::MasslessHistory const*>(l.history);
#define PB_LAUNCH(kOrder)                                                     \
  k::MasslessMultistepFlowKernel<kOrder, 128, 3>                              \
      <<<l.grid, 128, 0, l.stream>>>(                                         \
          view, method, l.h, l.stage_table, l.stage_times, l.burns,           \
          l.n_steps, l.n, l.i_begin, l.i_end, l.state, history, l.oldest,     \
          l.steps_before, l.sample_every, l.max_samples, l.samples, l.status)
  switch (l.order) {
    case 8: PB_LAUNCH(8); break;
    case 12: PB_LAUNCH(12); break;
    default: return cudaErrorInvalidValue;
  }
#undef PB_LAUNCH
  return cudaGetLastError();
}

cudaError_t LaunchFillStep(FillStepLaunch const& l) {
  auto const& view = *static_cast<k::EphemerisView const*>(l.view);
  auto const& history = *static_cast<k::MasslessHistory const*>(l.history);
  if (l.n > 0) {
    k::MasslessFillStepKernel<<<static_cast<unsigned>(
                                    (l.n + k::kFlowThreads - 1) /
                                    k::kFlowThreads),
                                k::kFlowThreads, 0, l.stream>>>(
        view, l.stage, l.time, l.burns, l.n, l.state, history, l.slot,
        l.status);
  }
  return cudaGetLastError();
}

}  // namespace pb_contracted_api
