// Large-N all-pairs point masses (BASELINE.json config 5): the accelerations of
// the owned targets [i_begin, i_end) due to all n sources.
//
// The arithmetic of one interaction is the reference's
// (physics/ephemeris_body.hpp:1366-1374): 18 algorithmic flops (3 sub,
// 5 norm², sqrt, mul, div, mul, 3 mul, 3 add); every ordered interaction is
// evaluated (no third-law sharing across threads).
//
// Work decomposition.  The sources are cut into L contiguous sub-ranges of
// `chunk` sources; L and chunk depend on n ONLY (AllPairsSplit), never on the
// sharding, so that a target's acceleration has the same bits whichever rank
// owns it:
//   a(i) = fold_g ( fold_{s in group g} ( sum_{j in sub-range s} term(i, j) ) )
// with sequential sums in increasing j and left folds in increasing s and g
// (8 sub-ranges per group).  The oracle restates exactly this order
// (oracle_capi.cpp, orc_allpairs_target_accelerations).  It differs from the
// reference's pair order (b1 < b2 with scattered third-law updates), hence a
// tolerance against the reference-ordered oracle, stated in the tests.
//
// A warp is one unit of work: 32 targets (one per lane) x one sub-range of
// sources.  The warp streams its sub-range through a private ring of
// shared-memory stages filled by 1-D bulk copies (cp.async.bulk = UBLKCP,
// completion on an mbarrier): no block barrier in the loop, the copy of stage
// k + kStages overlaps the arithmetic of stages k + 1 ...  A CTA is the 8
// sub-ranges of one group; its warps' partial sums are folded in order through
// shared memory.  With n / L sources per unit there are (n_local / 32) x L
// units -- at least 8 192 per GPU for n >= 2^17 sharded 8 ways -- so the 148
// SMs stay evenly loaded where one thread per target with a serial source
// loop left 128 CTAs for 148 SMs.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "pb_math.cuh"
#include "pb_tma.cuh"

namespace pb {

constexpr int kApWarps = 8;           // Sub-ranges per group = warps per CTA.
constexpr int kApThreads = 32 * kApWarps;
constexpr int kApTargets = 32;        // Targets per tile: one per lane.
constexpr int kApStageSources = 64;   // Sources per bulk copy (2 KiB).
constexpr int kApStages = 2;
constexpr int kApSourceDoubles = 4;   // x, y, z, mu.
constexpr int kApStageBytes =
    kApStageSources * kApSourceDoubles * static_cast<int>(sizeof(double));
constexpr int kApSharedBytes =
    kApWarps * kApStages * kApStageBytes +                        // Stages.
    kApWarps * 3 * kApTargets * static_cast<int>(sizeof(double)) +  // Partials.
    kApWarps * kApStages * static_cast<int>(sizeof(std::uint64_t));

struct AllPairsSplit {
  int sub_ranges;   // L: a multiple of kApWarps.
  long long chunk;  // Sources per sub-range: a multiple of kApStageSources.
};

// L = 2^23 / n clamped to [8, 64] (a power of two), chunk = ceil(n / L) rounded
// up to whole stages.
PB_HD AllPairsSplit MakeAllPairsSplit(long long const n) {
  AllPairsSplit split;
  int sub_ranges = 64;
  while (sub_ranges > 8 && static_cast<long long>(sub_ranges) * n > (1LL << 23)) {
    sub_ranges /= 2;
  }
  split.sub_ranges = sub_ranges;
  long long const per_range = (n + sub_ranges - 1) / sub_ranges;
  split.chunk = (per_range + kApStageSources - 1) / kApStageSources *
                kApStageSources;
  return split;
}

struct AllPairsForceView {
  long long n;
  long long i_begin;
  long long n_local;
  AllPairsSplit split;
  // [n][4]: x, y, z, mu of every body, 32-byte aligned.
  double const* sources;
  // groups == 1: the accelerations, 3 planes of n_local.  Otherwise the group
  // partials, [tile][group][xyz][lane].
  double* out;
};

#if defined(__CUDACC__)

// One interaction with the reference's operations (the branch-free correctly
// rounded square root and division of pb_math.cuh).
struct ExactInteraction {
  // kMaybeSelf: the source may be the target itself; its interaction is
  // computed and discarded by a select, so that the loop has no branch (its
  // Δq is exactly zero, so a zero factor makes its addends +0.0, which leave
  // the sums unchanged).  Stages that cannot contain the target run without
  // the test and the select.
  template<bool kMaybeSelf>
  __device__ __forceinline__ static void Accumulate(
      double const sx, double const sy, double const sz, double const smu,
      double const xi, double const yi, double const zi, bool const is_self,
      bool& in_range, double& ax, double& ay, double& az) {
    // Δq = position of the source - position of the target: the acceleration
    // on b2 = i due to b1 = j.
    double const dx = sx - xi;
    double const dy = sy - yi;
    double const dz = sz - zi;
    double const d2 = dx * dx + dy * dy + dz * dz;
    if constexpr (kMaybeSelf) {
      in_range = in_range && (InFastRange(d2) || is_self);
    } else {
      in_range = in_range && InFastRange(d2);
    }
    double const d = FastSqrt(d2);
    double const one_over_d3 = FastDivide(d, d2 * d2);
    double mu_over_d3 = smu * one_over_d3;
    if constexpr (kMaybeSelf) {
      mu_over_d3 = is_self ? 0.0 : mu_over_d3;
    }
    ax += dx * mu_over_d3;
    ay += dy * mu_over_d3;
    az += dz * mu_over_d3;
  }
  // The IEEE operators, for the targets that met a distance outside the range
  // of the branch-free ones.
  __device__ __forceinline__ static void AccumulateSlow(
      double const sx, double const sy, double const sz, double const smu,
      double const xi, double const yi, double const zi, double& ax,
      double& ay, double& az) {
    double const dx = sx - xi;
    double const dy = sy - yi;
    double const dz = sz - zi;
    double const d2 = dx * dx + dy * dy + dz * dz;
    double const d = sqrt(d2);
    double const one_over_d3 = d / (d2 * d2);
    double const mu_over_d3 = smu * one_over_d3;
    ax += dx * mu_over_d3;
    ay += dy * mu_over_d3;
    az += dz * mu_over_d3;
  }
};

// PB_ARITHMETIC_CONTRACTED: the same formula with fused multiply-adds and
// r^-3 = (rsqrt r²)³, the reciprocal square root from the hardware seed and one
// third-order correction (relative error below 2^-50).  17 FP64 instructions
// per interaction instead of 40.  Not bit-identical to the reference; within
// BASELINE.json's tolerance (tests/test_gpu_contracted.py).
struct ContractedInteraction {
  template<bool kMaybeSelf>
  __device__ __forceinline__ static void Accumulate(
      double const sx, double const sy, double const sz, double const smu,
      double const xi, double const yi, double const zi, bool const is_self,
      bool& in_range, double& ax, double& ay, double& az) {
    double const dx = sx - xi;
    double const dy = sy - yi;
    double const dz = sz - zi;
    double const d2 = fma(dz, dz, fma(dy, dy, dx * dx));
    if constexpr (kMaybeSelf) {
      in_range = in_range && (InFastRange(d2) || is_self);
    } else {
      in_range = in_range && InFastRange(d2);
    }
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d2));
    double const e = fma(d2 * y0, -y0, 1.0);
    double const p = fma(e, 0.375, 0.5);
    double const y1 = fma(p, y0 * e, y0);
    double mu_over_d3 = smu * (y1 * y1 * y1);
    if constexpr (kMaybeSelf) {
      mu_over_d3 = is_self ? 0.0 : mu_over_d3;
    }
    ax = fma(dx, mu_over_d3, ax);
    ay = fma(dy, mu_over_d3, ay);
    az = fma(dz, mu_over_d3, az);
  }
  __device__ __forceinline__ static void AccumulateSlow(
      double const sx, double const sy, double const sz, double const smu,
      double const xi, double const yi, double const zi, double& ax,
      double& ay, double& az) {
    ExactInteraction::AccumulateSlow(sx, sy, sz, smu, xi, yi, zi, ax, ay, az);
  }
};

template<typename Interaction>
__global__ void __launch_bounds__(kApThreads, 4)
AllPairsAccelerationKernel(AllPairsForceView const v) {
  extern __shared__ __align__(128) unsigned char shared[];
  int const warp = threadIdx.x >> 5;
  int const lane = threadIdx.x & 31;
  int const groups = v.split.sub_ranges / kApWarps;
  long long const tile = blockIdx.x / groups;
  int const group = static_cast<int>(blockIdx.x % groups);

  double* const stages = reinterpret_cast<double*>(shared) +
                         warp * kApStages * kApStageSources * kApSourceDoubles;
  double* const partials = reinterpret_cast<double*>(
      shared + kApWarps * kApStages * kApStageBytes);
  std::uint64_t* const barriers =
      reinterpret_cast<std::uint64_t*>(
          shared + kApWarps * kApStages * kApStageBytes +
          kApWarps * 3 * kApTargets * sizeof(double)) +
      warp * kApStages;

  // This warp's sub-range of the sources.
  long long const sub_range = static_cast<long long>(group) * kApWarps + warp;
  long long const j_begin = sub_range * v.split.chunk;
  long long const j_end =
      j_begin + v.split.chunk < v.n ? j_begin + v.split.chunk : v.n;
  int const n_stages =
      j_begin < v.n ? static_cast<int>((j_end - j_begin + kApStageSources - 1) /
                                       kApStageSources)
                    : 0;

  // This lane's target; a lane beyond the end shadows the first target.
  long long const i_local = tile * kApTargets + lane;
  long long const i = v.i_begin + (i_local < v.n_local ? i_local : 0);
  double const xi = v.sources[kApSourceDoubles * i];
  double const yi = v.sources[kApSourceDoubles * i + 1];
  double const zi = v.sources[kApSourceDoubles * i + 2];

  auto issue = [&](int const k) {
    long long const j0 = j_begin + static_cast<long long>(k) * kApStageSources;
    std::uint32_t const bytes = static_cast<std::uint32_t>(
        (j_end - j0 < kApStageSources ? j_end - j0 : kApStageSources) *
        kApSourceDoubles * sizeof(double));
    int const slot = k % kApStages;
    MbarrierArriveExpectTx(&barriers[slot], bytes);
    BulkCopyGlobalToShared(
        stages + slot * kApStageSources * kApSourceDoubles,
        v.sources + kApSourceDoubles * j0, bytes, &barriers[slot]);
  };

  if (lane == 0) {
    for (int slot = 0; slot < kApStages; ++slot) {
      MbarrierInit(&barriers[slot], 1);
    }
    MbarrierInitFence();
    for (int k = 0; k < kApStages && k < n_stages; ++k) {
      issue(k);
    }
  }
  __syncwarp();

  double ax = 0.0, ay = 0.0, az = 0.0;
  bool in_range = true;
  for (int k = 0; k < n_stages; ++k) {
    int const slot = k % kApStages;
    MbarrierWait(&barriers[slot], (k / kApStages) & 1);
    long long const j0 = j_begin + static_cast<long long>(k) * kApStageSources;
    int const count = static_cast<int>(
        j_end - j0 < kApStageSources ? j_end - j0 : kApStageSources);
    // Index of the target within the stage, if it is there.
    long long const self_offset = i - j0;
    int const self = self_offset >= 0 && self_offset < count
                         ? static_cast<int>(self_offset)
                         : -1;
    double2 const* const stage = reinterpret_cast<double2 const*>(
        stages + slot * kApStageSources * kApSourceDoubles);
    // `self` differs between the lanes, but a stage holds the target of every
    // lane or of none when the tile lies inside it, and of no lane in all the
    // other stages: the vote keeps the warp on one path.
    if (count == kApStageSources && __all_sync(0xffffffffu, self < 0)) {
#pragma unroll 4
      for (int jj = 0; jj < kApStageSources; ++jj) {
        double2 const xy = stage[2 * jj];
        double2 const zm = stage[2 * jj + 1];
        Interaction::template Accumulate<false>(xy.x, xy.y, zm.x, zm.y, xi, yi,
                                                zi, false, in_range, ax, ay,
                                                az);
      }
    } else {
#pragma unroll 2
      for (int jj = 0; jj < count; ++jj) {
        double2 const xy = stage[2 * jj];
        double2 const zm = stage[2 * jj + 1];
        Interaction::template Accumulate<true>(xy.x, xy.y, zm.x, zm.y, xi, yi,
                                               zi, jj == self, in_range, ax,
                                               ay, az);
      }
    }
    // Every lane has read the stage: it may be refilled.
    __syncwarp();
    if (lane == 0 && k + kApStages < n_stages) {
      AsyncProxyFence();
      issue(k + kApStages);
    }
  }
  if (!in_range) {
    // Some distance is outside the range of the branch-free operators (two
    // bodies at the same place, a non-finite state): redo this target's
    // sub-range with the IEEE operators, in the same order.
    ax = ay = az = 0.0;
    for (long long j = j_begin; j < j_end; ++j) {
      if (j == i) {
        continue;
      }
      Interaction::AccumulateSlow(
          v.sources[kApSourceDoubles * j], v.sources[kApSourceDoubles * j + 1],
          v.sources[kApSourceDoubles * j + 2],
          v.sources[kApSourceDoubles * j + 3], xi, yi, zi, ax, ay, az);
    }
  }

  // Left fold of the 8 sub-ranges of the group, in order.
  partials[(warp * 3 + 0) * kApTargets + lane] = ax;
  partials[(warp * 3 + 1) * kApTargets + lane] = ay;
  partials[(warp * 3 + 2) * kApTargets + lane] = az;
  __syncthreads();
  if (threadIdx.x < 3 * kApTargets) {
    int const c = threadIdx.x / kApTargets;
    int const l = threadIdx.x % kApTargets;
    double sum = partials[c * kApTargets + l];
#pragma unroll
    for (int w = 1; w < kApWarps; ++w) {
      sum = sum + partials[(w * 3 + c) * kApTargets + l];
    }
    long long const target = tile * kApTargets + l;
    if (groups == 1) {
      if (target < v.n_local) {
        v.out[c * v.n_local + target] = sum;
      }
    } else {
      v.out[((tile * groups + group) * 3 + c) * kApTargets + l] = sum;
    }
  }
}

// Left fold of the group partials of every target, in order.
static __global__ void AllPairsCombineKernel(long long const n_local,
                                             int const groups,
                                             double const* __restrict__ partials,
                                             double* __restrict__ g) {
  long long const index =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (index >= 3 * n_local) {
    return;
  }
  int const c = static_cast<int>(index / n_local);
  long long const target = index % n_local;
  long long const tile = target / kApTargets;
  int const lane = static_cast<int>(target % kApTargets);
  double sum = partials[((tile * groups) * 3 + c) * kApTargets + lane];
  for (int group = 1; group < groups; ++group) {
    sum = sum + partials[((tile * groups + group) * 3 + c) * kApTargets + lane];
  }
  g[index] = sum;
}

#endif  // __CUDACC__

}  // namespace pb
