// The fixed-step SRKN flow with the two halves of the right-hand side on
// different warps: SymplecticRungeKuttaNyströmIntegrator::Instance::Solve,
// integrators/symplectic_runge_kutta_nyström_integrator_body.hpp:97-141, over
// Ephemeris::ComputeGravitationalAccelerationByAllMassiveBodiesOnMasslessBodies,
// physics/ephemeris_body.hpp:1412-1462,1554-1591.
//
// Why.  In SrknFlowKernel one thread does everything for its probe.  The
// unrolled degree-10 geopotential wants ~200 registers, the 34 point-mass
// terms and the integrator want ~50 but are bound by instruction issue and by
// the latency of their serial glue (state in thread-local memory around the
// harmonics call); at the 128 registers that 16 warps per SM allow, the
// harmonics spill and the rest of the code runs at 4 warps per scheduler: the
// FP64 pipe is 71 % busy (profiles/r02t).  Here a block is two warpgroups:
//   * the HARMONICS warpgroup (setmaxnreg 192: no spills) evaluates nothing but
//     mu * (spherical harmonics) of the hot and warm bodies;
//   * the POINT-MASS warpgroup (setmaxnreg 64) evaluates the 34 point-mass terms,
//     adds everything in the reference's order and advances the integrator,
//     whose state it keeps in shared memory.
// 2 blocks per SM: (192 + 64) x 128 x 2 = 65 536 registers.  Each thread pair
// owns TWO probes (sets A and B of 128 probes per block): while the
// point-mass warps finish stage k of set A the harmonics warps are already on
// stage k of set B, so the harmonics warps never wait.  The hand-over is
// through shared memory under named barriers (bar.sync / bar.arrive, 256
// threads each): Q_READY[set] (position of the stage published) and
// H_READY[set] (harmonics published).
//
// The arithmetic is that of AccelerationOnMasslessBodyGrouped, term for term,
// so the bits are the reference's.  The kernel only handles the common case:
// when a probe needs anything else (harmonics of a body that is neither hot
// nor warm, a distance outside the range of the branch-free square root, a
// zero dividend in a Legendre recurrence) its block leaves the state untouched
// and clears its entry of `block_done`; SrknFlowKernel, launched right after
// with the same geometry and `block_done` as its skip list, then flows
// exactly those blocks.
#pragma once

#include <cuda_runtime.h>

#include "pb_kernels_massless.cuh"

namespace pb {

constexpr int kSpecializedThreads = 256;
constexpr int kSpecializedSet = 128;  // One warpgroup.
constexpr int kHarmonicsRegisters = 192;
constexpr int kPointMassRegisters = 64;
// Bodies per group of the point-mass warpgroup's loops (register budget).
#ifndef PB_SPECIALIZED_GROUP
#define PB_SPECIALIZED_GROUP 2
#endif
constexpr int kSpecializedGroup = PB_SPECIALIZED_GROUP;

// Named barriers (0 is __syncthreads).
constexpr int kBarrierQReady = 1;  // + set.
constexpr int kBarrierHReady = 3;  // + set.

struct SpecializedShared {
  // Position at the current stage, published by the point-mass warpgroup.
  double q_stage[2][3][kSpecializedSet];
  // Published by the harmonics warpgroup, for the hot (0) and warm (1) body:
  // point-mass term and mu1 * harmonics, xyz each.
  double terms[2][2][6][kSpecializedSet];
  // Bit 0: a special body is inside its collision radius; bits 8..: the range
  // offset of the squared distances did not fit, see below.
  unsigned h_flags[2][kSpecializedSet];
  // Integrator state: q.value, q.error, v.value, v.error, dq, dv (xyz each).
  double state[2][18][kSpecializedSet];
  int failed;
};

__device__ __forceinline__ void NamedBarrierSync(int const id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kSpecializedThreads)
               : "memory");
}
__device__ __forceinline__ void NamedBarrierArrive(int const id) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kSpecializedThreads)
               : "memory");
}

// mu1 * harmonics of the hot body with the degree-kHotDegree non-zonal
// evaluation inline (so that it is compiled for the harmonics warpgroup's
// register budget); any other case out of line.
__device__ __forceinline__ Vec3 SpecializedHotHarmonics(
    EphemerisView const& e, double const* const stage, int const b,
    Vec3 const& dq, double const dq_norm, double const dq2,
    double const one_over_dq3, bool& failed) {
  int const max_degree =
      LimitingDegree(c_hot.degree_damping, c_hot.n_damping, dq_norm) - 1;
  DeviceBody const& body = c_hot.hot_data;
  bool const is_zonal =
      body.is_zonal || dq_norm > c_hot.sectoral.outer_threshold;
  if (max_degree == kHotDegree && !is_zonal) {
    double const* const axes = stage + 3 * e.n_bodies + 6 * body.frame_slot;
    Vec3 const x_hat{axes[0], axes[1], axes[2]};
    Vec3 const y_hat{axes[3], axes[4], axes[5]};
    GeopotentialSetup s;
    InitializeSetup(body, x_hat, y_hat, -dq, dq_norm, dq2, one_over_dq3, s);
    HotTables const tables;
    FlaggedScratch<kHotDegree> scratch;
    Vec3 const acceleration =
        GeopotentialUnrolledAcceleration<kHotDegree, false>(tables, s, scratch);
    failed = failed || scratch.Failed();
    return acceleration;
  }
  return FlowHarmonics(e, stage, b, dq, dq_norm, dq2, one_over_dq3);
}

template<int kDummy>
__global__ void __launch_bounds__(kSpecializedThreads, 2)
SrknFlowSpecializedKernel(EphemerisView const e, SrknTableau const tableau,
                          double const h,
                          double const* __restrict__ stage_table,
                          int const n_steps, long long const n,
                          long long const i_begin, long long const i_end,
                          double* __restrict__ state,
                          long long const sample_every,
                          // The first step of this launch (1-based) after which
                          // a sample is due, and its slot.
                          long long const first_sample_step,
                          long long const first_sample_slot,
                          long long const max_samples,
                          double* __restrict__ samples, int32_t* status,
                          int32_t* __restrict__ block_done) {
  extern __shared__ __align__(16) unsigned char shared_bytes[];
  SpecializedShared& shared =
      *reinterpret_cast<SpecializedShared*>(shared_bytes);
  int const lane = threadIdx.x % kSpecializedSet;
  bool const harmonics_role = threadIdx.x < kSpecializedSet;
  if (threadIdx.x == 0) {
    shared.failed = 0;
  }
  __syncthreads();
  int const evaluations = tableau.stages - tableau.first_stage;
  int const total = n_steps * evaluations;
  int const n_bodies = e.n_bodies;
  // The special bodies in internal order; n_bodies: none.
  int const hot = c_hot.body >= 0 ? c_hot.body : n_bodies;
  int const warm = c_hot.warm_body >= 0 ? c_hot.warm_body : n_bodies;

  if (harmonics_role) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(
        kHarmonicsRegisters));
    bool failed = false;
    for (int g = 0; g < total; ++g) {
      double const* const stage =
          stage_table + static_cast<long long>(g) * e.stage_stride;
#pragma unroll 1
      for (int set = 0; set < 2; ++set) {
        NamedBarrierSync(kBarrierQReady + set);
        Vec3 const q{shared.q_stage[set][0][lane], shared.q_stage[set][1][lane],
                     shared.q_stage[set][2][lane]};
        unsigned range = 0;
        bool ok = true;
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {
          int const special = k == 0 ? hot : warm;
          double* const out = &shared.terms[set][k][0][lane];
          if (special >= n_bodies) {
            continue;
          }
          int const b = special;
          double const mu1 = c_bodies.mu[b];
          double const dx = __ldg(stage + 3 * b) - q.x;
          double const dy = __ldg(stage + 3 * b + 1) - q.y;
          double const dz = __ldg(stage + 3 * b + 2) - q.z;
          double const dq2 = dx * dx + dy * dy + dz * dz;
          range = FastRangeOffsetMax(dq2, range);
          double const safe_dq2 = range < kFastRangeLimit ? dq2 : 1.0;
          double const dq_norm = FastSqrt(safe_dq2);
          ok = ok && dq_norm > c_bodies.collision_radius[b];
          double const one_over_dq3 = FastDivide(dq_norm, safe_dq2 * safe_dq2);
          double const mu1_over_dq3 = mu1 * one_over_dq3;
          out[0 * kSpecializedSet] = dx * mu1_over_dq3;
          out[1 * kSpecializedSet] = dy * mu1_over_dq3;
          out[2 * kSpecializedSet] = dz * mu1_over_dq3;
          Vec3 effect{0.0, 0.0, 0.0};
          if (range < kFastRangeLimit &&
              !(dq_norm >= c_bodies.outer_threshold_degree_2[b])) {
            effect = k == 0 ? SpecializedHotHarmonics(
                                  e, stage, b, Vec3{dx, dy, dz}, dq_norm,
                                  safe_dq2, one_over_dq3, failed)
                            : FlowHarmonics(e, stage, b, Vec3{dx, dy, dz},
                                            dq_norm, safe_dq2, one_over_dq3);
          }
          out[3 * kSpecializedSet] = mu1 * effect.x;
          out[4 * kSpecializedSet] = mu1 * effect.y;
          out[5 * kSpecializedSet] = mu1 * effect.z;
        }
        shared.h_flags[set][lane] =
            (ok ? 0u : 1u) | (range >= kFastRangeLimit || failed ? 2u : 0u);
        __threadfence_block();
        NamedBarrierArrive(kBarrierHReady + set);
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(
        kPointMassRegisters));
    int const n_oblate = e.n_oblate;
    int const special[2] = {min(hot, warm), max(hot, warm)};
    // Which published slot (0: hot, 1: warm) the k-th special body is.
    int const slot_of[2] = {hot <= warm ? 0 : 1, hot <= warm ? 1 : 0};
    // Probes of this thread: set A and set B (clamped past the end; those
    // results are not written).
    long long probe[2];
    bool live[2];
#pragma unroll
    for (int set = 0; set < 2; ++set) {
      long long const index = i_begin +
                              static_cast<long long>(blockIdx.x) *
                                  kSpecializedThreads +
                              set * kSpecializedSet + lane;
      live[set] = index < i_end;
      probe[set] = live[set] ? index : i_end - 1;
    }
    bool ok = true;
    bool failed = false;
    // Load the state; start the first step; publish the first positions.
#pragma unroll 1
    for (int set = 0; set < 2; ++set) {
      long long const i = probe[set];
      double (&s)[18][kSpecializedSet] = shared.state[set];
#pragma unroll
      for (int c = 0; c < 12; ++c) {
        s[c][lane] = state[c * n + i];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double dq = 0.0;
        double const dv = 0.0;
        if (tableau.first_stage == 1) {
          // exp(a0 h A), :108.
          double const ha = h * tableau.a[0];
          dq += ha * (s[6 + c][lane] + dv);
        }
        s[12 + c][lane] = dq;
        s[15 + c][lane] = dv;
        shared.q_stage[set][c][lane] = s[c][lane] + dq;
      }
      __threadfence_block();
      NamedBarrierArrive(kBarrierQReady + set);
    }
    int step = 0;
    int st = tableau.first_stage;
    // Sampling without integer divisions here (their out-of-line routines are
    // not compiled for this warpgroup's register budget).
    long long next_sample = first_sample_step;
    long long sample_slot = first_sample_slot;
    for (int g = 0; g < total; ++g) {
      double const* const stage =
          stage_table + static_cast<long long>(g) * e.stage_stride;
      double const hb = h * tableau.b[st];
      double const ha = h * tableau.a[st];
      bool const last_stage = st == tableau.stages - 1;
#pragma unroll 1
      for (int set = 0; set < 2; ++set) {
        double (&s)[18][kSpecializedSet] = shared.state[set];
        Vec3 const q{shared.q_stage[set][0][lane], shared.q_stage[set][1][lane],
                     shared.q_stage[set][2][lane]};
        NamedBarrierSync(kBarrierHReady + set);
        unsigned const h_flags = shared.h_flags[set][lane];
        ok = ok && (h_flags & 1u) == 0;
        failed = failed || (h_flags & 2u) != 0;
        double ax = 0.0, ay = 0.0, az = 0.0;
        unsigned range = 0;
        bool active = false;
        int begin = 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (special[k] < n_bodies) {
            PointMassSegment<kSpecializedGroup>(stage, begin, special[k],
                                                n_oblate, q, ax, ay, az, ok,
                                                range, active);
            double const* const t = &shared.terms[set][slot_of[k]][0][lane];
            ax += t[0 * kSpecializedSet];
            ay += t[1 * kSpecializedSet];
            az += t[2 * kSpecializedSet];
            ax += t[3 * kSpecializedSet];
            ay += t[4 * kSpecializedSet];
            az += t[5 * kSpecializedSet];
            begin = special[k] + 1;
          }
        }
        PointMassSegment<kSpecializedGroup>(stage, begin, n_bodies, n_oblate, q,
                                            ax, ay, az, ok, range, active);
        failed = failed || active || range >= kFastRangeLimit;
        // The stage: dv += h b g, dq += h a (v + dv), :121-129.
        double const g3[3] = {ax, ay, az};
        double new_q_stage[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double dv = s[15 + c][lane];
          double dq = s[12 + c][lane];
          double const v = s[6 + c][lane];
          dv += hb * g3[c];
          dq += ha * (v + dv);
          if (last_stage) {
            // End of the step: the compensated increments, :131-135, and the
            // beginning of the next step.
            DP q_dp{s[c][lane], s[3 + c][lane]};
            DP v_dp{v, s[9 + c][lane]};
            Increment(q_dp, dq);
            Increment(v_dp, dv);
            s[c][lane] = q_dp.value;
            s[3 + c][lane] = q_dp.error;
            s[6 + c][lane] = v_dp.value;
            s[9 + c][lane] = v_dp.error;
            dq = 0.0;
            dv = 0.0;
            if (tableau.first_stage == 1) {
              double const ha0 = h * tableau.a[0];
              dq += ha0 * (v_dp.value + dv);
            }
            new_q_stage[c] = q_dp.value + dq;
          } else {
            new_q_stage[c] = s[c][lane] + dq;
          }
          s[12 + c][lane] = dq;
          s[15 + c][lane] = dv;
        }
        if (g + 1 < total) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            shared.q_stage[set][c][lane] = new_q_stage[c];
          }
          __threadfence_block();
          NamedBarrierArrive(kBarrierQReady + set);
        }
        if (last_stage && sample_every > 0 && step + 1 == next_sample &&
            sample_slot < max_samples && live[set]) {
          double* const out = samples + sample_slot * 6 * n;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            out[c * n + probe[set]] = s[c][lane];
            out[(3 + c) * n + probe[set]] = s[6 + c][lane];
          }
        }
      }
      if (last_stage) {
        if (step + 1 == next_sample) {
          next_sample += sample_every;
          ++sample_slot;
        }
        ++step;
        st = tableau.first_stage;
      } else {
        ++st;
      }
    }
    // Bit 0: this block has to be flowed by the general kernel; bit 2: a
    // collision (reported only if the block completes, below).
    if (failed) {
      atomicOr(&shared.failed, 1);
    }
    if (!ok) {
      atomicOr(&shared.failed, 4);
    }
  }
  __syncthreads();
  int const outcome = shared.failed;
  bool const done = (outcome & 3) == 0;
  if (!harmonics_role && done) {
#pragma unroll 1
    for (int set = 0; set < 2; ++set) {
      long long const index = i_begin +
                              static_cast<long long>(blockIdx.x) *
                                  kSpecializedThreads +
                              set * kSpecializedSet + lane;
      if (index < i_end) {
#pragma unroll
        for (int c = 0; c < 12; ++c) {
          state[c * n + index] = shared.state[set][c][lane];
        }
      }
    }
  }
  if (threadIdx.x == 0) {
    block_done[blockIdx.x] = done ? 1 : 0;
    if (done && (outcome & 4) != 0) {
      atomicOr(status, 11);
    }
  }
}

}  // namespace pb
