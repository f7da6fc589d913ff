"""Static count of the ALGORITHMIC floating-point operations of the hot path:
the operations the reference's formulae prescribe (and both the oracle and the
kernels execute), each add/sub/mul/div/sqrt/compare = 1, an FMA = 2.  This is
the numerator of `roofline.achieved` (SURVEY.md §8d asks for an exact static
count of the restatement instead of the "≈" figures); instruction-level
expansion of div/sqrt on the GPU is deliberately NOT counted.

The counts follow principia_b200/csrc/pb_geopotential.cuh / pb_accel.cuh
statement by statement.
"""

POINT_MASS_PAIR = 19  # 3 sub, 5 norm², sqrt, cmp, mul+div, mul, 3 mul, 3 add.


def geopotential_setup():
  # InitializeSetup: 3 dots (5 each), x²+y² (3), sqrt, compare, 1/r_eq,
  # cos/sin lambda (2), 1/r, r_normalized (3), cos/sin beta (2),
  # grad_B (1+3 + 1+3 + 3 + 3 + 3), grad_L (3+3+3), R1/r (1).
  return 15 + 3 + 1 + 1 + 1 + 2 + 1 + 3 + 2 + 17 + 9 + 1


def order_update(n, m):
  if n == 2 and m == 1:
    return 0
  flops = 0
  if m == n:
    flops += 6 if m % 2 == 0 else 7
  if m == 0:
    flops += 5
  if m == n:
    pass
  elif m == n - 1:
    flops += 2
  elif m == n - 2:
    flops += 5
  else:
    flops += 7
  return flops


def order_term(n, m):
  flops = 1                       # B
  if m < n:
    flops += 2                    # cos_beta * cos_beta^m * P
  if m > 0:
    flops += 4                    # m * sin_beta * cos_beta^(m-1) * P, subtract
    flops += 3                    # L
  flops += 1 + 3                  # (B L) grad_sigma_R
  flops += 2 + 3                  # (sigma R/r L gradBpoly) grad_B
  flops += 3                      # sum
  if m > 0:
    flops += 4 + 3 + 3 + 3        # scalar chain, (S cos - C sin), * grad_L, +=
  flops += 3                      # normalisation
  return flops


def damping(damped):
  # compare; undamped: R' r_normalized (3); damped: r³, 3 products, sigma (2),
  # sigma' r (4), sigma R/r, (.. + ..) (3), * r_normalized (3).
  return 1 + (1 + 3 + 2 + 4 + 1 + 3 + 3 if damped else 3)


def geopotential(max_degree, zonal, damped_degrees=()):
  """One GeneralSphericalHarmonicsAcceleration evaluation."""
  flops = 2 + 1 + max_degree.bit_length()  # NaN test, zonal test, partition.
  flops += geopotential_setup()
  for n in range(2, max_degree + 1):
    size = 1 if zonal else n + 1
    flops += 2 + 1                # R_n/r, R'
    if n == 2 and size > 1:
      flops += 2 * damping(n in damped_degrees)
      flops += order_update(2, 0) + order_term(2, 0)
      flops += order_update(2, 2) + order_term(2, 2)
      flops += 3
    else:
      flops += damping(n in damped_degrees)
      for m in range(size):
        flops += order_update(n, m) + order_term(n, m)
      flops += 3 * (size - 1)     # right fold over orders
  flops += 3 * (max_degree - 2)   # right fold over degrees
  flops += 6                      # + zero vectors of degrees 0 and 1
  return flops


def massless_probe_stage_sol_leo():
  """One right-hand-side evaluation for a LEO probe in the 34-body Sol field:
  34 point masses, Earth at degree/order 10 (undamped), the Moon's damped J2
  (zonal regime at lunar distance), the other 9 oblate bodies beyond their
  outer thresholds (NaN test + partition only)."""
  flops = 34 * POINT_MASS_PAIR
  flops += geopotential(10, zonal=False) + 6   # Earth, acc += mu * effect
  flops += geopotential(2, zonal=True, damped_degrees=(2,)) + 6   # Moon
  flops += 9 * (2 + 6)                         # far oblate bodies
  return flops


SRKN_STAGE = 3 + 2 + 6 + 9       # q_stage, h a_i & h b_i, dv, dq
SRKN_STEP = 6 * 4                # six compensated increments


def srkn14a_probe_step_sol_leo():
  return 14 * (massless_probe_stage_sol_leo() + SRKN_STAGE) + 3 * 3 + SRKN_STEP


if __name__ == '__main__':
  print('geopotential setup', geopotential_setup())
  print('geopotential degree/order 10', geopotential(10, False))
  print('geopotential zonal J2 damped', geopotential(2, True, (2,)))
  print('probe-stage (Sol, LEO)', massless_probe_stage_sol_leo())
  print('SRKN14A probe-step', srkn14a_probe_step_sol_leo())


# ---------------------------------------------------------------------------
# Secondary configurations (bench.py `secondary`).

ALL_PAIRS_INTERACTION = 18  # POINT_MASS_PAIR without the collision test.

# DoublePrecision arithmetic (pb_math.cuh): TwoSum / TwoDifference 6,
# QuickTwoSum 3, Add / Subtract = 2 x 6 + 1 + 3 + 1 + 3, Scale 2, Increment 4.
DP_ADD = 20
DP_SCALE = 2
DP_INCREMENT = 4


def multistep_component_step(order):
  """One coordinate of one body through one step of
  SymmetricLinearMultistepIntegrator::Instance::Solve
  (pb_kernels_nbody.cuh EnsembleMultistepStepKernel): the double-double sum of
  the displacements, the weighted accelerations, the new position and the
  Cohen-Hubbard-Oesterwinter velocity."""
  half = order // 2
  flops = DP_SCALE + 1                                  # j = 0.
  flops += (half - 1) * (2 * (DP_SCALE + DP_ADD) + 3)   # j paired with k - j.
  flops += DP_SCALE + DP_ADD + 2                        # j = k / 2.
  flops += 3 + DP_INCREMENT                             # h² Σ / denominator.
  flops += DP_ADD                                       # Position.
  flops += DP_ADD + 2 + 2 * order + 3                   # Velocity.
  return flops


def ensemble_system_step(n_bodies, order, active_harmonics, damped_harmonics,
                         idle_harmonics):
  """One multistep step of one system of an ensemble: every ordered pair's
  point-mass term (the kernel evaluates each pair from both bodies), the
  degree-2 zonal harmonics of the oblate sources within their outer threshold
  (`damped_harmonics` of the `active_harmonics` beyond the inner threshold),
  the range tests of the others, and the multistep arithmetic."""
  flops = n_bodies * (n_bodies - 1) * ALL_PAIRS_INTERACTION
  undamped = active_harmonics - damped_harmonics
  flops += undamped * (geopotential(2, zonal=True) + 6)
  flops += damped_harmonics * (geopotential(2, zonal=True,
                                            damped_degrees=(2,)) + 6)
  flops += idle_harmonics * (2 + 6)
  flops += 3 * n_bodies * multistep_component_step(order)
  return flops


def ensemble_system_step_bytes(n_bodies, order):
  """HBM bytes of the same step: the history ring (displacement value, error
  and acceleration of `order` steps) is read once for the sum and its
  accelerations once more for the velocity; one slot and the 12 planes of the
  state are written."""
  per_component = (3 * order + order + 2) * 8 + 3 * 8 + 4 * 8
  return 3 * n_bodies * per_component


ERKN_STEP_OVERHEAD = 4 * 3 * 8 + 2 * 3 * 4 + 2 * 6 + 7 * DP_INCREMENT + 30


def adaptive_rhs_evaluation(earth_degree, polynomial_degree, n_bodies=34,
                            moon_j2=True):
  """One right-hand-side evaluation of the adaptive flow at a per-probe time:
  the Estrin evaluation of every body's three polynomials (2 flops per
  coefficient), the point masses, the Earth's field truncated at
  `earth_degree` (0: beyond the outer threshold of degree 2) and the stage
  arithmetic."""
  flops = n_bodies * (3 * 2 * polynomial_degree + 1 + POINT_MASS_PAIR)
  if earth_degree >= 2:
    flops += geopotential(earth_degree, zonal=False) + 6
  else:
    flops += 2 + 6
  if moon_j2:
    flops += geopotential(2, zonal=True, damped_degrees=(2,)) + 6
  flops += 9 * (2 + 6)
  flops += 3 * 8 + 9  # q_stage from the stage sums.
  return flops
