"""GPU parity of the Чебышёв-basis series evaluation (SURVEY.md §8 row a3'):
the reference's exact known answers and random series against the oracle, bit
for bit, through the C ABI."""
import numpy as np
import pytest

from principia_b200 import series

import oracle_lib
from chebyshev_kats import KATS

pytestmark = pytest.mark.gpu


def pack(coefficient_lists, dimension=1):
  c = np.zeros((len(coefficient_lists), series.MAX_DEGREE + 1, dimension))
  for i, coefficients in enumerate(coefficient_lists):
    c[i, :len(coefficients), :] = np.asarray(coefficients).reshape(
        len(coefficients), -1)
  return c


def test_reference_known_answers():
  degree = [len(k[0]) - 1 for k in KATS]
  lower = [k[1] for k in KATS]
  upper = [k[2] for k in KATS]
  coefficients = pack([k[0] for k in KATS])
  index, arguments, expected_values = [], [], []
  d_index, d_arguments, expected_derivatives = [], [], []
  for i, (_, _, _, values, derivatives) in enumerate(KATS):
    for argument, value in values:
      index.append(i)
      arguments.append(argument)
      expected_values.append(value)
    for argument, derivative in derivatives:
      d_index.append(i)
      d_arguments.append(argument)
      expected_derivatives.append(derivative)
  values, _ = series.chebyshev_series_evaluate(
      degree, lower, upper, coefficients, index, arguments,
      want_derivatives=False)
  assert list(values[:, 0]) == expected_values
  _, derivatives = series.chebyshev_series_evaluate(
      degree, lower, upper, coefficients, d_index, d_arguments,
      want_values=False)
  assert list(derivatives[:, 0]) == expected_derivatives


def test_random_series_match_oracle_bitwise():
  orc = oracle_lib.library()
  rng = np.random.default_rng(3)
  n_series, n_evaluations, dimension = 200, 20000, 3
  degree = rng.integers(0, series.MAX_DEGREE + 1, n_series).astype(np.int32)
  lower = rng.uniform(-1e6, 1e6, n_series)
  upper = lower + rng.uniform(1.0, 1e5, n_series)
  coefficients = (rng.normal(size=(n_series, series.MAX_DEGREE + 1, dimension))
                  * 10.0**rng.uniform(-3, 3, (n_series, 1, 1)))
  index = rng.integers(0, n_series, n_evaluations).astype(np.int32)
  arguments = lower[index] + rng.random(n_evaluations) * (
      upper[index] - lower[index])
  # The bounds themselves, where the scaled argument is +-1 within 2 ulps.
  arguments[:n_series] = lower[index[:n_series]]
  arguments[n_series:2 * n_series] = upper[index[n_series:2 * n_series]]
  values, derivatives = series.chebyshev_series_evaluate(
      degree, lower, upper, coefficients, index, arguments)
  expected_values = np.zeros_like(values)
  expected_derivatives = np.zeros_like(derivatives)
  for j in range(n_evaluations):
    i = index[j]
    for c in range(dimension):
      column = np.ascontiguousarray(coefficients[i, :, c])
      expected_values[j, c] = orc.orc_chebyshev_evaluate(
          oracle_lib.dptr(column), int(degree[i]), lower[i], upper[i],
          arguments[j])
      expected_derivatives[j, c] = orc.orc_chebyshev_evaluate_derivative(
          oracle_lib.dptr(column), int(degree[i]), lower[i], upper[i],
          arguments[j])
  assert values.tobytes() == expected_values.tobytes()
  assert derivatives.tobytes() == expected_derivatives.tobytes()


def test_bad_arguments_are_rejected():
  from principia_b200.ephemeris import StatusError
  with pytest.raises(StatusError):
    series.chebyshev_series_evaluate([21], [0.0], [1.0], pack([[1.0]]), [0],
                                     [0.5])
  with pytest.raises(StatusError):
    series.chebyshev_series_evaluate([0], [0.0], [1.0], pack([[1.0]]), [1],
                                     [0.5])
