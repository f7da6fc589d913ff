"""Series in the Чебышёв basis (PolynomialInЧебышёвBasis of the reference,
numerics/polynomial_in_чебышёв_basis.hpp), evaluated in batches on the GPU
through the C ABI."""
import ctypes

import numpy as np

from . import _capi
from .ephemeris import _check as check

MAX_DEGREE = 20


def chebyshev_series_evaluate(degree, lower_bound, upper_bound, coefficients,
                              series_index, arguments, device=0,
                              want_values=True, want_derivatives=True):
  """coefficients: [n_series][MAX_DEGREE + 1][dimension]; evaluation j is of
  series series_index[j] at arguments[j].  Returns (values, derivatives), each
  [n_evaluations][dimension] or None."""
  degree = np.ascontiguousarray(degree, dtype=np.int32)
  lower_bound = np.ascontiguousarray(lower_bound, dtype=np.float64)
  upper_bound = np.ascontiguousarray(upper_bound, dtype=np.float64)
  coefficients = np.ascontiguousarray(coefficients, dtype=np.float64)
  series_index = np.ascontiguousarray(series_index, dtype=np.int32)
  arguments = np.ascontiguousarray(arguments, dtype=np.float64)
  n_series = len(degree)
  assert coefficients.shape[:2] == (n_series, MAX_DEGREE + 1)
  dimension = coefficients.shape[2]
  n = len(arguments)
  values = np.zeros((n, dimension)) if want_values else None
  derivatives = np.zeros((n, dimension)) if want_derivatives else None
  dptr = lambda a: (a.ctypes.data_as(_capi.c_double_p)
                    if a is not None else None)
  check(_capi.library().pb_chebyshev_series_evaluate(
      device, n_series, dimension, degree.ctypes.data_as(_capi.c_int32_p),
      dptr(lower_bound), dptr(upper_bound), dptr(coefficients), n,
      series_index.ctypes.data_as(_capi.c_int32_p), dptr(arguments),
      dptr(values), dptr(derivatives)))
  return values, derivatives
