#!/usr/bin/env python3
"""Extracts the pinned figures of numerics/newhall_test.cpp (the
ApproximationInMonomialBasis_<function>_<degree> tests, :511-780) into
tests/golden/newhall_kats.json: for each of the two test functions and each
degree 3..17, the IsNear intervals "x_(1)" of the error estimate, of the
absolute error of the values (with and without FMA) and of the variations.

  python tools/make_newhall_kats.py [/root/reference]
"""
import json
import os
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
  reference = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
  text = open(os.path.join(reference, 'numerics', 'newhall_test.cpp'),
              encoding='utf-8').read()
  pattern = re.compile(
      r'TEST_F\(NewhallTest, ApproximationInMonomialBasis_(\d)_(\d+)\) \{(.*?)\n\}',
      re.S)
  cases = []
  for function, degree, body in pattern.findall(text):
    def figure(name):
      m = re.search(r'/\*' + name + r'=\*/([-0-9.e]+)_\((\d)\)', body)
      return None if m is None else [m.group(1), int(m.group(2))]
    cases.append(dict(
        function=int(function), degree=int(degree),
        error_estimate=figure('expected_value_error_estimate'),
        value_error=figure('expected_value_absolute_error'),
        variation_error=figure('expected_variation_absolute_error'),
        value_error_with_fma=figure('expected_value_absolute_error_with_fma')))
  assert len(cases) == 30, len(cases)
  out = os.path.join(REPO, 'tests', 'golden', 'newhall_kats.json')
  with open(out, 'w') as f:
    json.dump(dict(source='numerics/newhall_test.cpp:511-780', cases=cases), f,
              indent=1)
  print('wrote', out, len(cases), 'cases')


if __name__ == '__main__':
  main()
