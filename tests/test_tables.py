"""The generated tables (tools/gen_tables.py, derived from first principles) must
equal, bit for bit, the tables the reference ships as Mathematica output.

Only runs where /root/reference exists (this container); on the GPU box the
test is skipped — the generated header travels, the reference does not.
"""
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/numerics'

pytestmark = pytest.mark.skipif(not os.path.isdir(REF),
                                reason='reference tree not present')


def parse_generated(path):
  text = open(path).read()
  arrays = {}
  for m in re.finditer(
      r'static const (double|int) (\w+)\[\d+\] = \{(.*?)\};', text, re.S):
    kind, name, body = m.groups()
    toks = [t.strip() for t in body.replace('\n', ' ').split(',') if t.strip()]
    arrays[name] = [float.fromhex(t) if kind == 'double' else int(t)
                    for t in toks]
  return arrays


def ref_numbers(text):
  text = re.sub(r'/\*.*?\*/', '', text)
  text = text.replace("'", '')
  return [float(t) for t in re.findall(
      r'(?<![\w.])[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?(?![\w.])', text)]


@pytest.mark.parametrize('header', ['oracle/generated_tables.h',
                                    'principia_b200/csrc/generated_tables.h'])
def test_legendre_tables_match_reference(header):
  gen = parse_generated(os.path.join(REPO, header))
  for ref_file, name in (
      ('legendre_normalization_factor.mathematica.h',
       'pb_legendre_normalization_factor'),
      ('max_abs_normalized_associated_legendre_function.mathematica.h',
       'pb_max_abs_normalized_legendre')):
    text = open(os.path.join(REF, ref_file), encoding='utf-8').read()
    body = text[text.index('{{{'):text.index('}}}')]
    ref = ref_numbers(body)
    assert len(ref) == 51 * 52 // 2
    assert gen[name] == ref, name


@pytest.mark.parametrize('header', ['oracle/generated_tables.h',
                                    'principia_b200/csrc/generated_tables.h'])
def test_newhall_matrices_match_reference(header):
  gen = parse_generated(os.path.join(REPO, header))
  text = open(os.path.join(REF, 'newhall_matrices.mathematica.h'),
              encoding='utf-8-sig').read()
  offsets = gen['pb_newhall_monomial_offset']
  err_pos = 0
  for degree in range(3, 18):
    for basis in ('чебышёв', 'monomial'):
      key = 'newhall_c_matrix_%s_degree_%d_divisions_8_w04(' % (basis, degree)
      start = text.index(key)
      start = text.index('{', text.index('std::array', start))
      end = text.index('}}});', start)
      ref = ref_numbers(text[start:end])
      assert len(ref) == (degree + 1) * 18
      if basis == 'monomial':
        off = offsets[degree - 3]
        assert gen['pb_newhall_monomial'][off:off + len(ref)] == ref, degree
      else:
        mine = gen['pb_newhall_chebyshev_last_row'][err_pos:err_pos + 18]
        assert mine == ref[-18:], degree
        err_pos += 18
