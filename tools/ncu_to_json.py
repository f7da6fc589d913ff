#!/usr/bin/env python3
"""Writes profiles/current_<name>.json from one kernel of an .ncu-rep: the
profiler-only figures that bench.py attaches to its roofline entries (DRAM
traffic per launch, FP64-pipe activity, executed instructions per unit of
work), so that the bench line never carries literals that could go stale.

  python tools/ncu_to_json.py gpurun_out/x.ncu-rep NAME --units U \
      [--unit-name probe_stage] [--flop-per-unit F] [--launch-index K]

`--units`: units of work (probe-stages, interactions ...) of the profiled
launch; `--flop-per-unit`: algorithmic flops of one unit
(principia_b200/flops.py), from which frac_at_full_fp64_pipe = F / (2 x FP64
instructions per unit): the roofline fraction the executed instruction mix
would reach with the FP64 pipe 100 % busy.
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ncu(args):
  return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def number(text):
  try:
    return float(text.replace(',', ''))
  except ValueError:
    return None


SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9,
         'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}


def main():
  parser = argparse.ArgumentParser()
  parser.add_argument('report')
  parser.add_argument('name')
  parser.add_argument('--units', type=float, required=True)
  parser.add_argument('--unit-name', default='probe_stage')
  parser.add_argument('--flop-per-unit', type=float, default=None)
  parser.add_argument('--launch-index', type=int, default=0)
  args = parser.parse_args()

  rows = list(csv.reader(io.StringIO(
      ncu(['-i', args.report, '--page', 'raw', '--csv']))))
  header, units, values = rows[0], rows[1], rows[2 + args.launch_index]
  raw = {h: (u, v) for h, u, v in zip(header, units, values)}

  def metric(name):
    unit, value = raw.get(name, ('', ''))
    x = number(value)
    return None if x is None else x * SCALE.get(unit, 1.0)

  # Executed warp-level instructions by pipe, from the SASS page.
  text = ncu(['-i', args.report, '--page', 'source', '--csv',
              '--print-source', 'sass'])
  executed = fp64 = 0.0
  by_opcode = {}
  kernel_index = -1
  for row in csv.reader(io.StringIO(text)):
    if not row:
      continue
    if row[0] == 'Kernel Name':
      kernel_index += 1
      continue
    if kernel_index != args.launch_index or not row[0].startswith('0x'):
      continue
    match = re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', row[1])
    count = number(row[5]) or 0.0
    executed += count
    if match:
      opcode = match.group(1)
      by_opcode[opcode] = by_opcode.get(opcode, 0.0) + count
      if opcode in ('DADD', 'DMUL', 'DFMA', 'DSETP', 'DMNMX'):
        fp64 += count
  lanes = 32.0
  per_unit = executed * lanes / args.units
  fp64_per_unit = fp64 * lanes / args.units
  out = {
      'source': os.path.basename(args.report),
      'kernel': raw.get('Kernel Name', ('', ''))[1],
      'launch': '%s x %s' % (raw.get('Grid Size', ('', ''))[1],
                             raw.get('Block Size', ('', ''))[1]),
      'units_in_launch': args.units, 'unit': args.unit_name,
      'duration_ms': metric('gpu__time_duration.sum'),
      'dram_bytes_per_launch': (metric('dram__bytes_read.sum') or 0.0) +
                               (metric('dram__bytes_write.sum') or 0.0),
      'fp64_pipe_active_pct': metric(
          'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'),
      'issue_active_pct': metric(
          'smsp__issue_active.avg.pct_of_peak_sustained_active'),
      'warps_active_pct': metric(
          'sm__warps_active.avg.pct_of_peak_sustained_active'),
      'registers_per_thread': metric('launch__registers_per_thread'),
      'instructions_per_%s' % args.unit_name: per_unit,
      'fp64_instructions_per_%s' % args.unit_name: fp64_per_unit,
      'fp64_opcodes_per_%s' % args.unit_name: {
          k: v * lanes / args.units for k, v in sorted(by_opcode.items())
          if k in ('DADD', 'DMUL', 'DFMA', 'DSETP', 'DMNMX')},
      'stall_per_issue': {
          k.split('issue_stalled_')[1].split('_per_issue')[0]: number(v[1])
          for k, v in raw.items()
          if 'issue_stalled' in k and k.endswith('_per_issue_active.ratio')
          and (number(v[1]) or 0) >= 0.05},
  }
  if args.flop_per_unit is not None and fp64_per_unit > 0:
    out['algorithmic_flop_per_%s' % args.unit_name] = args.flop_per_unit
    out['frac_at_full_fp64_pipe'] = args.flop_per_unit / (2.0 * fp64_per_unit)
  path = os.path.join(REPO, 'profiles', 'current_%s.json' % args.name)
  with open(path, 'w') as f:
    json.dump(out, f, indent=1, sort_keys=True)
    f.write('\n')
  print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == '__main__':
  main()
