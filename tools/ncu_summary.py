#!/usr/bin/env python3
"""Summarises an .ncu-rep: headline counters of the first kernel and the
source lines with the most stall samples.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep [top_lines]
"""
import csv
import io
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'launch__registers_per_thread',
    'launch__grid_size', 'launch__block_size',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
    'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
    'l1tex__data_pipe_lsu_wavefronts.sum',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
]


def run(args):
  return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
  rep = sys.argv[1]
  top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
  rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
  header, units, values = rows[0], rows[1], rows[2]
  for h, u, v in zip(header, units, values):
    if h in WANT or ('issue_stalled' in h and h.endswith('_per_warp_active.pct')
                     and 'not_issued' not in h):
      print(f'{h} [{u}] {v}')
  text = run(['-i', rep, '--page', 'source', '--csv', '--print-source',
              'cuda,sass'])
  current = None
  lines = {}
  total_samples = total_inst = 0
  for row in csv.reader(io.StringIO(text)):
    if not row:
      continue
    if row[0] == 'File Path':
      current = row[1].split('/')[-1]
      continue
    if row[0] in ('Function Name', 'Line No') or row[0] == '':
      continue
    try:
      line = int(row[0])
    except ValueError:
      continue
    def number(text):
      try:
        return int(text)
      except ValueError:
        return 0
    samples, inst, thread_inst = number(row[6]), number(row[7]), number(row[8])
    lines[(current, line)] = (samples, inst, thread_inst, row[1].strip()[:72])
    total_samples += samples
    total_inst += inst
  print(f'warp instructions {total_inst}, samples {total_samples}')
  by_file = {}
  for (f, _), (s, i, t, _) in lines.items():
    a = by_file.setdefault(f, [0, 0, 0])
    a[0] += s
    a[1] += i
    a[2] += t
  for f, (s, i, t) in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
    print(f'  {f}: samples {100 * s / max(total_samples, 1):.1f}% '
          f'inst {100 * i / max(total_inst, 1):.1f}% lanes {t / max(i, 1):.1f}')
  for (f, l), (s, i, t, src) in sorted(lines.items(),
                                       key=lambda kv: -kv[1][0])[:top]:
    print(f'{f}:{l:4d} samp {100 * s / max(total_samples, 1):5.1f}% '
          f'inst {100 * i / max(total_inst, 1):5.1f}% lanes {t / max(i, 1):4.1f}  {src}')


if __name__ == '__main__':
  main()
