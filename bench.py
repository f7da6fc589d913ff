#!/usr/bin/env python3
"""Headline benchmark: BASELINE.json config 2 -- 10^5 massless LEO probes per GPU
in the 34-body Sol field with the Earth's degree/order-10 geopotential,
fixed-step BlanesMoan2002SRKN14A at 10 s (Ephemeris::FlowWithFixedStep).

One bench "step" = one FlowWithFixedStep call advancing every probe of the rank
by FLOW_STEPS integrator steps (14 right-hand-side evaluations each).

  python bench.py --gpus N --steps K --warmup W          (this engine)
  python bench.py --impl reference ...                   (CPU arm, see below)

`value`: probe-RKN-stages/s, whole job, state resident in HBM.
`e2e`:   the same through the host-buffer C-ABI call (pb_flow_with_fixed_step):
         pinned host state -> HBM -> flow -> host, every step.
`roofline`: FP64 pipe.  achieved = algorithmic flops of one flow-kernel launch
         (principia_b200/flops.py) / its CUDA-event duration; peak = DFMA rate
         measured in this run (MEASURED_PEAKS.json has no FP64 figure).
`cpu_baseline` / `--impl reference`: the reference cannot be built in this
         image, so the CPU arm is the oracle -- the restatement of
         the reference's algorithm -- on all host threads, one integrator
         instance per thread over a bounded sample of the same workload.
"""
import argparse
import gc
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

PROBES_PER_GPU = 100000
FLOW_STEPS = 100          # Integrator steps per bench step.
STEP_SECONDS = 10.0
EVALUATIONS = 14          # BlanesMoan2002SRKN14A.
METRIC = 'massless_probe_rkn_stages_per_second'
UNIT = 'probe-stages/s'
WORKLOAD = ('config 2: 10^5 LEO probes per GPU, Sol 34 bodies, Earth '
            'geopotential 10x10, BlanesMoan2002SRKN14A h=10 s')


def bench_config(probes):
  """The `config` of both arms (the CPU arm flows a bounded sample of it per
  step, stated in its cpu_baseline.sample)."""
  return {'workload': WORKLOAD, 'probes_per_gpu': probes,
          'integrator_steps_per_step': FLOW_STEPS,
          'step_seconds': STEP_SECONDS,
          'parallelism': 'probes sharded by rank, no collective',
          'l2': 'flushed (256 MiB write) before every timed step'}


class ClockSampler:
  """SM clock and throttle reasons during the timed region
  (B200_PROFILING.md).  Two instruments:
   * the SM clock of EVERY timed step is measured on the device
     (pb_sm_clock_probe_start / _read: one warp reads the cycle counter against
     the global timer for 1 ms, 8 ms into the step; enqueued by the bench's own
     thread right before the step's launches) -- no driver query, no second
     thread, no synchronisation inside the step;
   * the throttle reasons (and NVML's own clocks.sm) come from in-process NVML
     (pynvml) queries that BRACKET the region under the same load: one while
     the last warm-up step runs, one while an extra, untimed step runs right
     after the timed region.  None inside it: an NVML query can hold a driver
     lock for tens of milliseconds on some boxes (measured here: 3 ms
     typically, 40-100 ms in one run out of six, every CUDA call of the step
     queued behind it), which would be charged to `value` as if it were flow
     time.  A throttle event inside the region would show in the per-step
     device clocks (sm_min_mhz).
  For an NVML sample the bench kicks the sampler thread when the step starts;
  the thread waits for the step's launches to be enqueued (5 ms) and queries
  while the flow kernels of that step run (45 ms).  nvidia-smi remains the
  fallback."""

  def __init__(self, index):
    self.index = index
    self.samples = []  # (sm MHz, max MHz, reasons bitmask), NVML.
    self.device_mhz = []  # Device-measured SM clock, one per kicked step.
    self.thread = None
    self.stop_event = threading.Event()
    self.kick_event = threading.Event()
    self.query_requested = False
    self.sm_max = None
    self.process = None
    self.lines = []

  def kick(self, query=True):
    """A step starts now: query NVML while it runs (sampler thread)."""
    self.query_requested = query
    self.kick_event.set()

  def probe_start(self):
    """A TIMED step starts now: enqueue, from this thread, the device-side
    clock measurement: 1 ms of cycles against the global timer, starting 8 ms
    from now -- in the middle of the flow kernels that the step launches right
    after."""
    from principia_b200.ephemeris import sm_clock_probe_start
    sm_clock_probe_start(self.index, 8000.0, 1000.0)

  def probe_read(self):
    """The step is complete: collect the measurement (long finished)."""
    from principia_b200.ephemeris import sm_clock_probe_read
    self.device_mhz.append(sm_clock_probe_read(self.index))

  def _query(self, nvml, handle):
    try:
      sm = nvml.nvmlDeviceGetClockInfo(handle, nvml.NVML_CLOCK_SM)
      try:
        reasons = nvml.nvmlDeviceGetCurrentClocksEventReasons(handle)
      except AttributeError:
        reasons = nvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
      self.samples.append((float(sm), float(self.sm_max), int(reasons)))
    except Exception:  # An NVML hiccup must not end the bench.
      pass

  def _nvml_loop(self, nvml, handle):
    while not self.stop_event.is_set():
      if not self.kick_event.wait(0.2):
        continue
      self.kick_event.clear()
      query = self.query_requested
      time.sleep(0.005)
      if query:
        self._query(nvml, handle)

  def start(self):
    try:
      import pynvml as nvml
      nvml.nvmlInit()
      visible = os.environ.get('CUDA_VISIBLE_DEVICES')
      index = self.index
      if visible:
        entries = [v for v in visible.split(',') if v.strip()]
        if self.index < len(entries) and entries[self.index].strip().isdigit():
          index = int(entries[self.index])
      handle = nvml.nvmlDeviceGetHandleByIndex(index)
      self.nvml = nvml
      self.handle = handle
      self.sm_max = nvml.nvmlDeviceGetMaxClockInfo(handle, nvml.NVML_CLOCK_SM)
      self._query(nvml, handle)  # Fails here rather than in the thread.
      if not self.samples:
        raise RuntimeError('NVML query failed')
      self.samples.clear()
      self.thread = threading.Thread(target=self._nvml_loop,
                                     args=(nvml, handle), daemon=True)
      self.thread.start()
      return
    except Exception:
      self.thread = None
    try:
      self.process = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.index),
           '--query-gpu=clocks.sm,clocks.max.sm,power.draw,'
           'clocks_event_reasons.hw_slowdown,'
           'clocks_event_reasons.hw_thermal_slowdown,'
           'clocks_event_reasons.sw_thermal_slowdown,'
           'clocks_event_reasons.sw_power_cap',
           '--format=csv,noheader,nounits', '-lms', '200'],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
      deadline = time.perf_counter() + 5.0
      while not self.lines and time.perf_counter() < deadline:
        time.sleep(0.01)
    except OSError:
      self.process = None

  def _read(self):
    for line in self.process.stdout:
      self.lines.append(line.strip())

  def stop(self):
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
             'sw_power_cap']
    if self.process is None and self.thread is not None:
      # NVML path.
      self.stop_event.set()
      self.thread.join(timeout=2)
      nvml = self.nvml
      bits = {
          'hw_slowdown': nvml.nvmlClocksEventReasonHwSlowdown,
          'hw_thermal_slowdown': nvml.nvmlClocksEventReasonHwThermalSlowdown,
          'sw_thermal_slowdown': nvml.nvmlClocksEventReasonSwThermalSlowdown,
          'sw_power_cap': nvml.nvmlClocksEventReasonSwPowerCap,
      }
      reasons = set()
      for _, _, mask in self.samples:
        for name in names:
          if mask & bits[name]:
            reasons.add(name)
      sm = [s[0] for s in self.samples]
      sm_max = [s[1] for s in self.samples]
      device = self.device_mhz
      return {'sm_mhz': float(np.median(device)) if device else (
                  float(np.median(sm)) if sm else None),
              'sm_min_mhz': float(np.min(device)) if device else None,
              'sm_max_mhz': max(sm_max) if sm_max else None,
              'samples': len(device), 'reasons': sorted(reasons),
              'nvml_sm_mhz': float(np.median(sm)) if sm else None,
              'nvml_samples': len(sm),
              'source': 'SM clock of every timed step measured on the device '
                        'for 1 ms, 8 ms into the step (pb_sm_clock_probe_*); throttle '
                        'reasons and '
                        'nvml_sm_mhz from NVML queries that bracket the timed '
                        'region under the same load (during the last warm-up '
                        'step and during an untimed step right after it)'}
    if self.process is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
    time.sleep(0.25)
    self.process.terminate()
    self.thread.join(timeout=2)
    sm, sm_max, reasons = [], [], set()
    for line in self.lines:
      parts = [p.strip() for p in line.split(',')]
      if len(parts) < 7:
        continue
      try:
        sm.append(float(parts[0]))
        sm_max.append(float(parts[1]))
      except ValueError:
        continue
      for name, value in zip(names, parts[3:7]):
        if value.lower().startswith('active'):
          reasons.add(name)
    return {'sm_mhz': float(np.median(sm)) if sm else None,
            'sm_max_mhz': max(sm_max) if sm_max else None,
            'samples': len(sm), 'reasons': sorted(reasons),
            'source': 'nvidia-smi'}


def host_cores():
  return len(os.sched_getaffinity(0))


def profile_counters(name):
  """profiles/current_<name>.json (tools/ncu_to_json.py): the ncu-derived
  figures of the roofline entry, with the capture they come from."""
  path = os.path.join(REPO, 'profiles', 'current_%s.json' % name)
  try:
    with open(path) as f:
      data = json.load(f)
  except (OSError, ValueError):
    return {'traffic': None, 'ncu': None}
  return {'traffic': data.get('dram_bytes_per_launch'),
          'ncu': {k: data.get(k) for k in (
              'source', 'kernel', 'launch', 'duration_ms',
              'fp64_pipe_active_pct', 'issue_active_pct',
              'warps_active_pct', 'registers_per_thread',
              'fp64_instructions_per_probe_stage',
              'instructions_per_probe_stage',
              'frac_at_full_fp64_pipe')}}


def make_workload(n, seed):
  from principia_b200 import workloads
  system = workloads.sol_system()
  earth = system.bodies[system.index('Earth')]
  state = workloads.leo_probes(n, earth['q'], earth['v'], earth['mu'], seed)
  return system, state


def oracle_with_golden_segments(system):
  sys.path.insert(0, os.path.join(REPO, 'tests'))
  from oracle_lib import OracleEphemeris
  from principia_b200 import workloads
  oracle = OracleEphemeris(system)
  for body, segments in enumerate(workloads.load_sol_ephemeris_segments()):
    oracle.append_segments(body, segments)
  return oracle


def cpu_flow(oracle, state, flow_steps, threads):
  """Times one CPU FlowWithFixedStep over `state`; returns seconds."""
  from principia_b200 import _capi
  time_dp = np.zeros(2)
  t0 = time.perf_counter()
  oracle.flow_with_fixed_step(_capi.BLANES_MOAN_2002_SRKN_14A, STEP_SECONDS,
                              flow_steps * STEP_SECONDS, state, time_dp,
                              threads=threads)
  return time.perf_counter() - t0


def cpu_sample_size(oracle, system, threads, target_seconds):
  """Probes per CPU step such that one step costs about target_seconds."""
  _, probe = make_workload(threads * 4, seed=7)
  seconds = cpu_flow(oracle, probe, 10, threads)
  per_probe_stage = seconds / (threads * 4 * 10 * EVALUATIONS)
  n = int(target_seconds / (per_probe_stage * FLOW_STEPS * EVALUATIONS))
  return max(threads, (n // threads) * threads)


def run_reference(args):
  """CPU arm: the oracle on all host threads, bounded sample per step."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  threads = host_cores()
  system, _ = make_workload(8, 42)
  oracle = oracle_with_golden_segments(system)
  n = cpu_sample_size(oracle, system, threads, target_seconds=8.0)
  _, state0 = make_workload(n, 42)
  for _ in range(min(args.warmup, 1)):
    cpu_flow(oracle, state0[:, :threads * 2].copy(), 10, threads)
  total = 0.0
  for _ in range(args.steps):
    total += cpu_flow(oracle, state0.copy(), FLOW_STEPS, threads)
  stages = args.steps * n * FLOW_STEPS * EVALUATIONS
  value = stages / total
  sample = ('%d probes x %d SRKN14A steps per bench step (bounded sample of the '
            '10^5-probe workload), oracle restatement, OpenMP %d threads, one '
            'integrator instance per thread' % (n, FLOW_STEPS, threads))
  print(json.dumps({
      'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
      'data': 'synthetic',
      'config': bench_config(args.probes),
      'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads,
                       'kind': 'port', 'sample': sample},
      'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0}}))


def run_native(args):
  import torch
  import torch.distributed as dist
  from principia_b200 import _capi, flops, workloads
  from principia_b200.ephemeris import (Ephemeris, kernel_launch_count,
                                        measure_fp64_peak)

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if not torch.cuda.is_available():
    raise SystemExit('bench.py: no CUDA device; there is no CPU fallback '
                     '(use --impl reference for the CPU arm)')
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  if world > 1:
    dist.init_process_group('nccl', device_id=device)

  n = args.probes
  system, state_host = make_workload(n, seed=42 + rank)
  ephemeris = Ephemeris(system, device=local_rank)
  # Ephemeris::Prolong by the product itself (massive-body flow and Newhall
  # fit on the device), over the two days the flows below span.
  workloads.prolong_ephemeris(ephemeris, 2 * 86400.0)
  method = _capi.BLANES_MOAN_2002_SRKN_14A
  span = FLOW_STEPS * STEP_SECONDS

  state_device = torch.from_numpy(state_host).to(device)
  pinned = torch.from_numpy(state_host.copy()).pin_memory()
  pinned_np = pinned.numpy()
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
  # A stream of our own for the flush, the timing events and the resident
  # flows: on the legacy default stream (torch's current stream at start-up)
  # one run in four had steps of 60-100 ms around kernels of 44 ms -- that
  # stream synchronises implicitly with every blocking stream of the process.
  torch.cuda.set_stream(torch.cuda.Stream(device=device))
  stream = torch.cuda.current_stream().cuda_stream

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # The ephemeris covers two days = 172 bench steps of 1000 s; the
  # flows advance continuously in time and restart from the initial state
  # after WRAP steps (a device/host copy of the pristine state, negligible).
  WRAP = 150
  pristine_device = state_device.clone()
  counters = {'resident': 0, 'e2e': 0}

  def resident_step(arithmetic=_capi.PB_ARITHMETIC_EXACT):
    k = counters['resident'] % WRAP
    if k == 0 and counters['resident'] > 0:
      state_device.copy_(pristine_device)
    counters['resident'] += 1
    time_dp = np.array([k * span, 0.0])
    result = ephemeris.flow_with_fixed_step_device(
        method, STEP_SECONDS, (k + 1) * span, n, state_device.data_ptr(),
        time_dp, stream, arithmetic=arithmetic)
    assert result['n_steps'] == FLOW_STEPS and result['status'] == 0, result
    return result

  def e2e_step():
    k = counters['e2e'] % WRAP
    if k == 0 and counters['e2e'] > 0:
      pinned_np[...] = state_host
    counters['e2e'] += 1
    time_dp = np.array([k * span, 0.0])
    result = ephemeris.flow_with_fixed_step(method, STEP_SECONDS,
                                            (k + 1) * span, pinned_np, time_dp)
    assert result['n_steps'] == FLOW_STEPS and result['status'] == 0, result
    return float(pinned_np[0, 0])  # The step's result, read on the host.

  # FP64 roofline denominator, measured on this device, before the timed runs.
  peak_flops, peak_ms = measure_fp64_peak(local_rank)

  sampler = ClockSampler(local_rank)
  sampler.start()
  for k in range(args.warmup):
    if k == args.warmup - 1:
      sampler.kick()
    # (The flush belongs to the warm-up too: the first launch of torch's fill
    # kernel loads its module, 8 ms with a warm file cache and up to 0.8 s on a
    # fresh box, which used to land in the first timed step.)
    flush.fill_(k & 0xff)
    resident_step()
  launches_before = kernel_launch_count()
  barrier()
  start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
  kernel_ms = 0.0
  kernel_launches = 0
  # No cyclic garbage collection inside the timed regions: with torch loaded a
  # full collection is a ~100 ms pause of the host thread, which showed up as
  # one slow run in five.
  gc.collect()
  gc.disable()
  step_wall_ms = []  # Host wall clock of each timed flow call (diagnostic).
  start.record()
  for k in range(args.steps):
    flush.fill_(k & 0xff)  # Evict L2 between steps.
    sampler.probe_start()
    t0 = time.perf_counter()
    resident_step()
    step_wall_ms.append((time.perf_counter() - t0) * 1e3)
    sampler.probe_read()
    # Device time (CUDA events on the launching stream) of the flow kernels of
    # this call: the call slices the batch into launches that overlap on
    # helper streams, so the figure is per call, not per launch.
    ms, launches = ephemeris.last_flow_kernel_ms()
    kernel_ms += ms
    kernel_launches += launches
  stop.record()
  barrier()
  launches_timed = kernel_launch_count() - launches_before
  # The closing NVML sample, under the same load, outside the timed region.
  sampler.kick()
  resident_step()
  clocks = sampler.stop()
  elapsed_ms = torch.tensor([start.elapsed_time(stop)], device=device,
                            dtype=torch.float64)
  if world > 1:
    dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
  elapsed_ms = float(elapsed_ms.item())

  # End to end through the host-buffer entry point.
  for k in range(min(args.warmup, 3)):
    e2e_step()
  barrier()
  start.record()
  for k in range(args.steps):
    flush.fill_(k & 0xff)
    e2e_step()
  stop.record()
  barrier()
  e2e_ms = torch.tensor([start.elapsed_time(stop)], device=device,
                        dtype=torch.float64)
  if world > 1:
    dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
  e2e_ms = float(e2e_ms.item())

  # End to end WITH the trajectory sink: FlowWithFixedStep appends every step
  # to a DiscreteTrajectory (physics/ephemeris_body.hpp:1118-1131), here the
  # device-side DiscreteTrajectorySegment with the plugin's downsampling
  # (10 m): Ephemeris::NewInstance from the host state, FlowWithFixedStep into
  # the sink, the final state and the length of every timeline back on the
  # host, every step.
  from principia_b200.trajectory import Downsampler
  sink_steps = min(args.steps, 10)
  sink = Downsampler(n, 10.0, (sink_steps + 2) * FLOW_STEPS + 2,
                     device=local_rank)
  sink.append([0.0], np.concatenate([state_host[0:3], state_host[6:9]])[None])
  sink_state = pinned_np
  sink_state[...] = state_host
  sink_time = np.zeros(2)
  sink_points = 0

  def sink_step(k):
    instance = ephemeris.new_instance(method, STEP_SECONDS, sink_state,
                                      sink_time)
    result = instance.flow_to_downsampler((k + 1) * span, sink)
    assert result['n_steps'] == FLOW_STEPS and result['status'] == 0, result
    instance.state_into(sink_state, sink_time)
    instance.close()
    return int(sink.counts().sum())

  sink_step(0)
  barrier()
  sink_step_ms = []
  start.record()
  for k in range(1, sink_steps + 1):
    flush.fill_(k & 0xff)
    t0 = time.perf_counter()
    sink_points = sink_step(k)  # Ends with a device-to-host read.
    sink_step_ms.append((time.perf_counter() - t0) * 1e3)
  stop.record()
  barrier()
  sink_ms = torch.tensor([start.elapsed_time(stop)], device=device,
                         dtype=torch.float64)
  if world > 1:
    dist.all_reduce(sink_ms, op=dist.ReduceOp.MAX)
  sink_ms = float(sink_ms.item())
  del sink

  stages_per_step = world * n * FLOW_STEPS * EVALUATIONS
  value = args.steps * stages_per_step / (elapsed_ms * 1e-3)
  e2e_value = args.steps * stages_per_step / (e2e_ms * 1e-3)

  call_flops = n * FLOW_STEPS * flops.srkn14a_probe_step_sol_leo()
  average_kernel_ms = kernel_ms / args.steps
  achieved = call_flops / (average_kernel_ms * 1e-3) / 1e12
  roofline = {
      'bound': 'fp64', 'achieved': achieved, 'peak': peak_flops / 1e12,
      'unit': 'TFLOP/s', 'frac': achieved / (peak_flops / 1e12),
      # Counters that only a profiler sees (DRAM traffic per launch, FP64-pipe
      # activity, executed FP64 instructions per probe-stage) are read from the
      # committed summary that tools/ncu_to_json.py wrote from the ncu capture
      # of this kernel -- never literals here; absent file: null.
      **profile_counters('srkn_flow'),
      'kernel': 'SrknFlowKernel',
      'kernel_ms': average_kernel_ms,
      'kernel_launches_per_step': kernel_launches / args.steps,
      'kernel_share_of_step': kernel_ms / elapsed_ms,
      'algorithmic_flop_per_probe_stage':
          flops.massless_probe_stage_sol_leo(),
      'peak_source': 'register-resident DFMA chain measured in this run '
                     '(pb_measure_fp64_peak, %.2f ms); MEASURED_PEAKS.json '
                     'has no FP64 figure; nominal 37.2 TFLOP/s' % peak_ms}

  # The opt-in contracted arithmetic (fused multiply-adds; within 1e-12 of the
  # reference, tests/test_gpu_contracted.py): the same workload, resident.  The
  # headline (`value`, `e2e`, `roofline`) stays the exact arithmetic.
  for k in range(2):
    resident_step(_capi.PB_ARITHMETIC_CONTRACTED)
  barrier()
  contracted_kernel_ms = 0.0
  start.record()
  for k in range(args.steps):
    flush.fill_(k & 0xff)
    resident_step(_capi.PB_ARITHMETIC_CONTRACTED)
    contracted_kernel_ms += ephemeris.last_flow_kernel_ms()[0]
  stop.record()
  barrier()
  contracted_ms = torch.tensor([start.elapsed_time(stop)], device=device,
                               dtype=torch.float64)
  if world > 1:
    dist.all_reduce(contracted_ms, op=dist.ReduceOp.MAX)
  contracted_ms = float(contracted_ms.item())
  contracted_achieved = (call_flops /
                         (contracted_kernel_ms / args.steps * 1e-3) / 1e12)
  roofline_contracted = {
      'arithmetic': 'PB_ARITHMETIC_CONTRACTED (opt-in; not bit-identical, '
                    'relative position error <= 1e-12 after 10^4 steps)',
      'value': args.steps * stages_per_step / (contracted_ms * 1e-3),
      'unit': UNIT, 'ms_per_step': contracted_ms / args.steps,
      'bound': 'fp64', 'achieved': contracted_achieved,
      'peak': peak_flops / 1e12, 'unit_achieved': 'TFLOP/s',
      'frac': contracted_achieved / (peak_flops / 1e12),
      'kernel': 'SrknFlowKernel (pb_contracted.cu, --fmad=true)',
      'kernel_ms': contracted_kernel_ms / args.steps,
      **profile_counters('srkn_flow_contracted')}

  cpu_baseline = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    threads = host_cores()
    oracle = oracle_with_golden_segments(system)
    n_cpu = cpu_sample_size(oracle, system, threads, target_seconds=12.0)
    _, sample_state = make_workload(n_cpu, 42)
    seconds = cpu_flow(oracle, sample_state, FLOW_STEPS, threads)
    cpu_baseline = {
        'value': n_cpu * FLOW_STEPS * EVALUATIONS / seconds, 'unit': UNIT,
        'cores': threads, 'kind': 'port',
        'sample': '%d of the 10^5 probes x %d steps, %.1f s of CPU work, '
                  'oracle restatement on %d OpenMP threads' %
                  (n_cpu, FLOW_STEPS, seconds, threads)}

  secondary = None
  if not args.no_secondary:
    import bench_secondary
    secondary = bench_secondary.run_secondary(torch, dist, world, rank,
                                              local_rank, peak_flops, args)

  if rank == 0:
    print(json.dumps({
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': bench_config(n),
        'probe_steps_per_second': value / EVALUATIONS,
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT,
                'h2d_bytes_per_step': world * 12 * n * 8,
                'd2h_bytes_per_step': world * 12 * n * 8,
                'ms_per_step': e2e_ms / args.steps},
        'e2e_with_sink': {
            'value': sink_steps * stages_per_step / (sink_ms * 1e-3),
            'unit': UNIT, 'ms_per_step': sink_ms / sink_steps,
            # Host wall clock of the individual steps on this rank (every
            # step creates and destroys an integrator instance).
            'ms_per_step_median': float(np.median(sink_step_ms)),
            'ms_per_step_max': float(np.max(sink_step_ms)),
            'h2d_bytes_per_step': world * 12 * n * 8,
            'd2h_bytes_per_step': world * (12 * n * 8 + n * 8),
            'sink': 'every step of every probe appended to a device '
                    'DiscreteTrajectorySegment with 10 m downsampling '
                    '(pb_fixed_step_instance_flow_to_downsampler)',
            'points_kept_per_probe':
                sink_points / n if sink_points else None,
            'points_appended_per_probe': (sink_steps + 1) * FLOW_STEPS + 1},
        'step_wall_ms': [round(x, 2) for x in step_wall_ms],
        'gpu_launches': int(launches_timed),
        'roofline': roofline,
        'roofline_contracted': roofline_contracted,
        'cpu_baseline': cpu_baseline,
        'secondary': secondary}))
  if world > 1:
    dist.destroy_process_group()


def main():
  parser = argparse.ArgumentParser()
  parser.add_argument('--gpus', type=int, default=1)
  parser.add_argument('--steps', type=int, default=10)
  parser.add_argument('--warmup', type=int, default=3)
  parser.add_argument('--impl', default='native',
                      choices=['native', 'reference'])
  parser.add_argument('--probes', type=int, default=PROBES_PER_GPU)
  parser.add_argument('--no-cpu-baseline', action='store_true')
  parser.add_argument('--no-secondary', action='store_true',
                      help='skip configs 3, 4 and 5 (the `secondary` block)')
  parser.add_argument('--adaptive-probes', type=int, default=100000)
  parser.add_argument('--ensemble-systems', type=int, default=100000)
  parser.add_argument('--allpairs-bodies', type=int, default=2**20)
  parser.add_argument('--allpairs-flow-bodies', type=int, default=2**17)
  args = parser.parse_args()
  # The timing rules ask for at least three warm-up steps (module loads, the
  # choice of the hot bodies, clocks): fewer are not accepted silently.
  if args.impl == 'native' and args.warmup < 3:
    print('bench.py: --warmup %d raised to 3' % args.warmup, file=sys.stderr)
    args.warmup = 3
  if args.steps < 1:
    raise SystemExit('bench.py: --steps must be at least 1')
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_native(args)


if __name__ == '__main__':
  main()
