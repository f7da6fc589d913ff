#!/usr/bin/env python3
"""How long do the NVML queries of bench.py's ClockSampler take on this box,
and do they delay CUDA calls issued meanwhile?  (One run of bench.py measured
120 ms per step with the sampler against 49 ms without.)"""
import threading
import time

import pynvml as nvml
import torch


def main():
  nvml.nvmlInit()
  handle = nvml.nvmlDeviceGetHandleByIndex(0)
  calls = {
      'clock': lambda: nvml.nvmlDeviceGetClockInfo(handle, nvml.NVML_CLOCK_SM),
      'max_clock': lambda: nvml.nvmlDeviceGetMaxClockInfo(handle,
                                                          nvml.NVML_CLOCK_SM),
      'reasons': lambda: nvml.nvmlDeviceGetCurrentClocksEventReasons(handle),
      'power': lambda: nvml.nvmlDeviceGetPowerUsage(handle),
  }
  x = torch.zeros(1 << 20, device='cuda')
  torch.cuda.synchronize()
  for name, call in calls.items():
    durations = []
    for _ in range(10):
      t0 = time.perf_counter()
      call()
      durations.append(time.perf_counter() - t0)
    print('%s: median %.3f ms max %.3f ms' %
          (name, 1e3 * sorted(durations)[5], 1e3 * max(durations)))
  # CUDA call latency with and without a concurrent query loop.
  def cuda_loop(label):
    worst = total = 0.0
    for _ in range(300):
      t0 = time.perf_counter()
      x.add_(1.0)
      torch.cuda.synchronize()
      d = time.perf_counter() - t0
      worst = max(worst, d)
      total += d
    print('%s: 300 launch+sync: mean %.3f ms worst %.3f ms' %
          (label, 1e3 * total / 300, 1e3 * worst))
  cuda_loop('quiet')
  for name, call in calls.items():
    stop = threading.Event()
    def loop():
      while not stop.is_set():
        call()
        stop.wait(0.02)
    thread = threading.Thread(target=loop, daemon=True)
    thread.start()
    cuda_loop('with %s every 20 ms' % name)
    stop.set()
    thread.join()


if __name__ == '__main__':
  main()
