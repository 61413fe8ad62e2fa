#!/usr/bin/env python
"""Tuning aid: how long NVML queries take, and how much a background sampler perturbs the step."""
import os, sys, time, threading, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
for name, fn in [("clock", lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                 ("reasons", lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)),
                 ("power", lambda: pynvml.nvmlDeviceGetPowerUsage(h))]:
    ts = []
    for _ in range(20):
        t = time.perf_counter(); fn(); ts.append(1e3 * (time.perf_counter() - t))
    print(name, "ms min/med/max", round(min(ts), 3), round(sorted(ts)[10], 3), round(max(ts), 3), flush=True)
for mode, q in (("off", "crp"), ("inline", "crp"), ("thread", "c"), ("thread", "r"), ("thread", "p"), ("off", "crp"),
                ("inline", "crp")):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3",
                          "--no-e2e", "--no-cpu-baseline", "--clock-mode", mode, "--clock-queries", q],
                         capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print("mode", mode, q, "ms_per_step", round(d["ms_per_step"], 2), "wall", d["step_wall_ms"], d["clocks"], flush=True)
    except Exception as e:
        print("mode", mode, q, "failed", e, out.stderr[-500:])
