#!/usr/bin/env python
"""
bench.py -- the reference's headline metric on its headline config, on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): sorted+counted k-mers per second.  One "step" = one pass of the hot path
(both-strand layout -> key pack -> onesweep radix sort -> ambiguous-window refinement -> run-length
grouping -> group-size histogram) over one synthetic genome.

Workload at N=1 = BASELINE.json configs[1] ("C2"): synthetic 100 Mbp, 10 equal records, 20 N-runs per
record with log-uniform lengths 10^3..10^5, both strands, k=31 -> ~2.0e8 k-mers.
At N>1 every GPU brings its own 100 Mbp of genome (weak scaling): the N x 100 Mbp collection is
sorted and counted as ONE index, key-range partitioned over the ranks with one NCCL all-to-all.

Printed JSON line:
  value / ms_per_step  device-timed (CUDA events on the launching stream), inputs resident in HBM
  e2e                  same metric through the Python API with host buffers: H2D of the byte array from
                       pinned memory and D2H of the sorted start indices + histogram inside the timed region
  roofline             dominant kernel (one onesweep pass): algorithmic 2*W*N bytes / average pass time
  cpu_baseline         the CPU oracle (a port of the reference's algorithm) on a bounded sample
`--impl reference` times that CPU port with all host threads on the same kind of workload.
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genome-kmers_b200"))

METRIC = "sorted+counted k-mers/sec"
UNIT = "Gkmer/s"
K = 31
BASES_PER_GPU = 100_000_000
N_RECORDS = 10
RUNS_PER_RECORD = 20
MAX_BIN = 1_000_000
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md


# ---------------------------------------------------------------------------------------------
# synthetic workload: iid uniform ACGT, equal-length records, N runs
# ---------------------------------------------------------------------------------------------
def make_genome(n_bases, n_records, runs_per_record, seed, out=None):
    """Forward sequence byte array (records joined by '$') + segment starts + names."""
    rng = np.random.default_rng(seed)
    avg = n_bases // n_records
    lengths = [avg] * (n_records - 1) + [n_bases - avg * (n_records - 1)]
    total = n_bases + n_records - 1
    sba = out if out is not None else np.empty(total, dtype=np.uint8)
    assert len(sba) == total
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    starts, pos = [], 0
    for r, ln in enumerate(lengths):
        starts.append(pos)
        chunk = 1 << 24
        for lo in range(0, ln, chunk):
            hi = min(ln, lo + chunk)
            sba[pos + lo:pos + hi] = lut[rng.integers(0, 4, hi - lo, dtype=np.uint8)]
        for _ in range(runs_per_record):
            run = int(np.exp(rng.uniform(np.log(1e3), np.log(1e5))))
            run = min(run, max(1, ln // 4))
            st = int(rng.integers(0, ln - run))
            sba[pos + st:pos + st + run] = ord("N")
        pos += ln
        if r != n_records - 1:
            sba[pos] = ord("$")
            pos += 1
    names = [f"chr{i}" for i in range(n_records)]
    return sba, np.asarray(starts, dtype=np.uint64), names


def n_kmers(n_bases, n_records, k, strands=2):
    return strands * (n_bases - n_records * (k - 1))


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, through NVML.

    mode "inline": sample() is called by the timing loop between steps (no second thread touching the
    driver while kernels are being launched); mode "thread": a background thread polls every period_s.
    `queries` selects what is read ("c" clock, "r" throttle reasons, "p" power)."""

    def __init__(self, gpu_index=0, period_s=0.02, mode="inline", queries="crp"):
        self.gpu_index = gpu_index
        self.period_s = period_s
        self.mode = mode
        self.queries = queries
        self.samples = []
        self.thread = None
        self.stop_flag = False
        self.nvml = None

    def start(self):
        if self.mode == "off":
            return
        try:
            import threading

            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.sample()
            if self.mode == "thread":
                self.thread = threading.Thread(target=self._run, daemon=True)
                self.thread.start()
        except Exception:
            self.nvml = None

    def sample(self):
        nv = self.nvml
        if nv is None:
            return
        try:
            sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM) if "c" in self.queries else 0
            reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if "r" in self.queries else 0
            power = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0 if "p" in self.queries else 0.0
            self.samples.append((sm, reasons, power))
        except Exception:
            pass

    def _run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(self.period_s)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.nvml is None:
            return out
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
        nv = self.nvml
        if self.samples:
            out["sm_mhz"] = float(np.median([s[0] for s in self.samples]))
            out["sm_max_mhz"] = float(self.max_sm)
            out["samples"] = len(self.samples)
            out["power_w_max"] = float(max(s[2] for s in self.samples))
            out["sampler"] = self.mode
            bits = 0
            for s in self.samples:
                bits |= int(s[1])
            names = {"hw_slowdown": "nvmlClocksEventReasonHwSlowdown",
                     "hw_thermal_slowdown": "nvmlClocksEventReasonHwThermalSlowdown",
                     "sw_thermal_slowdown": "nvmlClocksEventReasonSwThermalSlowdown",
                     "sw_power_cap": "nvmlClocksEventReasonSwPowerCap"}
            alt = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown",
                   "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
                   "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown",
                   "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap"}
            for key in names:
                mask = getattr(nv, names[key], None) or getattr(nv, alt[key], 0)
                if bits & int(mask):
                    out["reasons"].append(key)
        return out


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """dram bytes per onesweep launch from the committed ncu capture, if one exists for this workload."""
    path = os.path.join(ROOT, "profiles", "onesweep_traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's algorithm (quicksort + comparator + group walk)
# ---------------------------------------------------------------------------------------------
def cpu_port_run(sample_bases, threads, seed=7):
    import oracle

    n_rec = 2
    sba, starts, _ = make_genome(sample_bases, n_rec, 2, seed)
    both, both_starts = oracle.both_strands(sba, starts)
    init = oracle.init_indices(both_starts, len(both), K)
    t0 = time.perf_counter()
    srt = oracle.sort_indices(both, init, K, K, break_ties=False, validate=True, threads=threads)
    hist, total = oracle.group_hist(both, srt, K, max_bin=MAX_BIN)
    dt = time.perf_counter() - t0
    assert total == len(init)
    return len(init), dt


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    import oracle

    threads = oracle.max_threads()
    sample_bases = args.ref_sample_bases
    for _ in range(args.warmup):
        cpu_port_run(min(sample_bases, 200_000), threads)
    times, n = [], 0
    for s in range(args.steps):
        n, dt = cpu_port_run(sample_bases, threads, seed=100 + s)
        times.append(dt)
    total_t = sum(times)
    value = n * args.steps / total_t / 1e9
    sample = (f"{sample_bases} bp sample of the workload generator (2 records, N runs), both strands, k={K}: "
              f"{n} k-mers per step; C port of the reference's quicksort+comparator+group walk "
              f"(oracle/gk_oracle.c) parallelised over {threads} OpenMP threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {
        "workload": ("C2 (BASELINE.json configs[1]): synthetic %d Mbp per GPU, %d records per GPU, N runs, "
                     "forward + reverse-complement strands, k=%d sort + unique counts"
                     % (BASES_PER_GPU // 1_000_000, N_RECORDS, K)),
        "k": K, "bases_per_gpu": BASES_PER_GPU, "records_per_gpu": N_RECORDS, "strands": "both",
        "kmers_total": n_kmers(BASES_PER_GPU * n_gpus, N_RECORDS * n_gpus, K),
        "max_counts_bin": MAX_BIN,
        "parallelism": "single GPU" if n_gpus == 1 else f"key-range sharded x{n_gpus}, one NCCL all-to-all",
        "l2_policy": "inputs larger than L2 (>= 2.4 GB of key/index pairs per pass vs 126 MB L2)",
    }


# ---------------------------------------------------------------------------------------------
# GPU arm, one GPU
# ---------------------------------------------------------------------------------------------
def run_single_gpu(args):
    import torch

    from genome_kmers import _native
    from genome_kmers.kmers import Kmers
    from genome_kmers.sequence_collection import SequenceCollection

    torch.cuda.set_device(0)
    lib = _native.lib()
    n_bases = args.bases
    total_len = n_bases + N_RECORDS - 1
    pinned = torch.empty(total_len, dtype=torch.uint8).pin_memory()
    host_sba, starts, names = make_genome(n_bases, N_RECORDS, RUNS_PER_RECORD, 42, out=pinned.numpy())
    n = n_kmers(n_bases, N_RECORDS, K)

    both_len = 2 * total_len + 1
    rc_starts = (total_len - 1 - np.concatenate([starts[1:].astype(np.int64) - 2, [total_len - 1]])[::-1])
    both_starts = np.ascontiguousarray(
        np.concatenate([starts, rc_starts.astype(np.uint64) + np.uint64(total_len + 1)]), dtype=np.uint64)

    d_fwd = pinned.to("cuda")
    d_both = torch.empty(both_len, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    sp = int(stream.cuda_stream)
    stats = _native.GkSortStats()
    per_step_stats = []
    last_hist = [None]

    host_phase_ms = []

    def device_step():
        """inputs (forward byte array) resident in HBM; outputs stay on the device except the histogram"""
        t = [time.perf_counter()]
        _native.check(lib.gk_sba_both_strands(d_fwd.data_ptr(), total_len, d_both.data_ptr(), sp))
        handle = ctypes.c_void_p()
        _native.check(lib.gk_index_create(d_both.data_ptr(), both_len, _native.host_ptr(both_starts),
                                          len(both_starts), K, K, ctypes.byref(handle)))
        t.append(time.perf_counter())
        try:
            _native.check(lib.gk_index_sort(handle, ctypes.byref(stats), sp))
            t.append(time.perf_counter())
            hist = np.zeros(MAX_BIN + 1, dtype=np.int64)   # fresh zero pages, as Kmers.get_kmer_group_counts does
            total, top = ctypes.c_int64(0), ctypes.c_uint64(0)
            _native.check(lib.gk_index_group_counts_zeroed(handle, K, None, 1, 0, MAX_BIN, _native.host_ptr(hist),
                                                           ctypes.byref(total), ctypes.byref(top), sp))
            assert total.value == n, (total.value, n)
            last_hist[0] = hist
            t.append(time.perf_counter())
        finally:
            lib.gk_index_destroy(handle)
        t.append(time.perf_counter())
        host_phase_ms.append([round(1e3 * (b - a), 3) for a, b in zip(t[:-1], t[1:])])
        return stats.as_dict()

    for _ in range(args.warmup):
        device_step()
    torch.cuda.synchronize()
    clocks = ClockSampler(0, period_s=args.clock_period, mode=args.clock_mode, queries=args.clock_queries)
    clocks.start()
    _native.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    step_wall_ms = []
    for _ in range(args.steps):
        t_step = time.perf_counter()
        per_step_stats.append(device_step())
        if args.clock_mode == "inline":
            clocks.sample()                  # GPU still draining the step's last kernels
        step_wall_ms.append(1e3 * (time.perf_counter() - t_step))
    ev1.record(stream)
    torch.cuda.synchronize()
    launches = _native.launch_count()
    total_ms = ev0.elapsed_time(ev1)
    clock_info = clocks.stop()
    ms_per_step = total_ms / args.steps
    value = n / (ms_per_step * 1e-3) / 1e9
    hist = last_hist[0]
    n_distinct = int(hist.sum())

    # ---- roofline of the dominant kernel: one onesweep pass moves 2*W*N bytes ----
    passes = per_step_stats[-1]["sort_passes"]
    pass_ms = float(np.mean([s["sort_ms"] for s in per_step_stats])) / max(passes, 1)
    w_bytes = 12
    algo_bytes = 2 * w_bytes * n
    achieved = algo_bytes / (pass_ms * 1e-3) / 1e9
    peak, peak_src = measured_hbm_peak()
    roofline = {
        "bound": "hbm", "kernel": "gk::onesweep_kernel (one 8-bit digit pass over (u64 key, u32 index) pairs)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": ncu_traffic_per_launch(), "peak_source": peak_src,
        "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": pass_ms, "launches_per_step": passes,
        "stage_ms": {k: float(np.mean([s[k] for s in per_step_stats]))
                     for k in ("pack_ms", "hist_ms", "sort_ms", "fixup_ms", "total_ms")},
        "whole_step_bytes_per_kmer_model": 221,
        "whole_step_frac_of_peak": (221 * n / (ms_per_step * 1e-3) / 1e9) / peak,
    }

    # ---- e2e through the Python API with host buffers ---------------------------------------------
    sc = SequenceCollection.from_sba(host_sba, starts.astype(np.uint32), names, strands_to_load="both",
                                     validate=False)

    e2e_phases = []

    def e2e_step():
        t = [time.perf_counter()]
        km = Kmers(sc, K, K, source_strand="both")
        km._ensure_device()                  # H2D of the forward byte array, both-strand layout
        t.append(time.perf_counter())
        km.sort()
        t.append(time.perf_counter())
        h, total = km.get_kmer_group_counts(K, max_counts_bin=MAX_BIN)
        t.append(time.perf_counter())
        idx = km.kmer_sba_start_indices      # D2H of the sorted start indices
        t.append(time.perf_counter())
        assert total == n and len(idx) == n
        e2e_phases.append([round(1e3 * (b - a), 3) for a, b in zip(t[:-1], t[1:])])
        return idx, h

    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        for _ in range(min(args.warmup, 2)):
            e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            idx = h = None                   # release the previous result before the next step
            idx, h = e2e_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        e2e = {"value": n / e2e_s / 1e9, "unit": UNIT, "ms_per_step": 1e3 * e2e_s, "steps": e2e_steps,
               "h2d_bytes_per_step": int(total_len),
               "d2h_bytes_per_step": int(idx.nbytes + 8 * (int(np.flatnonzero(h).max()) + 1)),
               "api": "Kmers(seq_coll, 31, 31, 'both'); sort(); get_kmer_group_counts(31); kmer_sba_start_indices",
               "phase_ms_upload_sort_count_download": e2e_phases[-e2e_steps:]}
        assert np.array_equal(h, hist)
        del idx

    # ---- CPU baseline on a bounded sample -----------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        n_cpu, dt = cpu_port_run(args.cpu_sample_bases, 1)
        cpu = {"value": n_cpu / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": (f"{args.cpu_sample_bases} bp sample of the same generator, both strands, k={K}: {n_cpu} "
                          f"k-mers in {dt:.1f} s; single-threaded C port of the reference's quicksort + comparator "
                          "+ group walk (oracle/gk_oracle.c), like the single-threaded numba reference"),
               "host_cores_available": os.cpu_count()}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload_config(1),
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "clocks": clock_info,
        "result": {"kmers": int(n), "distinct_kmers": n_distinct,
                   "ambiguous_windows": int(per_step_stats[-1]["n_ambiguous"]),
                   "key_bits": per_step_stats[-1]["key_bits"]},
        "step_wall_ms": [round(v, 3) for v in step_wall_ms],
        "host_phase_ms_create_sort_count_destroy": host_phase_ms[-min(3, len(host_phase_ms)):],
    }
    if args.bases != BASES_PER_GPU:
        line["config"]["workload"] += f" [REDUCED to {args.bases} bp: not a valid bench number]"
        line["config"]["bases_per_gpu"] = args.bases
    print(json.dumps(line), flush=True)


def run_multi_gpu(args, rank, world):
    from genome_kmers import distributed as gkd

    gkd.bench_main(args, rank, world, make_genome, workload_config, ClockSampler, METRIC, UNIT, K,
                   N_RECORDS, RUNS_PER_RECORD, MAX_BIN, measured_hbm_peak)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bases", type=int, default=BASES_PER_GPU, help="bases per GPU (default = the C2 workload)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample-bases", type=int, default=2_000_000)
    ap.add_argument("--ref-sample-bases", type=int, default=4_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clock-period", type=float, default=0.02, help="NVML sampling period in s (thread mode)")
    ap.add_argument("--clock-mode", default="inline", choices=["inline", "thread", "off"])
    ap.add_argument("--clock-queries", default="crp")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rules: at least 3 warm-up steps
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world == 1 and args.gpus == 1:
        run_single_gpu(args)
    else:
        run_multi_gpu(args, rank, world)


if __name__ == "__main__":
    main()
