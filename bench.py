#!/usr/bin/env python
"""
bench.py -- the reference's headline metric on its headline config, on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3] [--workload ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): sorted+counted k-mers per second.  One "step" = one pass of the hot path
(both-strand layout -> key pack -> onesweep radix sort -> tie repair + flags -> ambiguous-window fragments
-> group-size histogram) over one synthetic genome.

Workload at N=1 = BASELINE.json configs[1] ("C2"): synthetic 100 Mbp, 10 equal records, 20 N-runs per
record with log-uniform lengths 10^3..10^5, both strands, k=31 -> ~2.0e8 k-mers.
At N>1 every GPU brings its own 100 Mbp of genome (weak scaling): the N x 100 Mbp collection is
sorted and counted as ONE index, key-range partitioned over the ranks with one exchange over NVLink.
`--config c3` is BASELINE.json configs[2]: 3.1 Gbp, 25 records, both strands, spread over the ranks.
`--workload repeats` swaps the iid genome for one with diverged repeat families and low-complexity tracts.

Printed JSON line:
  value / ms_per_step  device-timed (CUDA events on the launching stream), inputs resident in HBM
  e2e                  same metric through the Python API with host buffers: H2D of the byte array from
                       pinned memory and D2H of the sorted start indices + histogram inside the timed region
  roofline             dominant kernel (one onesweep pass): algorithmic 2*W*N bytes / average pass time
  verified             an UNTIMED leg after the timed loop: the sorted index checked on the device against the
                       sequence bytes (gk_index_verify: neighbour order with the reference's comparator, tie
                       order, every start valid and distinct) plus the histogram identities
  cpu_baseline         the reference itself (numba, from oracle/_ref) or the C port on a bounded sample
`--impl reference` times the reference's CPU implementation of the path on a bounded sample of the workload.
"""
import argparse
import ctypes
import json
import os
import sys
import time
import types

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
C3_BASES = 3_100_000_000   # BASELINE.json configs[2]
C3_RECORDS = 25
PORT_THREADS = 16          # the C port's thread count is fixed (never taken from OMP_NUM_THREADS)


# ---------------------------------------------------------------------------------------------
# synthetic workloads
# ---------------------------------------------------------------------------------------------
def _fill_random_bases(sba, pos, ln, rng):
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    chunk = 1 << 24
    for lo in range(0, ln, chunk):
        hi = min(ln, lo + chunk)
        sba[pos + lo:pos + hi] = lut[rng.integers(0, 4, hi - lo, dtype=np.uint8)]


def _add_n_runs(sba, pos, ln, runs, rng):
    for _ in range(runs):
        run = int(np.exp(rng.uniform(np.log(1e3), np.log(1e5))))
        run = min(run, max(1, ln // 4))
        st = int(rng.integers(0, ln - run))
        sba[pos + st:pos + st + run] = ord("N")


def make_genome(n_bases, n_records, runs_per_record, seed, out=None):
    """iid uniform ACGT, equal-length records, N runs: forward byte array (records joined by '$'),
    segment starts, names."""
    rng = np.random.default_rng(seed)
    avg = n_bases // n_records
    lengths = [avg] * (n_records - 1) + [n_bases - avg * (n_records - 1)]
    total = n_bases + n_records - 1
    sba = out if out is not None else np.empty(total, dtype=np.uint8)
    assert len(sba) == total
    starts, pos = [], 0
    for r, ln in enumerate(lengths):
        starts.append(pos)
        _fill_random_bases(sba, pos, ln, rng)
        _add_n_runs(sba, pos, ln, runs_per_record, rng)
        pos += ln
        if r != n_records - 1:
            sba[pos] = ord("$")
            pos += 1
    names = [f"chr{i}" for i in range(n_records)]
    return sba, np.asarray(starts, dtype=np.uint64), names


def make_repeat_genome(n_bases, n_records, runs_per_record, seed, out=None):
    """A repeat-rich genome: the same record layout and N runs as make_genome, and on top of the iid
    background (per 100 Mbp) three families of diverged interspersed repeats -- 300 bp x 100 000 copies,
    1 kb x 5 000, 6 kb x 1 000, each copy with its own substitution rate drawn from 0-15 % -- plus 20 000
    tandem / low-complexity tracts (period 1-6, 20-500 bp).  About 40 % of the bases end up in repeats."""
    sba, starts, names = make_genome(n_bases, n_records, 0, seed, out=out)
    rng = np.random.default_rng(seed + 7919)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    scale = n_bases / 1e8
    bounds = np.concatenate([starts.astype(np.int64), [len(sba) + 1]])

    def place(unit_len, copies):
        unit = lut[rng.integers(0, 4, unit_len)]
        copies = max(1, int(copies * scale))
        batch = max(1, (1 << 24) // unit_len)
        for lo in range(0, copies, batch):
            m = min(batch, copies - lo)
            mat = np.broadcast_to(unit, (m, unit_len)).copy()
            rate = rng.uniform(0.0, 0.15, m)[:, None]
            mut = rng.random((m, unit_len)) < rate
            mat[mut] = lut[rng.integers(0, 4, int(mut.sum()))]
            rec = rng.integers(0, n_records, m)
            rec_lo, rec_hi = bounds[rec], bounds[rec + 1] - 1 - unit_len      # inside one record
            ok = rec_hi > rec_lo
            pos = (rec_lo + (rng.random(m) * np.maximum(rec_hi - rec_lo, 1)).astype(np.int64))[ok]
            sba[pos[:, None] + np.arange(unit_len)[None, :]] = mat[ok]

    place(300, 100_000)
    place(1000, 5_000)
    place(6000, 1_000)
    for _ in range(int(20_000 * scale)):
        period, ln = int(rng.integers(1, 7)), int(rng.integers(20, 501))
        rec = int(rng.integers(0, n_records))
        lo, hi = int(bounds[rec]), int(bounds[rec + 1]) - 1 - ln
        if hi <= lo:
            continue
        p = int(rng.integers(lo, hi))
        sba[p:p + ln] = np.resize(lut[rng.integers(0, 4, period)], ln)
    for r in range(n_records):
        lo, hi = int(bounds[r]), int(bounds[r + 1]) - 1
        _add_n_runs(sba, lo, hi - lo, runs_per_record, rng)
    return sba, starts, names


WORKLOADS = {"c2": make_genome, "repeats": make_repeat_genome}


def n_kmers(n_bases, n_records, k, strands=2):
    return strands * (n_bases - n_records * (k - 1))


def workload_config(n_gpus, bases_per_gpu=BASES_PER_GPU, records_per_gpu=N_RECORDS, name="c2"):
    label = {"c2": "C2 (BASELINE.json configs[1]): synthetic %d Mbp per GPU, %d records per GPU, N runs",
             "repeats": "repeat-rich variant of C2 (NOT the bench workload): %d Mbp per GPU, %d records per GPU, "
                        "diverged repeat families + low-complexity tracts + N runs",
             "c3": "C3 (BASELINE.json configs[2]): human-sized 3.1 Gbp, 25 records, %d Mbp and %d records per GPU, "
                   "N runs"}[name]
    return {
        "workload": (label + ", forward + reverse-complement strands, k=%d sort + unique counts")
                    % (bases_per_gpu // 1_000_000, records_per_gpu, K),
        "k": K, "bases_per_gpu": bases_per_gpu, "records_per_gpu": records_per_gpu, "strands": "both",
        "kmers_total": n_kmers(bases_per_gpu * n_gpus, records_per_gpu * n_gpus, K),
        "max_counts_bin": MAX_BIN,
        "parallelism": "single GPU" if n_gpus == 1 else f"key-range sharded x{n_gpus}, one exchange over NVLink",
        "l2_policy": "inputs larger than L2 (>= 2.4 GB of key/index pairs per pass vs 126 MB L2)",
    }


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
    """(dram bytes per onesweep launch, where it comes from): a constant read from the committed ncu capture of
    this workload, NOT a measurement of the run being reported (ncu cannot run inside the timed region)."""
    path = os.path.join(ROOT, "profiles", "onesweep_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return d.get("dram_bytes_per_launch"), "profile constant: " + d.get("source", "profiles/onesweep_traffic.json")
    except Exception:
        return None, None


# ---------------------------------------------------------------------------------------------
# CPU arms.  (1) the UNMODIFIED reference (pure Python + numba), vendored from /root/reference into
# oracle/_ref by oracle/make_ref.py in the build container; it needs numba, which the image has.
# (2) the C port of its algorithm (oracle/gk_oracle.c) when the reference cannot be imported.
# ---------------------------------------------------------------------------------------------
_COMP = bytes.maketrans(b"ACGTRYSWKMBDHVN", b"TGCAYRSWMKVHDBN")


def load_reference():
    """(kmers module, sequence_collection module) of the reference, or (None, why).  Must be called in a
    process that has not imported this repo's own `genome_kmers` package (same name)."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "genome_kmers", "kmers.py")):
        return None, "oracle/_ref is missing (made by oracle/make_ref.py where /root/reference exists)"
    if "genome_kmers" in sys.modules:
        return None, "this repo's genome_kmers is already imported in this process"
    try:
        import numba  # noqa: F401
    except Exception as exc:  # pragma: no cover
        return None, f"numba is not importable: {exc}"
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))   # absent in the image; only save/load use it
    saved = list(sys.path)
    try:
        sys.path.insert(0, ref_dir)
        from genome_kmers import kmers as ref_kmers
        from genome_kmers import sequence_collection as ref_sc
    except Exception as exc:
        for name in [m for m in sys.modules if m == "genome_kmers" or m.startswith("genome_kmers.")]:
            del sys.modules[name]
        return None, f"the reference failed to import: {exc}"
    finally:
        sys.path[:] = saved
    return (ref_kmers, ref_sc), None


def reference_sequence_list(sba, starts):
    """The reference's own input for 'both strands' (SURVEY.md 8c): forward records followed by the
    reverse-complemented records in reversed order; its sba is then forward || '$' || revcomp."""
    bounds = list(starts.astype(np.int64)) + [len(sba) + 1]
    recs = [sba[bounds[i]:bounds[i + 1] - 1].tobytes() for i in range(len(starts))]
    fwd = [(f"chr{i}", r.decode()) for i, r in enumerate(recs)]
    rc = [(f"chr{i}_rc", recs[i].translate(_COMP)[::-1].decode()) for i in reversed(range(len(recs)))]
    return fwd + rc


def reference_run(ref, sample_bases, seed=7, n_rec=2, runs=2):
    """One sort() + get_kmer_group_counts() of the real reference (single-threaded, numba JIT included)."""
    ref_kmers, ref_sc = ref
    if sample_bases <= 0:     # JIT-only calibration: the same calls on a 200-bp collection
        rng = np.random.default_rng(seed)
        sba = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 200)].copy()
        starts = np.zeros(1, dtype=np.uint64)
    else:
        sba, starts, _ = make_genome(sample_bases, n_rec, runs, seed)
    sc = ref_sc.SequenceCollection(sequence_list=reference_sequence_list(sba, starts), strands_to_load="forward")
    km = ref_kmers.Kmers(sc, min_kmer_len=K, max_kmer_len=K)
    n = len(km)
    t0 = time.perf_counter()
    km.sort()
    hist, total = km.get_kmer_group_counts(K)
    dt = time.perf_counter() - t0
    assert total == n
    return n, dt


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
    ref, why = (None, "--ref-kind port") if args.ref_kind == "port" else load_reference()
    config = workload_config(args.gpus)
    if ref is not None:
        # every sort() / count call of the reference compiles a fresh numba closure (kmers.py:1641-1645, :1156):
        # the warm-up steps run the same calls on a 200-bp collection and measure exactly that
        # (at least three calibration calls, the fastest counts: the first also compiles the module-level
        # functions once per process, which no later call pays)
        sample_bases = args.ref_sample_bases or 1_000_000
        jit = [reference_run(ref, 0, seed=50 + s)[1] for s in range(max(3, args.warmup))]
        jit_s = float(min(jit))
        raw, n = [], 0
        for s in range(args.steps):
            n, dt = reference_run(ref, sample_bases, seed=100 + s)
            raw.append(dt)
        net_total = sum(raw) - jit_s * args.steps
        jit_clamped = net_total < 0.05 * sum(raw)     # (cannot happen at the default sample size: ~1.5 s of sorting)
        net_total = max(net_total, 0.05 * sum(raw))
        value = n * args.steps / net_total / 1e9
        cores, kind = 1, "reference"
        ms_per_step = 1e3 * net_total / args.steps
        sample = (f"UNMODIFIED reference (numba, oracle/_ref): Kmers.sort() + get_kmer_group_counts({K}) on a "
                  f"{sample_bases} bp sample of the workload generator (2 records, N runs), forward records + "
                  f"reverse-complemented records = both strands, {n} k-mers per step, single-threaded like the "
                  f"reference; numba compile time ({jit_s:.1f} s per step, measured on a 200-bp collection) is "
                  f"subtracted; with it the value is {n * args.steps / sum(raw) / 1e9:.6f} {UNIT}")
        extra = {"jit_s_per_step": jit_s, "raw_s_per_step": float(np.mean(raw)), "jit_estimate_clamped": bool(jit_clamped),
                 "jit_calibration_s": [round(t, 3) for t in jit]}
    else:
        threads = max(1, min(PORT_THREADS, os.cpu_count() or 1))
        sample_bases = args.ref_sample_bases or 4_000_000
        for _ in range(args.warmup):
            cpu_port_run(min(sample_bases, 200_000), threads)
        times, n = [], 0
        for s in range(args.steps):
            n, dt = cpu_port_run(sample_bases, threads, seed=100 + s)
            times.append(dt)
        value = n * args.steps / sum(times) / 1e9
        cores, kind = threads, "port"
        ms_per_step = 1e3 * sum(times) / args.steps
        sample = (f"C port of the reference's quicksort + comparator + group walk (oracle/gk_oracle.c) on {threads} "
                  f"OpenMP threads (fixed), {sample_bases} bp sample of the workload generator (2 records, N runs), "
                  f"both strands, {n} k-mers per step; the reference itself was not used: {why}")
        extra = {}
    config.update(sampled=True, sample_bases=sample_bases, sample_kmers_per_step=int(n),
                  note="the CPU arm runs a bounded sample of the workload named above, not its full size")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": config,
        "cpu_baseline": dict({"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}, **extra),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_for_our_arm(args):
    """Rank 0, N=1: the reference (or the port) on a bounded sample, in a child process so that the reference's
    `genome_kmers` never meets this repo's package of the same name."""
    import subprocess

    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--ref-sample-bases", str(args.cpu_sample_bases)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
        line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
        cpu = line["cpu_baseline"]
        cpu["host_cores_available"] = os.cpu_count()
        return cpu
    except Exception as exc:  # the bench line must survive a failing baseline leg
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": f"failed: {exc}"}


# ---------------------------------------------------------------------------------------------
# GPU arm, one GPU
# ---------------------------------------------------------------------------------------------
def bytes_moved_per_step(n, passes, both_len, idx_bytes=4):
    """HBM bytes the kernels of one step actually move (algorithmic): both-strand layout
    (read L, write 2L) + alphabet scan (2L) + pack (1 B + key + start per window) + `passes` radix passes
    (read + write a pair each) + tie repair / flags (key read, flag write) + histogram (flags read twice)."""
    w = 8 + idx_bytes
    return int(2.5 * both_len + n * ((1 + w) + 2 * w * passes + 9 + 2))


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
    host_sba, starts, names = WORKLOADS[args.workload](n_bases, N_RECORDS, RUNS_PER_RECORD, 42, out=pinned.numpy())
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

    def device_step(keep=False):
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
            hist = _native.zeros_int64(MAX_BIN + 1)        # fresh zero pages, as Kmers.get_kmer_group_counts does
            total, top = ctypes.c_int64(0), ctypes.c_uint64(0)
            _native.check(lib.gk_index_group_counts_zeroed(handle, K, None, 1, 0, MAX_BIN, _native.host_ptr(hist),
                                                           ctypes.byref(total), ctypes.byref(top), sp))
            assert total.value == n, (total.value, n)
            last_hist[0] = hist
            t.append(time.perf_counter())
            if keep:
                return handle
        finally:
            if not keep:
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

    # ---- verification leg (untimed): the order of one more step, checked on the device against the bytes ----
    verification = {"checked": False}
    if not args.no_verify:
        handle = device_step(keep=True)
        try:
            report = np.zeros(8, dtype=np.uint64)
            _native.check(lib.gk_index_verify(handle, K, _native.host_ptr(report), None, sp))
        finally:
            lib.gk_index_destroy(handle)
        rep = dict(zip(("kmers", "out_of_order", "tie_order", "invalid_starts", "duplicate_starts", "groups",
                        "flag_mismatches", "flags_compared"), (int(v) for v in report)))
        sizes = np.flatnonzero(last_hist[0])
        checks = {
            "every window start present exactly once (permutation of the init set)":
                rep["kmers"] == n and rep["invalid_starts"] == 0 and rep["duplicate_starts"] == 0,
            "neighbours in non-decreasing order under the reference's byte comparator": rep["out_of_order"] == 0,
            "equal k-mers in ascending start order (break_ties=True order)": rep["tie_order"] == 0,
            "head flags of the sort agree with the bytes": rep["flags_compared"] == 1 and rep["flag_mismatches"] == 0,
            "histogram: number of groups equals the groups counted from the bytes":
                int(last_hist[0].sum()) == rep["groups"],
            "histogram: sum of size x count equals the number of k-mers (no bin clamped)":
                int((last_hist[0][sizes] * sizes).sum()) == n or int(sizes.max()) == MAX_BIN,
        }
        verification = {"checked": True, "report": rep, "checks": checks}
    verified = bool(verification["checked"] and all(verification["checks"].values()))

    # ---- roofline of the dominant kernel: one onesweep pass moves 2*W*N bytes ----
    passes = per_step_stats[-1]["sort_passes"]
    pass_ms = float(np.mean([s["sort_ms"] for s in per_step_stats])) / max(passes, 1)
    w_bytes = 12
    algo_bytes = 2 * w_bytes * n
    achieved = algo_bytes / (pass_ms * 1e-3) / 1e9
    peak, peak_src = measured_hbm_peak()
    traffic, traffic_src = ncu_traffic_per_launch()
    moved = bytes_moved_per_step(n, passes, both_len)
    roofline = {
        "bound": "hbm", "kernel": "gk::onesweep_kernel (one 8-bit digit pass over (u64 key, u32 index) pairs)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": pass_ms, "launches_per_step": passes,
        "stage_ms": {k: float(np.mean([s[k] for s in per_step_stats]))
                     for k in ("pack_ms", "hist_ms", "sort_ms", "fixup_ms", "total_ms")},
        "whole_step": {"bytes_moved": moved, "bytes_moved_per_kmer": moved / n,
                       "frac_of_peak": moved / (ms_per_step * 1e-3) / 1e9 / peak,
                       "note": "bytes the step's kernels actually move (bytes_moved_per_step in bench.py)",
                       "model_221B_per_kmer_equivalent_frac": (221 * n / (ms_per_step * 1e-3) / 1e9) / peak,
                       "model_note": "SURVEY.md 8d's end-to-end model assumes 8 radix passes; this code runs "
                                     f"{passes}, so the model-equivalent figure is not a bandwidth"},
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

    cpu = None if args.no_cpu_baseline else cpu_baseline_for_our_arm(args)

    last = per_step_stats[-1]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(1, args.bases, N_RECORDS, args.workload),
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "clocks": clock_info, "verified": verified, "verification": verification,
        "result": {"kmers": int(n), "distinct_kmers": n_distinct,
                   "ambiguous_windows": int(last["n_ambiguous"]), "fragments": int(last["n_fragments"]),
                   "refine_flags": int(last["refine_flags"]), "key_bits": last["key_bits"]},
        "step_wall_ms": [round(v, 3) for v in step_wall_ms],
        "host_phase_ms_create_sort_count_destroy": host_phase_ms[-min(3, len(host_phase_ms)):],
    }
    if args.bases != BASES_PER_GPU:
        line["config"]["workload"] += f" [REDUCED to {args.bases} bp: not a valid bench number]"
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm, N > 1: weak scaling, one index over all ranks' genomes
# ---------------------------------------------------------------------------------------------
def run_multi_gpu(args, rank, world):
    import torch
    import torch.distributed as dist

    from genome_kmers import _native
    from genome_kmers.distributed import NativeEngine, PeerExchange, ShardedKmers

    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = NativeEngine()
    n_bases, n_records = args.bases, args.records
    make = WORKLOADS[args.workload]

    # every rank generates its own chunk of the genome and the ranks all-gather the forward byte array;
    # the trailing '$' of a rank's chunk separates it from the next rank's first record
    chunk_len = n_bases + n_records
    pinned = torch.empty(chunk_len, dtype=torch.uint8).pin_memory()
    host = pinned.numpy()
    sba, starts, _ = make(n_bases, n_records, RUNS_PER_RECORD, 42 + rank, out=host[:chunk_len - 1])
    host[chunk_len - 1] = ord("$")
    all_starts = np.concatenate([starts + np.uint64(r * chunk_len) for r in range(world)])
    total_fwd = world * chunk_len - 1
    n_total = 2 * (n_bases * world - n_records * world * (K - 1))

    def load_inputs():
        d_chunk = pinned.to("cuda", non_blocking=True)
        d_all = torch.empty(world * chunk_len, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(d_all, d_chunk)
        return d_all[:total_fwd]

    d_fwd = load_inputs()
    hist = None

    def step(d_forward, keep=False):
        m0 = eng.mark()
        sk = ShardedKmers(d_forward, all_starts, K, "both", engine=eng)
        m1 = eng.mark()
        sk.sort()
        h, total = sk.get_kmer_group_counts(K, max_counts_bin=MAX_BIN)
        assert total == n_total, (total, n_total)
        stats, sent = dict(sk.stats), sk.exchange_bytes_sent
        stats["_mode"] = sk.exchange_mode
        if keep:
            return h, sk
        sk.close()
        m3 = eng.mark()
        stats["_sk_marks"] = [("begin", m0), ("both_strands", m1)] + sk._marks[1:] + [("count_allgather", m3)]
        return h, stats, sent

    for _ in range(args.warmup):
        step(d_fwd)
    _native.launch_count(reset=True)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    stream = torch.cuda.current_stream()
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    per_step = []
    for _ in range(args.steps):
        hist, stats, sent = step(d_fwd)
        if rank == 0:
            clocks.sample()
        per_step.append((stats, sent))
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = _native.launch_count()
    ms_per_step = float(ms.item()) / args.steps
    clock_info = clocks.stop() if rank == 0 else None

    # ---- verification leg (untimed): every shard checked on its device, the shard boundaries and the global
    # ---- permutation property checked across ranks ----------------------------------------------------------
    verification = {"checked": False}
    if not args.no_verify:
        h, sk = step(d_fwd, keep=True)
        verification = sk.verify(h, n_total)
        sk.close()
    verified = bool(verification.get("checked") and all(verification["checks"].values()))

    # e2e: host chunk -> H2D -> all-gather -> sort/count -> shard of sorted starts back on the host
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        shard_bytes = 0

        def e2e_step():
            d_in = load_inputs()
            sk = ShardedKmers(d_in, all_starts, K, "both", engine=eng)
            sk.sort()
            h, total = sk.get_kmer_group_counts(K, max_counts_bin=MAX_BIN)
            local = sk.local_start_indices()     # D2H of this rank's shard of the sorted starts (pinned)
            sk.close()
            assert total == n_total
            return int(local.nbytes)

        for _ in range(2):                       # warm the pinned-buffer cache and the NCCL channels
            e2e_step()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            shard_bytes = e2e_step()
        torch.cuda.synchronize()
        dist.barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": n_total / float(dt.item()) / 1e9, "unit": UNIT, "ms_per_step": 1e3 * float(dt.item()),
               "steps": e2e_steps, "h2d_bytes_per_step": int(chunk_len) * world,
               "d2h_bytes_per_step": shard_bytes * world,
               "api": "ShardedKmers(...).sort(); get_kmer_group_counts(); local_start_indices() on every rank"}

    idx_bytes = 8 if (2 * (world * chunk_len - 1) + 1 > 0xFFFFFFFF
                      or os.environ.get("GK_FORCE_IDX64", "0") not in ("", "0")) else 4
    last = per_step[-1][0]
    mine = [last.get("total_ms", 0.0), last.get("fixup_ms", 0.0), float(last.get("n_shard", 0)),
            float(last.get("n_ambiguous", 0)), float(last.get("n_fragments", 0)), float(last.get("refine_flags", 0)),
            last.get("pack_ms", 0.0), last.get("hist_ms", 0.0), last.get("sort_ms", 0.0)]
    per_rank = [None] * world
    dist.all_gather_object(per_rank, mine)
    sent_all = torch.tensor([float(np.mean([s for _, s in per_step]))], dtype=torch.float64, device="cuda")
    dist.all_reduce(sent_all, op=dist.ReduceOp.SUM)
    phase_ms = {}
    exchange_mode = per_step[-1][0].pop("_mode", "nccl")
    for stats, _ in per_step:
        stats.pop("_mode", None)
        marks = stats.pop("_sk_marks", [])
        for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
            phase_ms[name] = phase_ms.get(name, 0.0) + a.elapsed_time(b) / len(per_step)
    # fused partition + peer writes (its own mark when the peer path ran) + fragment gather + the ordering all-reduce
    exchange_ms = phase_ms.get("exchange", 0.0) + phase_ms.get("partition_kernel", 0.0)
    if rank == 0:
        passes = per_step[-1][0]["sort_passes"]
        pass_ms = float(np.mean([s["sort_ms"] for s, _ in per_step])) / max(passes, 1)
        n_shard = per_step[-1][0]["n_shard"]
        peak, peak_src = measured_hbm_peak()
        pair_bytes = 8 + idx_bytes
        achieved = 2 * pair_bytes * n_shard / (pass_ms * 1e-3) / 1e9
        name = "c3" if args.config == "c3" else args.workload
        line = {
            "metric": METRIC, "value": n_total / (ms_per_step * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(world, n_bases, n_records, name),
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "gk::onesweep_kernel on rank 0's key range",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "avg_launch_ms": pass_ms,
                         "launches_per_step": passes, "pairs_on_rank0": int(n_shard),
                         "pair_bytes": pair_bytes},
            "exchange": {"bytes_over_nvlink_per_step": float(sent_all.item()), "mode": exchange_mode,
                         # rank 0's exchange phase: fused partition + peer writes + the ordering all-reduce
                         "exchange_ms_rank0": round(float(exchange_ms), 3),
                         "partition_kernel_ms_rank0": round(float(phase_ms.get("partition_kernel", 0.0)), 3),
                         "nvlink_gbs_per_gpu_outbound": (
                             round(float(sent_all.item()) / world / (exchange_ms * 1e-3) / 1e9, 1)
                             if exchange_ms > 0 else None),
                         "nvlink_peak_gbs_per_direction": 900.0, "nvlink_measured_peer_copy_gbs": 770.0,
                         "note": "(G-1)/G of the pure (u64 key, start) pairs cross NVLink once: written by the "
                                 "partition kernel into peer memory (mode peer) or one NCCL all-to-all (mode nccl); "
                                 "ambiguous windows travel as run-length fragments"},
            "cpu_baseline": None, "clocks": clock_info, "verified": verified, "verification": verification,
            "phase_ms_rank0": {k_: round(v, 3) for k_, v in phase_ms.items()},
            "per_rank_local_sort": {"total_ms": [round(r[0], 3) for r in per_rank],
                                    "refine_ms": [round(r[1], 3) for r in per_rank],
                                    "pairs": [int(r[2]) for r in per_rank],
                                    "ambiguous": [int(r[3]) for r in per_rank],
                                    "fragments": [int(r[4]) for r in per_rank],
                                    "refine_flags": [int(r[5]) for r in per_rank],
                                    "placeholders_hist_passes_ms": [[round(r[6], 3), round(r[7], 3), round(r[8], 3)]
                                                                    for r in per_rank]},
            "local_sort_stats_rank0": {k_: v for k_, v in per_step[-1][0].items()},
            "result": {"kmers": int(n_total), "distinct_kmers": int(hist.sum())},
        }
        if args.config != "c3" and n_bases != BASES_PER_GPU:
            line["config"]["workload"] += f" [REDUCED to {n_bases} bp per GPU: not a valid bench number]"
        print(json.dumps(line), flush=True)
    dist.barrier()
    PeerExchange.close_all()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3"],
                    help="c2: 100 Mbp per GPU (the bench workload); c3: 3.1 Gbp, 25 records over all GPUs")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="genome generator")
    ap.add_argument("--bases", type=int, default=BASES_PER_GPU, help="bases per GPU (default = the C2 workload)")
    ap.add_argument("--records", type=int, default=N_RECORDS, help="records per GPU")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample-bases", type=int, default=0,
                    help="sample size of the cpu_baseline leg (0: 4 Mbp for the reference, 2 Mbp for the port)")
    ap.add_argument("--ref-sample-bases", type=int, default=0,
                    help="sample size per step of --impl reference (0: 1 Mbp for the reference, 4 Mbp for the port)")
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "port"],
                    help="auto: the real reference from oracle/_ref when numba imports, else the C port")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the untimed verification leg")
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
    if args.cpu_sample_bases == 0:
        args.cpu_sample_bases = 4_000_000 if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "genome_kmers")) \
            else 2_000_000
    if args.config == "c3":
        n = max(world, args.gpus)
        args.bases = C3_BASES // n
        args.records = max(1, (C3_RECORDS + n - 1) // n)
    if world == 1 and args.gpus == 1:
        if args.config == "c3":
            raise SystemExit("--config c3 needs several GPUs: 6.2e9 (key, start) pairs and their ping-pong copies "
                             "(198 GB) do not fit one 180 GB B200; launch with torchrun --nproc-per-node 8")
        run_single_gpu(args)
    else:
        run_multi_gpu(args, rank, world)


if __name__ == "__main__":
    main()
