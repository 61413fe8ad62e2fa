"""
CPU checks of bench.py's contract that need no GPU: the `--impl reference` arm (the CPU port of the reference's
algorithm on a bounded sample) prints ONE JSON line with the agreed keys, also when launched the way the driver
launches N > 1 (torchrun: rank 0 alone works, the other ranks exit 0 without output); the workload generator
is deterministic and has the shape BASELINE.json's configs[1] asks for.
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
            "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _json_lines(text):
    return [json.loads(line) for line in text.splitlines() if line.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--ref-sample-bases", "40000"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1
    line = lines[0]
    assert REQUIRED <= set(line), REQUIRED - set(line)
    assert line["impl"] == "reference" and line["unit"] == "Gkmer/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["n_gpus"] == 1 and line["gpu_launches"] == 0
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "C2" in line["config"]["workload"] and line["config"]["k"] == 31
    # the CPU arm runs a bounded sample and says so: the config is not passed off as the full workload
    assert line["config"]["sampled"] is True and line["config"]["sample_bases"] == 40000
    if os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "genome_kmers")):
        assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == 1


def test_reference_arm_port_has_a_fixed_thread_count_under_torchrun_environment():
    """OMP_NUM_THREADS=1 (what torchrun exports) must not change the port's thread count."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--ref-kind", "port", "--steps", "1",
                          "--warmup", "1", "--ref-sample-bases", "40000"], cwd=ROOT, capture_output=True, text=True,
                         timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = _json_lines(out.stdout)[0]
    assert line["cpu_baseline"]["kind"] == "port"
    assert line["cpu_baseline"]["cores"] == max(1, min(16, os.cpu_count() or 1))


def test_reference_arm_under_torchrun_only_rank0_reports():
    import socket

    with socket.socket() as sock:                      # a free port for the rendezvous
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
         "127.0.0.1", "--master-port", str(port), "bench.py", "--impl", "reference", "--ref-kind", "port", "--gpus", "2",
         "--steps", "1", "--warmup", "1", "--ref-sample-bases", "30000"],
        cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2
    assert lines[0]["config"]["kmers_total"] == 2 * 199999400


def test_workload_generator_is_deterministic_and_has_the_config2_shape():
    sys.path.insert(0, ROOT)
    import bench

    sba, starts, names = bench.make_genome(200_000, 10, 20, 42)
    again, starts2, _ = bench.make_genome(200_000, 10, 20, 42)
    assert np.array_equal(sba, again) and np.array_equal(starts, starts2)
    assert len(sba) == 200_000 + 9 and len(starts) == 10 and names[0] == "chr0"
    assert int((sba == ord("$")).sum()) == 9
    assert set(np.unique(sba).tolist()) <= set(b"ACGTN$")
    assert 0 < float((sba == ord("N")).mean()) < 1.0        # N runs are present
    assert bench.n_kmers(100_000_000, 10, 31) == 199_999_400
    assert bench.workload_config(1)["kmers_total"] == 199_999_400


def test_repeat_rich_generator_has_repeats_and_the_same_layout():
    sys.path.insert(0, ROOT)
    import bench

    sba, starts, names = bench.make_repeat_genome(2_000_000, 10, 2, 42)
    again, _, _ = bench.make_repeat_genome(2_000_000, 10, 2, 42)
    assert np.array_equal(sba, again)
    assert len(sba) == 2_000_000 + 9 and int((sba == ord("$")).sum()) == 9
    assert np.array_equal(np.flatnonzero(sba == ord("$")) + 1, starts[1:].astype(np.int64))
    assert set(np.unique(sba).tolist()) <= set(b"ACGTN$")
    # repeats: many 20-mers occur more than once (an iid genome of this size has almost none)
    w = np.lib.stride_tricks.sliding_window_view(sba[:400_000], 20)[::7]
    keys = np.unique(w.view(np.dtype((np.void, 20))).ravel(), return_counts=True)[1]
    assert (keys > 1).sum() > 500
