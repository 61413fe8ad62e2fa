timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -m gpu -k "long_runs or repeat or Repeat" 2>&1 | tail -8
timeout 300 python bench.py --workload repeats --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > /tmp/bench_dbg.log 2>/dev/null
tail -1 /tmp/bench_dbg.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), d['verified'], [round(x,1) for x in d['step_wall_ms']], d['roofline']['stage_ms'], d['result'])"
