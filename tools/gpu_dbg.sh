set -x
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/tests_r02r.log 2>&1; tail -5 gpurun_out/tests_r02r.log
timeout 300 python tools/owner_like.py 0.35 2>&1 | tail -3 | cut -c1-700
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02r.log 2> gpurun_out/bench_r02r.err
tail -1 gpurun_out/bench_r02r.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['verified'], d['step_wall_ms'], d['roofline']['stage_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
