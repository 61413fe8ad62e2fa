run() { env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-verify > /tmp/bench_dbg.log 2>/dev/null
tail -1 /tmp/bench_dbg.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],2), [round(x,1) for x in d['step_wall_ms']])"; }
for rep in 1 2 3; do
run GK_BLOCK_CACHE=1
run GK_BLOCK_CACHE=0
done
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/tests_r02u.log 2>&1; tail -4 gpurun_out/tests_r02u.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r02u.log 2>/dev/null; tail -1 gpurun_out/bench_r02u.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['verified'], [round(x,1) for x in d['step_wall_ms']], d['e2e']['ms_per_step'], d['e2e']['phase_ms_upload_sort_count_download'][-2:])"
