set -x
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['step_wall_ms'], d['roofline']['stage_ms'])"
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --clock-mode off 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['step_wall_ms'], d['roofline']['stage_ms'])"
GK_TRACE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --clock-mode off 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['step_wall_ms'], d['roofline']['stage_ms'])"
