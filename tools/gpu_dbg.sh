for env in "A=1" "GK_PACK_NOHIST=1"; do echo "== $env"
env $env timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-verify --clock-mode off 2>/dev/null | tail -1 | python -c "import sys,json,statistics; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'median wall', round(statistics.median(d['step_wall_ms']),3), {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
done
