L=genome-kmers_b200/lib
for rep in 1 2; do for v in base t5 t6; do
cp $L/libgkb200_$v.so $L/libgkb200.so
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-verify --clock-mode off > /tmp/b.log 2>/dev/null
tail -1 /tmp/b.log | python -c "import sys,json,statistics; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), round(statistics.median(d['step_wall_ms']),3))"
done; done
cp $L/libgkb200_base.so $L/libgkb200.so
