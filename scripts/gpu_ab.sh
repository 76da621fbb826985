B="--steps 2 --warmup 1 --batch 2368 --no-e2e --no-cpu-baseline"
for r in 1 2; do
for d in gpurun_ab .; do
  (cd $d && python bench.py $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$d', d['value'], d['ms_per_step'], d['solved_ok'])")
done; done
