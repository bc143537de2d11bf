O=gpurun_out/async1; mkdir -p $O
for i in 1 2 3; do
MXADV_ASYNC_HOST=1 timeout -k 10 600 python bench.py --size 128 --steps 5 --warmup 3 --no-cpu --e2e-threads 2 > $O/bench_$i.json 2> $O/err.log
python - <<PY
import json
d=json.loads(open('$O/bench_$i.json').read().strip().splitlines()[-1])
print('async run $i e2e', d['e2e']['value'], 'members', d['e2e']['members']['value'])
PY
done
tail -3 $O/err.log
