O=gpurun_out/zc1; mkdir -p $O
timeout -k 10 900 python -m pytest tests/test_gpu_advance.py -x -q -m gpu 2>&1 | tail -8 | cut -c1-300 | tee $O/pytest.log
for v in 0 1; do
if [ $v = 1 ]; then export MXADV_NO_ZEROCOPY=1; fi
timeout -k 10 600 python bench.py --size 256 --steps 5 --warmup 3 --no-cpu > $O/bench_256_$v.json 2> $O/bench_256_$v.err
python - <<PY
import json
d=json.loads(open('$O/bench_256_$v.json').read().strip().splitlines()[-1])
print('nozerocopy=$v e2e', d['e2e']['value'], 'members', d['e2e']['members']['value'], 'launches', d['gpu_launches'])
PY
done
