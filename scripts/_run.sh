O=gpurun_out/e2e2; mkdir -p $O
timeout -k 10 900 python -m pytest tests/test_gpu_advance.py tests/test_gpu_pipeline.py tests/test_velpred.py -x -q -m gpu 2>&1 | tail -3 | tee $O/pytest.log
for t in 1 2 3 4 8; do
timeout -k 10 600 python bench.py --size 256 --steps 5 --warmup 3 --no-cpu --e2e-threads $t > $O/bench_t$t.json 2> $O/bench_t$t.err
python - <<PY
import json
d=json.loads(open('$O/bench_t$t.json').read().strip().splitlines()[-1])
print($t, 'e2e', d['e2e']['value'], 'members', d['e2e']['members']['value'])
PY
done
