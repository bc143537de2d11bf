for t in 1 3 2 3 1 4 6 8; do
python bench.py --size 128 --steps 4 --warmup 3 --no-cpu --e2e-threads $t 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('threads $t', round(d['e2e']['value']/1e9,3))"
done
