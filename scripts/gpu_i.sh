timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][18:], round(d['ms_per_step'],2), round(d['roofline']['ms_per_launch'],2), round(d['roofline']['frac'],4))"; }
run --size 256
run --size 512
run --size 512 --spherical --bc outflow
