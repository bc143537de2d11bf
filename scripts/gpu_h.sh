timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spherical or fused" 2>&1 | tail -3
run() { python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][18:], round(d['ms_per_step'],2), round(d['roofline']['ms_per_launch'],2), round(d['roofline']['frac'],4))"; }
run --size 512
run --size 256 --spherical --bc outflow
run --size 512 --spherical --bc outflow
run --size 512 --bc outflow
