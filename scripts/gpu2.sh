timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for cfg in "v1 16 32" "v2 16 32" "v2 8 32" "v2 16 16" "v2 16 64"; do
  set -- $cfg
  echo "== kernel=$1 GY=$2 zc=$3"
  MXADV_KERNEL=$1 MXADV_GY=$2 MXADV_ZCHUNK=$3 python bench.py --size 256 --steps 3 --warmup 3 --no-cpu --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"
done
