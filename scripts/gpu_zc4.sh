for zc in 32 36 42 46 51 56 63; do
  MXADV_ZCHUNK=$zc python bench.py --size 512 --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('512 zc=$zc', round(d['roofline']['ms_per_launch'],2))"
done
for zc in 43 52 62; do
  MXADV_ZCHUNK=$zc python bench.py --size 256 --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('256 zc=$zc', round(d['roofline']['ms_per_launch'],2))"
done
