for bc in reacting_bubble periodic; do
for zc in 16 20 24 26 28 32 36 40 43 48 52; do
  MXADV_ZCHUNK=$zc python bench.py --size 256 --bc $bc --steps 4 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$bc zc=$zc', round(d['roofline']['ms_per_launch'],3))"
done; done
