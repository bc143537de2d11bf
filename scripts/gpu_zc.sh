mkdir -p gpurun_out
for zc in 8 16 32 64; do
  echo "zc=$zc"
  MXADV_ZCHUNK=$zc python bench.py --size 256 --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"
done
