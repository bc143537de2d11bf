mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for kv in v4 v5; do
 for zc in 16 32; do
  echo "kernel=$kv zc=$zc"
  MXADV_KERNEL=$kv MXADV_ZCHUNK=$zc python bench.py --size 256 --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"
 done
done
