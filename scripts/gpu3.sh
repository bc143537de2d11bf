timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for cfg in "v2 16" "v2 32" "v2 8"; do
  set -- $cfg
  echo "== kernel=$1 zc=$2"
  MXADV_KERNEL=$1 MXADV_ZCHUNK=$2 python bench.py --size 256 --steps 3 --warmup 3 --no-cpu --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"
done
ncu --set full --clock-control none --import-source on -k regex:k_stream -s 6 -c 1 -o gpurun_out/prof_r1c python bench.py --size 256 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_c.log 2>&1
tail -2 gpurun_out/ncu_full_c.log
