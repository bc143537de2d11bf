set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_overlay.py tests/test_gpu_halo.py -m gpu -x -q 2>&1 | tail -5
for zc in 8 16 32 64; do
  echo "== zc=$zc"
  MXADV_ZCHUNK=$zc python bench.py --size 256 --steps 3 --warmup 3 --no-cpu --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"
done
for zc in 16 32; do
MXADV_ZCHUNK=$zc ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_stream -s 6 -c 2 --csv --log-file gpurun_out/dram_zc$zc.csv python bench.py --size 256 --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
cat gpurun_out/dram_zc$zc.csv | tail -8
done
