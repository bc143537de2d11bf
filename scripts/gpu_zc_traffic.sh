mkdir -p gpurun_out
for zc in 8 16 32; do
  MXADV_ZCHUNK=$zc ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_stream -s 2 -c 1 --csv --log-file gpurun_out/traffic_zc$zc.csv python bench.py --size 256 --steps 1 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
  echo "zc=$zc"; grep -E "dram__bytes|gpu__time|hit_rate" gpurun_out/traffic_zc$zc.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
