timeout 900 python -m pytest tests/test_update_scal.py -m gpu -x -q 2>&1 | tail -2
python tools/bench_update_scal.py --size 512
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_update_scal -s 6 -c 2 --csv --log-file gpurun_out/us_traffic.csv python tools/bench_update_scal.py --size 512 --steps 2 > /dev/null 2>&1
grep -E "dram__|gpu__time|hit_rate|warps_active" gpurun_out/us_traffic.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | cut -c1-150
