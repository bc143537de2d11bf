for name in base gy12 gy8; do
  if [ "$name" = base ]; then unset MXADV_LIB; else export MXADV_LIB=/root/repo/exp/$name/libmxadv.so; fi
  echo "== $name"
  MXADV_KERNEL=v4 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:k_stream4 -s 2 -c 1 --csv --log-file gpurun_out/exp2_$name.csv python bench.py --size 256 --steps 1 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
  grep -E "gpu__time|inst_executed|issue_active|fp64|registers" gpurun_out/exp2_$name.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
