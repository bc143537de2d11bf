set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for kv in v3 v4; do
  echo "== kernel=$kv"
  MXADV_KERNEL=$kv python bench.py --size 256 --steps 3 --warmup 3 --no-cpu --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"
done
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_stream -s 6 -c 2 --csv --log-file gpurun_out/v4_stats.csv python bench.py --size 256 --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
cut -d, -f5,13- gpurun_out/v4_stats.csv | tail -12
