set -x
mkdir -p gpurun_out/r1d
O=gpurun_out/r1d
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv
nproc
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > $O/bench_512.json 2> $O/bench_512.err; tail -2 $O/bench_512.err; cat $O/bench_512.json
python bench.py --size 256 > $O/bench_256.json 2> $O/bench_256.err; cat $O/bench_256.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2>/dev/null; cat $O/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file $O/launches_r1d.csv python bench.py --size 256 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stream4 -s 6 -c 2 -f -o $O/prof_r1d python bench.py --size 256 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_stream4 -s 6 -c 2 --csv --log-file $O/traffic_512.csv python bench.py --size 512 --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_stream4 -s 6 -c 2 --csv --log-file $O/traffic_256.csv python bench.py --size 256 --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
grep -E "dram__|gpu__time" $O/traffic_512.csv $O/traffic_256.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | cut -c1-160
