set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv
nproc
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --size 256 --steps 5 --warmup 3 > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err; tail -3 gpurun_out/bench_256.err; cat gpurun_out/bench_256.json
python bench.py --size 512 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; tail -3 gpurun_out/bench_512.err; cat gpurun_out/bench_512.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --size 256 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stream -s 6 -c 2 -o gpurun_out/prof_r1b python bench.py --size 256 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
