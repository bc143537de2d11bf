mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_stream4 -s 6 -c 1 -f -o gpurun_out/prof_skew python bench.py --size 256 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_q.log 2>&1
tail -1 gpurun_out/ncu_q.log
