mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_stream5 -s 2 -c 1 -f -o gpurun_out/prof_v5 python bench.py --size 256 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_v5.log 2>&1
tail -3 gpurun_out/ncu_v5.log
ls -la gpurun_out/
