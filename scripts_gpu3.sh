timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:k_stream -s 6 -c 1 -o gpurun_out/prof_r1b python bench.py --n 256 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_b.log 2>&1
tail -2 gpurun_out/ncu_full_b.log
