mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_update_scal -s 6 -c 1 -f -o gpurun_out/prof_us python tools/bench_update_scal.py --size 256 --steps 2 > gpurun_out/ncu_us.log 2>&1
tail -2 gpurun_out/ncu_us.log
