O=gpurun_out/final2; mkdir -p $O
MXADV_FCOMPARE_OUT=$O/fcompare timeout -k 10 2400 python -m pytest tests -q -m gpu -rs 2>&1 | tail -22 | cut -c1-200 | tee $O/pytest_gpu.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > $O/bench_512.json 2> $O/bench_512.err; tail -c 300 $O/bench_512.json
timeout -k 10 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; tail -c 200 $O/bench_reference_arm.json
timeout -k 10 600 python bench.py --size 256 --steps 5 --warmup 3 --no-cpu > $O/bench_256.json 2> $O/bench_256.err
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_n256.csv python bench.py --size 256 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/launches.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:k_stream4 -s 6 -c 1 -o $O/prof_k_stream4_n256 -f python bench.py --size 256 --steps 1 --warmup 3 --no-cpu --no-e2e > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log
timeout -k 10 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_stream4 -s 6 -c 2 --csv --log-file $O/traffic_n512.csv python bench.py --size 512 --steps 1 --warmup 3 --no-cpu --no-e2e > $O/traffic.log 2>&1; tail -3 $O/traffic_n512.csv | cut -c150-300
timeout -k 10 300 python tools/bench_velpred.py --size 256 --ppm 1 > $O/velpred_256_ppm1.json 2> $O/velpred.err; cat $O/velpred_256_ppm1.json
