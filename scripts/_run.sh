O=gpurun_out/vp2; mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_velpred.py -q -m gpu 2>&1 | tail -4 | cut -c1-300 | tee $O/pytest_velpred.log
MXADV_FCOMPARE_OUT=$O/fcompare timeout -k 10 1500 python -m pytest tests/test_gpu_fcompare.py tests/test_ci_extrema.py -q -m gpu -k "cx" 2>&1 | tail -15 | cut -c1-400 | tee $O/pytest_cx.log
timeout -k 10 300 python tools/bench_velpred.py --size 256 --ppm 1 > $O/velpred_256_ppm1.json 2> $O/velpred.err; cat $O/velpred_256_ppm1.json; tail -3 $O/velpred.err
timeout -k 10 300 python tools/bench_velpred.py --size 256 --ppm 0 --cpu-size 0 > $O/velpred_256_ppm0.json 2>> $O/velpred.err; cat $O/velpred_256_ppm0.json
