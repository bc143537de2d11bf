O=gpurun_out/vp7; mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_velpred.py -q -m gpu -x 2>&1 | tail -3 | cut -c1-300 | tee $O/pytest_velpred.log
timeout -k 10 300 python tools/bench_velpred.py --size 256 --ppm 1 | tee $O/velpred_256_ppm1.json
timeout -k 10 300 python tools/bench_velpred.py --size 256 --ppm 0 --cpu-size 0 | tee $O/velpred_256_ppm0.json
