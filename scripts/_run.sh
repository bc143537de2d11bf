O=gpurun_out/vp1; mkdir -p $O
timeout -k 10 900 python -m pytest tests/test_velpred.py -q -m gpu 2>&1 | tail -40 | cut -c1-300 | tee $O/pytest_velpred.log
