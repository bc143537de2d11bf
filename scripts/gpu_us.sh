timeout 900 python -m pytest tests/test_update_scal.py -m gpu -x -q 2>&1 | tail -5
python tools/bench_update_scal.py --size 256
python tools/bench_update_scal.py --size 512
