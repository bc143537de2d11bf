timeout -k 10 1500 python -m pytest tests/test_ci_extrema.py -x -q -m gpu -k wdconvect 2>&1 | tail -12 | cut -c1-300
