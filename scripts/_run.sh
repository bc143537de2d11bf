timeout -k 10 900 python -m pytest tests/test_cart.py -x -q -m gpu 2>&1 | tail -15
