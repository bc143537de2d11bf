timeout -k 10 1200 python scripts/_bisect.py convert fill3d 2>&1 | tail -4
mkdir -p gpurun_out/cx
MXADV_FCOMPARE_OUT=gpurun_out/cx timeout -k 10 1500 python -m pytest tests/test_gpu_fcompare.py tests/test_ci_extrema.py tests/test_cart.py -x -q -m gpu -k "cx or cart" 2>&1 | tail -5
