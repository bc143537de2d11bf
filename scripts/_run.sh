O=gpurun_out/final3; mkdir -p $O
timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.log
MXADV_FCOMPARE_OUT=$O/fcompare timeout -k 10 900 python -m pytest tests/test_gpu_fcompare.py tests/test_ci_extrema.py -q -m gpu -k "cx" 2>&1 | tail -4 | cut -c1-300 | tee $O/pytest_cx.log
