O=gpurun_out/san1; mkdir -p $O
timeout -k 10 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $O/memcheck_velpred.txt python -m pytest tests/test_velpred.py -q -m gpu -x -k "device or golden" 2>&1 | tail -4 | cut -c1-200
tail -5 $O/memcheck_velpred.txt
timeout -k 10 600 compute-sanitizer --tool initcheck --error-exitcode 9 --log-file $O/initcheck_velpred.txt python -m pytest tests/test_velpred.py -q -m gpu -x -k "device and (3d-ppm1 or 2d-ppm1 or sph-ppm1)" 2>&1 | tail -3 | cut -c1-200
tail -5 $O/initcheck_velpred.txt
