O=gpurun_out/x1; mkdir -p $O
# tolerance-mode A/B errors (the earlier run lacked the import path)
timeout -k 10 300 python tools/fmad_err.py dump $O/strict.npz; MXADV_LIB=exp/fmad/libmxadv.so timeout -k 10 300 python tools/fmad_err.py dump $O/fmad.npz; python tools/fmad_err.py compare $O/strict.npz $O/fmad.npz | tee $O/fmad_err.txt; rm -f $O/strict.npz $O/fmad.npz
# pair-sync variant: parity first
MXADV_LIB=exp/pair/libmxadv.so timeout -k 10 600 python -m pytest tests/test_gpu_tma.py tests/test_gpu_fullsize.py -x -q -m gpu -k "not 512" 2>&1 | tail -5 | tee $O/pytest_pair.log
TAG=x1 scripts/gpu.sh ab 256 tree pair fakein gy20 gy18
TAG=x1 scripts/gpu.sh ab 512 tree pair
