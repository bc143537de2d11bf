O=gpurun_out/cuda2; mkdir -p $O
ARMS="ref cx" NSTEPS=10 timeout -k 10 1500 bash oracle/e2e/run_cuda_compare.sh $O big128 "amr.n_cell=128 128 1024" amr.max_grid_size=128 2>&1 | tail -40 | cut -c1-220
