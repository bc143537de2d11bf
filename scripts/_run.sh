BENCH_ARGS="--ppm 2" TAG=x3 scripts/gpu.sh ab 256 tree gy12
BENCH_ARGS="--ppm 1" TAG=x3b scripts/gpu.sh ab 256 gy12
BENCH_ARGS="--ppm 0" TAG=x3c scripts/gpu.sh ab 256 tree gy12
