scripts/gpu.sh k6first 2>&1 | grep -v "^$" | tail -12
TAG=k6 scripts/gpu.sh prof 256 k_stream6
