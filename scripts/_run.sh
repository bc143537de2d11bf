TAG=tma1 scripts/gpu.sh first 2>&1 | grep -v "^$" | tail -14
