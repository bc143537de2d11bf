set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv
nproc
timeout 900 python -m pytest tests/test_w0.py -m gpu -x -q 2>&1 | tail -5
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --size 256 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_256_f.json 2> gpurun_out/bench_256_f.err; tail -3 gpurun_out/bench_256_f.err; cat gpurun_out/bench_256_f.json
