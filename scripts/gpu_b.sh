set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_halo.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --size 256 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_2gpu_256.json 2> gpurun_out/bench_2gpu_256.err; tail -5 gpurun_out/bench_2gpu_256.err; cat gpurun_out/bench_2gpu_256.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_512.json 2> gpurun_out/bench_2gpu_512.err; tail -5 gpurun_out/bench_2gpu_512.err; cat gpurun_out/bench_2gpu_512.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --impl reference
