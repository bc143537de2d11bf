O=gpurun_out/n2final; mkdir -p $O
timeout -k 10 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -rs 2>&1 | tail -8 | cut -c1-200 | tee $O/pytest_multi.log
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2_512.json 2> $O/bench_n2.err; tail -c 600 $O/bench_n2_512.json
