N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
for size in 256 512; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --size $size --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_n${N}_$size.json 2> gpurun_out/bench_n${N}_$size.err
tail -3 gpurun_out/bench_n${N}_$size.err; cat gpurun_out/bench_n${N}_$size.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --size $size --steps 3 --warmup 3 --no-e2e --no-halo > gpurun_out/bench_n${N}_${size}_nohalo.json 2>/dev/null
cat gpurun_out/bench_n${N}_${size}_nohalo.json
done
