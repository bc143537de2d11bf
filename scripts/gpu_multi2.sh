N=${1:-4}
mkdir -p gpurun_out
for size in 256 512; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --size $size --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_n${N}_$size.json 2> gpurun_out/bench_n${N}_$size.err
tail -2 gpurun_out/bench_n${N}_$size.err | cut -c1-300; cut -c1-700 gpurun_out/bench_n${N}_$size.json
done
