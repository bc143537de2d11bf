O=gpurun_out/adv3; mkdir -p $O
MXADV_TRACE=1 timeout -k 10 600 python bench.py --size 128 --steps 1 --warmup 1 --no-cpu --e2e-boxes 2 > $O/bench.json 2> $O/trace.err
grep -n "pipeline" $O/trace.err | tail -60
