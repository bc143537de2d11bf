mkdir -p gpurun_out/n8
timeout -k 10 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "4-1 or 8-0" 2>&1 | tail -4
run() {  # run <label> <bench args...>
  local label=$1; shift
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-e2e "$@" > gpurun_out/n8/bench_n8_$label.json 2> gpurun_out/n8/bench_n8_$label.err
  python - gpurun_out/n8/bench_n8_$label.json $label <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], "value %.4g ms/step %.2f"%(d["value"], d["ms_per_step"]), d["config"].get("halo_exchange"))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run 512_ppm1_nspec13_nccl
run 512_ppm1_nspec13_p2p --halo p2p
run 512_ppm1_nspec13_nohalo --no-halo
run 512_sph_ppm1_nspec13 --spherical --bc outflow
run 512_ppm0_nspec3 --ppm 0 --nspec 3
run 512_ppm2_nspec13 --ppm 2
run 448_ppm1_nspec21 --nspec 21 --size 448
