mkdir -p gpurun_out/adv
timeout -k 10 900 python -m pytest tests/test_gpu_advance.py tests/test_w0.py -x -q -m gpu 2>&1 | tail -12
timeout -k 10 900 python bench.py --size 256 --steps 5 --warmup 3 --no-cpu > gpurun_out/adv/bench_256_e2e.json 2> gpurun_out/adv/bench_256_e2e.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/adv/bench_256_e2e.json').read().strip().splitlines()[-1]); e=d["e2e"]; print("value %.4g e2e %.4g members %.4g"%(d["value"], e["value"], e["members"]["value"]), e["h2d_bytes_per_step"], e["d2h_bytes_per_step"])
except Exception as ex: print("FAILED", ex); print(open('gpurun_out/adv/bench_256_e2e.err').read()[-1500:])
PY
timeout -k 10 900 python bench.py --size 256 --steps 5 --warmup 3 --no-cpu --e2e-threads 4 > gpurun_out/adv/bench_256_e2e_t4.json 2> gpurun_out/adv/bench_256_e2e_t4.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/adv/bench_256_e2e_t4.json').read().strip().splitlines()[-1]); e=d["e2e"]; print("4 threads: e2e %.4g members %.4g"%(e["value"], e["members"]["value"]))
except Exception as ex: print("FAILED", ex)
PY
