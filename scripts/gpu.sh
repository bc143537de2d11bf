#!/bin/bash
# scripts/gpu.sh <task> [args]  — the GPU-box jobs of this repo, one parametrised script.
# Run as: gpurun --timeout N -- 'scripts/gpu.sh <task> ...'.  Everything is wrapped in its own
# `timeout` so that a hung kernel cannot hold the box.  Output goes to gpurun_out/<tag>/.
set -u
task=${1:-help}; shift || true
tag=${TAG:-$task}
out=gpurun_out/$tag
mkdir -p $out
bench() {  # bench <label> <env...> -- <bench args>
  local label=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout -k 10 900 python bench.py "$@" > $out/bench_$label.json 2> $out/bench_$label.err
  python - "$out/bench_$label.json" "$label" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    e2e = (d.get("e2e") or {}).get("value")
    print(f"{sys.argv[2]:28s} value {d['value']:.4g} ms/step {d['ms_per_step']:.2f} density-launch {r.get('ms_per_launch'):.3f} frac {r.get('frac'):.4f} kernel {r.get('kernel')} e2e {e2e}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
case $task in
  first)     # first contact of a changed kernel: smoke on a short leash, TMA parity, then A/B timing
    if ! timeout -k 5 150 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; then
      echo "SMOKE FAILED OR HUNG (rc $?)"; tail -5 $out/smoke.log; exit 1
    fi
    tail -2 $out/smoke.log
    timeout -k 10 900 python -m pytest tests/test_gpu_tma.py -x -q -m gpu 2>&1 | tail -15 | tee $out/pytest_tma.log
    bench 256_tma X=1 -- --size 256 --steps 5 --warmup 3 --no-cpu --no-e2e
    bench 256_cpasync MXADV_NO_TMA=1 -- --size 256 --steps 5 --warmup 3 --no-cpu --no-e2e
    bench 512_tma X=1 -- --size 512 --steps 5 --warmup 3 --no-cpu --no-e2e
    ;;
  tests)     # the whole GPU suite
    timeout -k 10 ${T:-2400} python -m pytest tests -x -q -m gpu "$@" 2>&1 | tail -25 | tee $out/pytest_gpu.log
    ;;
  ab)        # A/B of experiment builds: scripts/gpu.sh ab <size> <name1> <name2> ...   (exp/<name>/libmxadv.so; "tree" = in-tree)
    size=$1; shift
    for name in "$@"; do
      if [ "$name" = tree ]; then lib=maestroex_b200/csrc/libmxadv.so; else lib=exp/$name/libmxadv.so; fi
      bench ${size}_$name MXADV_LIB=$lib -- --size $size --steps 5 --warmup 3 --no-cpu --no-e2e ${BENCH_ARGS:-}
    done
    ;;
  prof)      # ncu --set full of the density launch: scripts/gpu.sh prof <size> <regex> [env...]
    size=$1; rx=$2; shift 2
    env "$@" timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s ${SKIP:-2} -c 1 \
      -o $out/prof_${tag} -f python bench.py --size $size --steps 1 --warmup 1 --no-cpu --no-e2e > $out/ncu_${tag}.log 2>&1
    tail -3 $out/ncu_${tag}.log
    ;;
  *) echo "tasks: first | tests | ab <size> <names...> | prof <size> <regex>"; exit 2;;
esac
