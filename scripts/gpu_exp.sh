# usage: gpu_exp.sh "<name:env...>" ...   runs bench 256 (and 512 if BIG=1) for each exp lib
mkdir -p gpurun_out
run() {
  python bench.py --size $1 --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  n=$1', round(d['ms_per_step'],3), round(d['roofline']['ms_per_launch'],3), round(d['roofline']['frac'],4))"
}
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}; [ "$envs" = "$spec" ] && envs=""
  echo "== $name $envs"
  if [ "$name" = base ]; then unset MXADV_LIB; else export MXADV_LIB=/root/repo/exp/$name/libmxadv.so; fi
  env $envs bash -c "$(declare -f run); run 256"
  [ -n "$BIG" ] && env $envs bash -c "$(declare -f run); run 512"
done
