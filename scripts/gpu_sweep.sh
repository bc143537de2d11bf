mkdir -p gpurun_out
out=gpurun_out/sweep.jsonl; : > $out
run() { python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e "$@" 2>/dev/null >> $out; }
for ppm in 0 1 2; do for ns in 3 13 21; do run --size 256 --ppm $ppm --nspec $ns; done; done
for ppm in 0 1 2; do for ns in 3 13; do run --size 512 --ppm $ppm --nspec $ns; done; done
run --size 256 --ppm 1 --nspec 13 --spherical --bc outflow
run --size 256 --ppm 0 --nspec 13 --spherical --bc outflow
run --size 512 --ppm 1 --nspec 13 --spherical --bc outflow
run --size 256 --ppm 1 --nspec 13 --bc periodic
run --size 256 --ppm 1 --nspec 13 --bc outflow
python - <<'PY'
import json
print("| workload | zone-species/s | ms/step | density ms | frac of HBM peak |")
print("|---|---|---|---|---|")
for l in open("gpurun_out/sweep.jsonl"):
    d=json.loads(l); r=d["roofline"]
    print(f"| {d['config']['workload']} | {d['value']:.3e} | {d['ms_per_step']:.2f} | {r['ms_per_launch']:.2f} | {r['frac']:.3f} |")
PY
