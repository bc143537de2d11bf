# scratch entry point of `gpurun -- 'bash scripts/_run.sh'`; the parametrised jobs are in scripts/gpu.sh
timeout -k 10 2400 python -m pytest tests -q -m gpu -rs 2>&1 | tail -12
