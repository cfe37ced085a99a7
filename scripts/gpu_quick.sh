# usage: bash scripts/gpu_quick.sh TAG  — full GPU test suite + one bench line + kernel table
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tests_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernels_$TAG.json > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python scripts/kernel_table.py gpurun_out/kernels_$TAG.json 14; python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print({k: d[k] for k in ('value','ms_per_step','step_tensor_frac','clocks','gpu_launches')}); print(d['e2e'])"; tail -3 gpurun_out/bench_$TAG.err
