# usage: bash scripts/gpu_bench.sh [tag]   (run under gpurun; writes into gpurun_out/)
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 --kernel-table gpurun_out/kernels_$TAG.json > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
# launch list (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu list rc=$?"
# full capture of the dominant kernels (fwd/dgrad linear and wgrad), 2 launches each, from the steady state
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 60 -c 2 -o gpurun_out/prof_linear_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_linear_$TAG.log 2>&1; echo "ncu linear rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 20 -c 2 -o gpurun_out/prof_wgrad_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_wgrad_$TAG.log 2>&1; echo "ncu wgrad rc=$?"
ls -la gpurun_out | tail -20
