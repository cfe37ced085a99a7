# usage: bash scripts/gpu_ncu.sh TAG — launch list + full captures of the three dominant GEMM kernels (1 GPU)
TAG=${1:-r01}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 755 -c 800 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_list_$TAG.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k linear_kernel -s 7 -c 1 -o gpurun_out/prof_fwd_$TAG -f $B > gpurun_out/ncu_fwd_$TAG.log 2>&1; echo "fwd rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k linear_kernel -s 52 -c 1 -o gpurun_out/prof_dgrad_$TAG -f $B > gpurun_out/ncu_dgrad_$TAG.log 2>&1; echo "dgrad rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k wgrad_kernel -s 12 -c 1 -o gpurun_out/prof_wgrad_$TAG -f $B > gpurun_out/ncu_wgrad_$TAG.log 2>&1; echo "wgrad rc=$?"
ls -la gpurun_out/*.ncu-rep
