# usage: bash scripts/gpu_ncu.sh TAG — launch list of one training iteration + full captures of the dominant GEMM kernels (1 GPU)
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-render --no-graph"
# the timed step of `bench.py --steps 1 --warmup 3` is the 4th iteration; list every launch of the run and cut afterwards
[ -s gpurun_out/launches_$TAG.csv ] || timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_list_$TAG.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"linear_kernel<.int.256, .int.0, .int.2" -s 7 -c 1 -o gpurun_out/prof_fwd_$TAG -f $B > gpurun_out/ncu_fwd_$TAG.log 2>&1; echo "fwd rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"linear_kernel<.int.256, .int.2, .int.2" -s 1 -c 1 -o gpurun_out/prof_fwdhead_$TAG -f $B > gpurun_out/ncu_fwdhead_$TAG.log 2>&1; echo "fwd+head rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"linear_kernel<.int.256, .int.1, .int.2" -s 3 -c 1 -o gpurun_out/prof_dgrad_$TAG -f $B > gpurun_out/ncu_dgrad_$TAG.log 2>&1; echo "dgrad rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"wgrad_kernel<.int.256, .int.2" -s 3 -c 1 -o gpurun_out/prof_wgrad_$TAG -f $B > gpurun_out/ncu_wgrad_$TAG.log 2>&1; echo "wgrad rc=$?"
ls -la gpurun_out/*.ncu-rep
