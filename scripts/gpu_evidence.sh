# usage: bash scripts/gpu_evidence.sh TAG — single-GPU evidence of the final code (under gpurun): sanitizer runs, launch list,
# one full ncu capture of the layer-fused proposal kernel
TAG=${1:-r02}
mkdir -p gpurun_out
bash scripts/gpu_sanitize.sh $TAG
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-render --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_list_$TAG.log 2>&1; echo "list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:prop_fused_pair -s 1 -c 1 -f -o gpurun_out/prof_propfused_$TAG python scripts/prop_ab.py > gpurun_out/ncu_propfused_$TAG.log 2>&1; echo "propfused infer rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:prop_fused_pair -s 4 -c 1 -f -o gpurun_out/prof_propfused_train_$TAG python scripts/prop_ab.py > gpurun_out/ncu_propfused_train_$TAG.log 2>&1; echo "propfused train rc=$?"
