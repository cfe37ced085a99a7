TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"resample_rg|bounds_rg|composite_fwd_rg|composite_bwd_rg|norm_sq" -s 5 -c 5 -o gpurun_out/prof_perray_$TAG -f python scripts/ncu_perray.py > gpurun_out/ncu_perray_$TAG.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_perray_$TAG.log
