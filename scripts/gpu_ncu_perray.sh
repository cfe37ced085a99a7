TAG=${1:-x}
mkdir -p gpurun_out
K="resample_rg|bounds_rg|composite_fwd_rg|composite_bwd_rg|norm_sq_rg|level0_t|distortion_rg|interlevel_kernel|cast_ipe_kernel|raygen_kernel|head_bwd_kernel"
# every kernel is launched twice by the driver: skip the first (cold) round
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 13 -c 13 -o gpurun_out/prof_perray_$TAG -f python scripts/ncu_perray.py > gpurun_out/ncu_perray_$TAG.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_perray_$TAG.log
