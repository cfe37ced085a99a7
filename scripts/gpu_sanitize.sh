# usage: bash scripts/gpu_sanitize.sh TAG — compute-sanitizer memcheck over the GPU test suite and racecheck over the kernels
# that communicate through shared memory / mbarriers (GEMMs incl. the CTA-pair protocol, scans, searches)
TAG=${1:-r02}
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $SAN --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer_memcheck_$TAG.log \
  python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "not full_size and not two_ranks and not default_model and not train_loop_shape and not cuda_graph" \
  > gpurun_out/sanitizer_memcheck_pytest_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_pytest_$TAG.log; tail -4 gpurun_out/sanitizer_memcheck_$TAG.log
timeout 1500 $SAN --tool racecheck --error-exitcode 9 --log-file gpurun_out/sanitizer_racecheck_$TAG.log \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -m gpu -p no:cacheprovider -x -k "linear or resample_vs or composite_fwd or losses_vs or cast_ipe_vs or layer_fused or fused_head" \
  > gpurun_out/sanitizer_racecheck_pytest_$TAG.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_pytest_$TAG.log; tail -4 gpurun_out/sanitizer_racecheck_$TAG.log
