mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_visualize.py tests/test_gpu_parity.py -k "visual or to8b or gpu_normals or gpu_depth" -q -m gpu -p no:cacheprovider > gpurun_out/vis.log 2>&1; echo "vis rc=$?"; tail -60 gpurun_out/vis.log
timeout 600 python scripts/microbench.py visualize --out gpurun_out/r01_visualize.json 2>&1 | tail -30
