mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_visualize.py -q -m gpu -p no:cacheprovider > gpurun_out/vis.log 2>&1; echo "vis rc=$?"; tail -60 gpurun_out/vis.log
