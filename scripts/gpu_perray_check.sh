TAG=${1:-x}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tests_$TAG.log
timeout 1200 python scripts/microbench.py micro --out gpurun_out/micro_$TAG.json > gpurun_out/micro_$TAG.log 2>&1; echo "micro rc=$?"
grep "B=  4194304" gpurun_out/micro_$TAG.log
