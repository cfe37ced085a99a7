# usage: bash scripts/gpu_final.sh TAG — the full single-GPU evidence set of a round (under gpurun)
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/tests_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 --kernel-table gpurun_out/kernels_$TAG.json > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python scripts/kernel_table.py gpurun_out/kernels_$TAG.json 8
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref_$TAG.json
timeout 600 python scripts/gemm_bench.py > gpurun_out/gemm_bench_$TAG.log 2>&1; echo "gemm rc=$?"
timeout 1200 python scripts/microbench.py micro --out gpurun_out/micro_$TAG.json > gpurun_out/micro_$TAG.log 2>&1; echo "micro rc=$?"
timeout 900 python scripts/microbench.py render --out gpurun_out/render_$TAG.json > gpurun_out/render_$TAG.log 2>&1; echo "render rc=$?"; grep case gpurun_out/render_$TAG.log | cut -c1-160
bash scripts/gpu_ncu.sh $TAG
