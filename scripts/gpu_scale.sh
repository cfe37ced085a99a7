# usage: bash scripts/gpu_scale.sh N TAG  (under gpurun --gpus N): train bench + sharded render on N GPUs
N=${1:-2}; TAG=${2:-r01}
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $R bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "bench rc=$?"
tail -n 1 gpurun_out/bench_n${N}_$TAG.json | cut -c1-260
timeout 900 $R scripts/microbench.py render --out gpurun_out/render_n${N}_$TAG.json > gpurun_out/render_n${N}_$TAG.log 2>&1; echo "render rc=$?"
grep "case" gpurun_out/render_n${N}_$TAG.log | cut -c1-220
