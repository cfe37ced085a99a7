mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
P="python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider"
timeout 900 $P -k "not linear and not model and not mutated" > gpurun_out/t1.log 2>&1; echo "t1 rc=$?"
timeout 300 $P -k "linear_fwd" > gpurun_out/t2.log 2>&1; echo "t2 rc=$?"
timeout 300 $P -k "linear_dgrad" > gpurun_out/t3.log 2>&1; echo "t3 rc=$?"
timeout 300 $P -k "linear_wgrad" > gpurun_out/t4.log 2>&1; echo "t4 rc=$?"
timeout 600 $P -k "model or mutated" > gpurun_out/t5.log 2>&1; echo "t5 rc=$?"
for f in t1 t2 t3 t4 t5; do echo "=== $f"; tail -5 gpurun_out/$f.log; done
