mkdir -p gpurun_out
for ms in 100 1000 100 1000; do
  MIP360_SMI_MS=$ms timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/smi_$ms.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/smi_$ms.json')); print('$ms', d['ms_per_step'], d.get("ms_per_step_with_kernel_events"), d['clocks'])"
done
