"""Print the per-kernel roofline table bench.py --kernel-table wrote."""
import json
import sys

d = json.load(open(sys.argv[1]))
print(f"ms/step {d['ms_per_step']:.2f}   sum of kernel ms/step {d['kernel_ms_per_step']:.2f}")
for r in d["kernels"][: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{r['kernel'][7:]:24s} {str(r['args']):34s} n/step={r['launches_per_step']:5.1f} avg={r['avg_ms']:8.4f}ms "
          f"share={r['share'] * 100:5.1f}% {r['bound']:6s} {r['achieved']:8.1f} {r['unit']:8s} frac={r['frac']:.3f}")
