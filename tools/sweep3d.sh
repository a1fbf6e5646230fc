#!/bin/bash
# BASELINE configs[4]: VoxelMorph-3D 128^3 batch sweep on one GPU (throughput vs batch), one bench.py line per batch size.
#   gpurun --timeout 1500 -- bash tools/sweep3d.sh
O=gpurun_out/r2_sweep3d.jsonl
: > $O
for B in 1 2 4 8 16 32; do
  python bench.py --workload 3d --batch3d $B --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 >> $O
done
python - <<'PY'
import json
print("# VoxelMorph-3D 128^3 (6-level features), fwd + bwd + Adam, one B200, CUDA-graph step; python bench.py --workload 3d --batch3d B")
print("# batch  ms/step  pairs/s  e2e pairs/s  conv share  hbm_view frac (fwd+dgrad kernels)")
for line in open("gpurun_out/r2_sweep3d.jsonl"):
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print(f"{d['config']['batch_per_gpu']:6d} {d['ms_per_step']:8.2f} {d['value']:8.1f} {d['e2e']['value']:10.1f} "
          f"{d['conv_total']['share_of_step']:10.2f} {d['roofline']['hbm_view']['frac']:8.3f}")
PY
