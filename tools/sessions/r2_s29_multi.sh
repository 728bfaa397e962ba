#!/bin/bash
# 8-GPU session after the sweep rewrite: slab scaling of 4097^2 / 8193^2 on 8 / 4 / 2 ranks (peer-memory transports, graphs)
O=gpurun_out/s29; mkdir -p $O
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  [ $n -le $N ] || continue
  for g in 8193 4097; do
    timeout 150 $TR --nproc-per-node $n --master-port $((29500+n+g%7)) bench.py --gpus $n --workload slab --grid $g --steps 5 > $O/r2_slab_n${n}_g${g}.json 2> $O/r2_slab_n${n}_g${g}.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/r2_slab_n${n}_g${g}.json").read().strip().splitlines()[-1])
    print("slab n=$n g=$g", "ms/vcycle", round(d["ms_per_vcycle"],3), "glups", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), d["config"]["halo_transport"], "graph", d["config"]["cuda_graph"], "cycles", d["config"]["cycles"], d["config"]["residual_linf"])
except Exception as e:
    print("slab n=$n g=$g FAILED", e)
PY
  done
done
