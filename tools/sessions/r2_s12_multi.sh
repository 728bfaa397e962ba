#!/bin/bash
# 2-GPU diagnosis: the same slab solve under torchrun and with two hand-started ranks (no elastic agent)
O=gpurun_out
export GSB_SLAB_TRACE=1
echo "== torchrun n=2"
timeout 300 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29502 bench.py --gpus 2 --workload slab --steps 5 2>&1 | grep -E "slab trace|ms_per_vcycle" | tail -4 | cut -c1-220
echo "== hand-started ranks n=2"
for r in 0 1; do
  RANK=$r LOCAL_RANK=$r WORLD_SIZE=2 MASTER_ADDR=127.0.0.1 MASTER_PORT=29503 timeout 300 python bench.py --gpus 2 --workload slab --steps 5 > $O/hand_$r.log 2>&1 &
done
wait
grep -E "slab trace|ms_per_vcycle" $O/hand_0.log | tail -4 | cut -c1-220
echo "== hand-started, 8193"
for r in 0 1; do
  RANK=$r LOCAL_RANK=$r WORLD_SIZE=2 MASTER_ADDR=127.0.0.1 MASTER_PORT=29504 timeout 300 python bench.py --gpus 2 --workload slab --grid 8193 --steps 3 > $O/hand8_$r.log 2>&1 &
done
wait
grep -E "slab trace|ms_per_vcycle" $O/hand8_0.log | tail -3 | cut -c1-220
echo "== env under torchrun"
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 1 --master-port 29505 --no-python env 2>/dev/null | grep -E "OMP|NCCL|TORCH|CUDA|MKL" | head -20
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
