#!/bin/bash
# N-GPU run (gpurun --gpus N): bench.py under torchrun with and without the NUMA binding; host facts of the box
N=${1:-8}
O=$GRAFT_REPO_ROOT/gpurun_out; mkdir -p $O; cd $GRAFT_REPO_ROOT
{ echo "nproc $(nproc)"; echo "cpu.max $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; free -g | head -2; nvidia-smi topo -m; } > $O/host_${N}gpu.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline --traffic off > $O/scale_N${N}_numa.json 2> $O/scale_N${N}_numa.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline --traffic off --no-numa > $O/scale_N${N}_nonuma.json 2> $O/scale_N${N}_nonuma.err
echo scale done
