#!/bin/bash
# 8-GPU A/B of the exchange on cfg5 (train only): peer-memory all-reduce (grid sizes) vs NCCL, pieces per step
O=gpurun_out; mkdir -p $O; T=${1:-r2r}; N=${2:-8}
run() { name=$1; shift; ch=$1; shift
  env "$@" NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload cfg5 --steps 10 --warmup 3 --no-cpu --no-rank --no-sub --chunks $ch > $O/${T}_$name.json 2> $O/${T}_$name.err
  python -c "
import json
d=json.loads(open('$O/${T}_$name.json').read().strip().splitlines()[-1])
print('$name', 'ms/step %.4f warm %.4f e2e %.4f'%(d['ms_per_step'], d['ms_per_step_warm'], d['e2e']['ms_per_step']), {k:round(v,3) for k,v in d['phases_ms_max_over_ranks'].items()})" 2>/dev/null || tail -3 $O/${T}_$name.err
}
run p2p16_c2 2 KGE_P2P_CTAS=16
run nccl_c2 2 KGE_P2P_ALLREDUCE=0
run p2p8_c2 2 KGE_P2P_CTAS=8
run p2p16_c3 3 KGE_P2P_CTAS=16
run p2p16_c1 1 KGE_P2P_CTAS=16
