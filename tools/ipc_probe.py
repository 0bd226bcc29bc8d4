"""2-process probe: row gathers through CUDA-IPC peer mappings (kge_score on a sharded table).
torchrun --nproc-per-node 2 tools/ipc_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from emgraph_b200.engine import get_engine, make_table
from emgraph_b200.distributed import PeerBuffer, exchange_peers

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = get_engine(local); dev = eng.tdev
K = 256
for rps in (16384, 262144, 2300000):
    buf = PeerBuffer(eng, (rps, K)); buf.tensor.uniform_(-0.1, 0.1)
    exchange_peers(eng, [buf], rank, world)
    E = rps * world
    tab = make_table(buf.peers, rows=E, rows_per_shard=rps, K=K)
    rel = torch.rand((8, K), device=dev)
    n = 300000
    g = torch.Generator(device=dev).manual_seed(rank)
    for name, lo in (("local", rank * rps), ("peer", ((rank + 1) % world) * rps)):
        tri = torch.stack([torch.randint(lo, lo + rps, (n,), device=dev, generator=g), torch.randint(0, 8, (n,), device=dev, generator=g),
                           torch.randint(lo, lo + rps, (n,), device=dev, generator=g)], 1).to(torch.int32).contiguous()
        dist.barrier(); torch.cuda.synchronize()
        best = 1e9
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.score(2, K, tab, rel, tri); e1.record(); torch.cuda.synchronize()
            if it: best = min(best, e0.elapsed_time(e1))
        if rank == 0:
            print("shard %.2f GB %s rows: %.3f ms, %.0f GB/s of entity rows" % (rps * K * 4 / 1e9, name, best, 2 * n * K * 4 / best / 1e6), flush=True)
    dist.barrier()
    del tab; buf.free()
dist.destroy_process_group()
