"""Row-sharded training + ranking on 2 GPUs against the oracle (needs >= 2 B200s; skipped otherwise).
Run with:  gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu -x -q"""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, cfg, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from emgraph_b200 import _lib
        from emgraph_b200.distributed import ShardedKGE
        model, loss, k, eta, E, R, n = cfg["model"], cfg["loss"], cfg["k"], cfg["eta"], cfg["E"], cfg["R"], cfg["n"]
        ent, rel = cfg["ent"], cfg["rel"]
        sk = ShardedKGE(model, k, eta, loss, "adam", E, R, n, lr=1e-2, init_ent=lambda b, e: ent[b:e], init_rel=lambda: rel,
                        device=rank)
        sk.exchange = cfg["exchange"]
        dev = sk.eng.tdev
        pos = torch.from_numpy(cfg["pos"][rank]).to(dev)
        repl = torch.from_numpy(cfg["repl"][rank]).to(dev)
        keep = torch.from_numpy(cfg["keep"][rank]).to(dev)
        loss_sum = sk.train_step(pos, repl=repl, keep_subj=keep)
        torch.cuda.synchronize()
        ent_new = sk.gather_entities()
        rel_new = sk.rel.cpu().numpy()
        sk.eng.filter_build(torch.from_numpy(cfg["filt"]).to(dev), E, R)
        out = {}
        for tc in (False, True):
            if tc and model == "TransE":
                continue
            r = sk.rank(torch.from_numpy(cfg["test"]).to(dev), side=0, strategy=0, filtered=True, use_tensor_cores=tc)
            out["ranks_tc%d" % int(tc)] = r.cpu().numpy()
        if rank == 0:
            q.put(dict(loss=float(loss_sum.item()), ent=ent_new, rel=rel_new, **out))
        dist.barrier()
    except BaseException as e:  # report instead of letting the parent time out
        import traceback
        q.put(dict(error="rank %d: %s\n%s" % (rank, e, traceback.format_exc())))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("model,loss,k,exchange", [("ComplEx", "nll", 12, "push"), ("TransE", "pairwise", 16, "push"),
                                                   ("DistMult", "multiclass_nll", 8, "push"), ("ComplEx", "nll", 12, "pull"),
                                                   ("TransE", "pairwise", 16, "pull")])
def test_sharded_step_and_ranking_match_oracle(model, loss, k, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, E, R, eta, n = 2, 301, 5, 6, 96
    rng = np.random.default_rng(21)
    K = ko.internal_k(model, k)
    ent = (rng.normal(size=(E, K)) * 0.4).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.4).astype(np.float32)
    pos = [np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32) for _ in range(world)]
    repl = [rng.integers(0, E, eta * n).astype(np.int32) for _ in range(world)]
    keep = [rng.integers(0, 2, eta * n).astype(np.uint8) for _ in range(world)]
    filt = ko.synthetic_triples(E, R, 1500, seed=5)
    test = filt[:40]
    cfg = dict(exchange=exchange, model=model, loss=loss, k=k, eta=eta, E=E, R=R, n=n, ent=ent, rel=rel, pos=pos, repl=repl, keep=keep, filt=filt, test=test)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, cfg, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=150)
    assert "error" not in res, res.get("error")
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # oracle: one step on the concatenated batch; negative (j, i) of rank r sits at row j*N + r*n + i
    N = world * n
    P = np.concatenate(pos, 0)
    RP = np.zeros(eta * N, np.int32)
    KP = np.zeros(eta * N, np.uint8)
    for r in range(world):
        for j in range(eta):
            RP[j * N + r * n:j * N + (r + 1) * n] = repl[r][j * n:(j + 1) * n]
            KP[j * N + r * n:j * N + (r + 1) * n] = keep[r][j * n:(j + 1) * n]
    o = ko.train_step(model, k, loss, eta, ent, rel, P, KP, RP, opt="adam", lr=1e-2,
                      state=((np.zeros_like(ent), np.zeros_like(ent)), (np.zeros_like(rel), np.zeros_like(rel))), step=1)
    np.testing.assert_allclose(res["loss"], o["loss"], rtol=1e-5)
    np.testing.assert_array_equal(res["ent"][~o["touched_ent"]], ent[~o["touched_ent"]])
    big = np.abs(o["grad_ent"]) > 1e-3  # first Adam step ~ lr*sign(g): ill-conditioned where |g| ~ eps
    np.testing.assert_allclose(res["ent"][big], o["ent_new"][big], rtol=1e-5, atol=1e-6)
    bigr = np.abs(o["grad_rel"]) > 1e-3
    np.testing.assert_allclose(res["rel"][bigr], o["rel_new"][bigr], rtol=1e-5, atol=1e-6)
    exp = ko.ranks(model, k, res["ent"], res["rel"], test, filt, "s,o", "worst")
    for key in ("ranks_tc0", "ranks_tc1"):
        if key in res:
            assert res[key].shape == exp.shape and (res[key] != exp).sum() <= 2, key
