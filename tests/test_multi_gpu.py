"""Multi-GPU training + ranking against the oracle (needs >= 2 B200s; skipped otherwise).
Run with:  gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu -x -q   (and --gpus 8)

The product path (`ShardedKGE`: column-sharded tables, one all-reduce of partial scores per step, row-range shards for
ranking) runs on 2, 4 and 8 ranks; the round-1 row-sharded exchange (`RowShardedKGE`) keeps one case."""
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
        from emgraph_b200.distributed import RowShardedKGE, ShardedKGE
        model, loss, k, eta, E, R, n = cfg["model"], cfg["loss"], cfg["k"], cfg["eta"], cfg["E"], cfg["R"], cfg["n"]
        ent, rel = cfg["ent"], cfg["rel"]
        out = {}
        if cfg["exchange"] in ("dim", "dim_p2p"):
            # "dim_p2p": the sum over the ranks through the library's peer-memory all-reduce instead of NCCL
            sk = ShardedKGE(model, k, eta, loss, "adam", E, R, n, lr=1e-2, seed=77, init_ent=ent, init_rel=rel, device=rank, chunks=cfg["chunks"],
                            p2p_allreduce=cfg["exchange"] == "dim_p2p")
            assert sk.p2p == (cfg["exchange"] == "dim_p2p")
            dev = sk.eng.tdev
            N = n * world
            P = np.concatenate(cfg["pos"], 0)
            # step 1: supplied corruptions of the GLOBAL batch (parity input); step 2: the in-kernel Philox stream
            l1 = sk.train_step(torch.from_numpy(cfg["pos"][rank]).to(dev), repl=torch.from_numpy(cfg["repl_g"]).to(dev),
                               keep_subj=torch.from_numpy(cfg["keep_g"]).to(dev))
            out["loss1"] = float(l1.item())
            P_dev = torch.from_numpy(P).to(dev)
            torch.cuda.synchronize()  # the pipelined prologue reads the batch from the library's side stream
            l2 = sk.train_step(P_dev, pos_is_global=True)
            out["loss2"] = float(l2.item())
            torch.cuda.synchronize()
            ent_new, rel_new = sk.gather_entities(), sk.gather_relations()
        else:
            sk = RowShardedKGE(model, k, eta, loss, "adam", E, R, n, lr=1e-2, init_ent=lambda b, e: ent[b:e], init_rel=lambda: rel, device=rank)
            sk.exchange = cfg["exchange"]
            dev = sk.eng.tdev
            ls = sk.train_step(torch.from_numpy(cfg["pos"][rank]).to(dev), repl=torch.from_numpy(cfg["repl"][rank]).to(dev),
                               keep_subj=torch.from_numpy(cfg["keep"][rank]).to(dev))
            out["loss1"] = float(ls.item())
            torch.cuda.synchronize()
            ent_new, rel_new = sk.gather_entities(), sk.rel.cpu().numpy()
        sk.eng.filter_build(torch.from_numpy(cfg["filt"]).to(dev), E, R)
        for tc in (False, True):
            if tc and model == "TransE":
                continue
            r = sk.rank(torch.from_numpy(cfg["test"]).to(dev), side=0, strategy=0, filtered=True, use_tensor_cores=tc)
            out["ranks_tc%d" % int(tc)] = r.cpu().numpy()
        if rank == 0:
            q.put(dict(ent=ent_new, rel=rel_new, **out))
        dist.barrier()
    except BaseException as e:  # report instead of letting the parent time out
        import traceback
        q.put(dict(error="rank %d: %s\n%s" % (rank, e, traceback.format_exc())))
        raise
    finally:
        dist.destroy_process_group()


def _run(world, cfg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, cfg, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    assert "error" not in res, res.get("error")
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def _global_corruptions(world, n, eta, repl, keep):
    """negative (j, i) of rank r sits at row j*N + r*n + i of the global batch"""
    N = world * n
    RP, KP = np.zeros(eta * N, np.int32), np.zeros(eta * N, np.uint8)
    for r in range(world):
        for j in range(eta):
            RP[j * N + r * n:j * N + (r + 1) * n] = repl[r][j * n:(j + 1) * n]
            KP[j * N + r * n:j * N + (r + 1) * n] = keep[r][j * n:(j + 1) * n]
    return RP, KP


@pytest.mark.parametrize("world,model,loss,k,exchange,chunks", [
    (2, "ComplEx", "nll", 12, "dim", 2), (2, "TransE", "pairwise", 16, "dim", 1), (2, "DistMult", "multiclass_nll", 8, "dim", 3),
    (2, "HolE", "self_adversarial", 10, "dim", 2), (4, "DistMult", "nll", 64, "dim", 2), (8, "ComplEx", "nll", 100, "dim", 2),
    (8, "DistMult", "nll", 256, "dim", 2), (8, "TransE", "multiclass_nll", 20, "dim", 1),
    (2, "ComplEx", "nll", 12, "push", 1), (2, "DistMult", "nll", 20, "dim_p2p", 2), (8, "ComplEx", "multiclass_nll", 24, "dim_p2p", 2),
])
def test_sharded_step_and_ranking_match_oracle(world, model, loss, k, exchange, chunks):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    E, R, eta, n = 301, 5, 6, 96
    rng = np.random.default_rng(21)
    K = ko.internal_k(model, k)
    ent = (rng.normal(size=(E, K)) * 0.4).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.4).astype(np.float32)
    pos = [np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32) for _ in range(world)]
    repl = [rng.integers(0, E, eta * n).astype(np.int32) for _ in range(world)]
    keep = [rng.integers(0, 2, eta * n).astype(np.uint8) for _ in range(world)]
    RP, KP = _global_corruptions(world, n, eta, repl, keep)
    filt = ko.synthetic_triples(E, R, 1500, seed=5)
    test = filt[:40]
    cfg = dict(exchange=exchange, chunks=chunks, model=model, loss=loss, k=k, eta=eta, E=E, R=R, n=n, ent=ent, rel=rel, pos=pos, repl=repl,
               keep=keep, repl_g=RP, keep_g=KP, filt=filt, test=test)
    res = _run(world, cfg)
    # oracle: the same steps on the concatenated batch
    P = np.concatenate(pos, 0)
    z = lambda: ((np.zeros_like(ent), np.zeros_like(ent)), (np.zeros_like(rel), np.zeros_like(rel)))  # noqa: E731
    o = ko.train_step(model, k, loss, eta, ent, rel, P, KP, RP, opt="adam", lr=1e-2, state=z(), step=1)
    np.testing.assert_allclose(res["loss1"], o["loss"], rtol=1e-5)
    touched = o["touched_ent"].copy()
    big = np.abs(o["grad_ent"]) > 1e-3  # first Adam step ~ lr*sign(g): ill-conditioned where |g| ~ eps
    e_exp, r_exp = o["ent_new"], o["rel_new"]
    if exchange.startswith("dim"):
        repl2, keep2 = ko.draw_corruptions(77, 2, P.shape[0], eta, E, "s,o")
        o2 = ko.train_step(model, k, loss, eta, o["ent_new"], o["rel_new"], P, keep2, repl2, opt="adam", lr=1e-2,
                           state=(o["state_ent"], o["state_rel"]), step=2)
        np.testing.assert_allclose(res["loss2"], o2["loss"], rtol=2e-5)
        touched |= o2["touched_ent"]
        big &= np.abs(o2["grad_ent"]) > 1e-3
        e_exp, r_exp = o2["ent_new"], o2["rel_new"]
    np.testing.assert_array_equal(res["ent"][~touched], ent[~touched])
    np.testing.assert_allclose(res["ent"][big], e_exp[big], rtol=1e-4, atol=2e-6)
    bigr = np.abs(o["grad_rel"]) > 1e-3
    if not exchange.startswith("dim"):
        np.testing.assert_allclose(res["rel"][bigr], r_exp[bigr], rtol=1e-5, atol=1e-6)
    exp = ko.ranks(model, k, res["ent"], res["rel"], test, filt, "s,o", "worst")
    for key in ("ranks_tc0", "ranks_tc1"):
        if key in res:
            assert res[key].shape == exp.shape and (res[key] != exp).sum() <= 2, key


@pytest.mark.parametrize("model,opt", [("ComplEx", "adam"), ("TransE", "adagrad")])
def test_fit_n_gpus_equals_single_gpu_fit_and_resumes(model, opt, tmp_path):
    """fit(engine_params={'n_gpus': 2}) from a plain process (spawns one worker per GPU) trains the model the single-GPU
    fit trains -- same batches, same corruption stream, same updates -- and 2 + 2 epochs through save / restore /
    resume equal 4 epochs (reference API: models/EmbeddingModel.py:1113, utils/model_utils.py:63-87, :139-154)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from emgraph_b200 import models, utils
    from toy_graph import TOY_QUERY
    rng = np.random.default_rng(4)
    tri = ko.synthetic_triples(60, 4, 500, seed=2)
    X = np.stack([np.char.add("e", tri[:, 0].astype(str)), np.char.add("r", tri[:, 1].astype(str)), np.char.add("e", tri[:, 2].astype(str))], 1)
    kw = dict(k=12, eta=4, batches_count=3, seed=9, optimizer=opt, optimizer_params={"lr": 0.02}, loss="nll")
    cls = getattr(models, model)
    m1 = cls(epochs=4, **kw).fit(X)
    m2 = cls(epochs=4, engine_params={"n_gpus": 2}, **kw).fit(X)
    np.testing.assert_allclose(m2.trained_model_params[0], m1.trained_model_params[0], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(m2.predict(X[:40]), m1.predict(X[:40]), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(m2.loss_history, m1.loss_history, rtol=1e-5)
    a = cls(epochs=2, engine_params={"n_gpus": 2}, **kw).fit(X)
    path = str(tmp_path / "m.pkl")
    utils.save_model(a, path, save_optimizer_state=True)
    b = utils.restore_model(path)
    b.engine_params = {"n_gpus": 2, "resume": True}
    b.fit(X)
    np.testing.assert_allclose(b.trained_model_params[0], m2.trained_model_params[0], rtol=1e-5, atol=1e-6)
    # a model trained on 2 GPUs evaluates like the single-GPU one
    from emgraph_b200.evaluation import evaluate_performance
    r1 = evaluate_performance(X[:30], m1, filter_triples=X, corrupt_side="s,o")
    r2 = evaluate_performance(X[:30], m2, filter_triples=X, corrupt_side="s,o")
    assert (r1 != r2).sum() <= 2
