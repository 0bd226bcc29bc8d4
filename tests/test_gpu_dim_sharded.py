"""Dimension-sharded training kernels on ONE GPU: the W ranks of a column-sharded model are played one after the
other through the same C-ABI calls the multi-process driver makes (kge_train_partial -> sum of the ranks' partial
sums -> kge_train_backward -> kge_train_reduce), and the merged result is compared with the oracle's step on the
whole model (1e-5 relative, BASELINE.json north_star).  The NCCL plumbing itself is covered by
tests/test_multi_gpu.py (>= 2 GPUs) and tests/test_distributed_host.py (gloo)."""
import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu


def _dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _close(a, b, rtol=1e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * float(np.abs(b).max()))


def virtual_step(engine, model, k, W, loss, eta, ent, rel, pos, keep=None, repl=None, *, norm=1, opt="adam", lr=1e-3, margin=1.0,
                 flags=0, state=None, step=1, chunks=1, nl=0, seed=0, alpha=0.5, sorted_partial=False):
    """One optimisation step of a model split over W virtual ranks; returns merged tables / gradients / scores."""
    from emgraph_b200 import _lib
    from emgraph_b200 import distributed as D
    from emgraph_b200.engine import model_id
    n = pos.shape[0]
    kc = D.dim_width(k, W)
    pos_d = _dev(pos, torch.int32)
    repl_d = _dev(repl, torch.int32) if repl is not None else None
    keep_d = _dev(keep, torch.uint8) if keep is not None else None
    bounds = D.chunk_bounds(n, chunks)
    ranks = []
    for r in range(W):
        e = _dev(D.slice_columns(ent, model, k, W, r))
        rl = _dev(D.slice_columns(rel, model, k, W, r))
        st = {}
        if state is not None:
            st = {nm: _dev(D.slice_columns(v, model, k, W, r)) for nm, v in state.items()}
        out = dict(loss=torch.zeros(1, device="cuda"), scores=torch.zeros(n * (1 + eta), device="cuda"),
                   g_ent=torch.zeros_like(e), g_rel=torch.zeros_like(rl))
        a = engine.train_args(model=model_id(model, norm), loss=_lib.LOSS_IDS[loss], opt=_lib.OPT_IDS[opt], k=kc, k_model=k, eta=eta,
                              ent=e, rel=rl, pos=pos_d, loss_out=out["loss"], flags=flags, margin=margin, alpha=alpha, lr=lr, step=step,
                              seed=seed, repl=repl_d, keep_subj=keep_d, dbg_scores=out["scores"], dbg_grad_ent=out["g_ent"],
                              dbg_grad_rel=out["g_rel"], non_linearity=nl, **st)
        flat = torch.zeros((1 + eta) * n, device="cuda")
        sums = [flat[(1 + eta) * lo:(1 + eta) * hi] for lo, hi in bounds]
        ranks.append(dict(a=a, ent=e, rel=rl, st=st, out=out, sums=sums, flat=flat))
    # phase 1 on every rank, then the "all-reduce" (fixed rank order), then phase 2 + reduction rank by rank (the ctx-owned
    # corruption / key buffers hold the same values for every rank: same batch, seed and step)
    def phase1(rk):
        if sorted_partial:  # the whole batch at once, entity rows streamed in sorted order, same chunk-major layout
            engine.train_partial_sorted(rk["a"], rk["flat"], len(bounds))
        else:
            for c, (lo, hi) in enumerate(bounds):
                engine.train_partial(rk["a"], rk["sums"][c], lo, hi)

    for rk in ranks:
        phase1(rk)
    totals = []
    for c in range(len(bounds)):
        t = ranks[0]["sums"][c].clone()
        for rk in ranks[1:]:
            t += rk["sums"][c]
        totals.append(t)
    for rk in ranks:
        phase1(rk)  # re-establish this rank's step (emit + sort)
        for c, (lo, hi) in enumerate(bounds):
            engine.train_backward(rk["a"], totals[c], lo, hi)
        engine.train_reduce(rk["a"])
    torch.cuda.synchronize()
    cat = lambda key, sub=None: D.merge_columns([(rk[key] if sub is None else rk[key][sub]).cpu().numpy() for rk in ranks], model, k)
    res = dict(ent=cat("ent"), rel=cat("rel"), g_ent=cat("out", "g_ent"), g_rel=cat("out", "g_rel"),
               loss=[float(rk["out"]["loss"].item()) for rk in ranks], scores=ranks[0]["out"]["scores"].cpu().numpy(),
               scores_all=[rk["out"]["scores"].cpu().numpy() for rk in ranks])
    if state is not None:
        res["state"] = {nm: D.merge_columns([rk["st"][nm].cpu().numpy() for rk in ranks], model, k) for nm in state}
    return res


CASES = [
    ("DistMult", "nll", 256, 64, 8, 1), ("DistMult", "nll", 256, 64, 2, 1), ("ComplEx", "nll", 200, 20, 8, 1),
    ("ComplEx", "pairwise", 200, 20, 4, 1), ("TransE", "pairwise", 100, 20, 8, 1), ("TransE", "multiclass_nll", 100, 20, 2, 2),
    ("HolE", "multiclass_nll", 256, 20, 4, 1), ("HolE", "self_adversarial", 64, 7, 8, 1), ("DistMult", "self_adversarial", 40, 33, 4, 1),
    ("TransE", "absolute_margin", 24, 5, 3, 1), ("DistMult", "pairwise", 10, 3, 4, 1), ("ComplEx", "multiclass_nll", 12, 9, 8, 1),
    ("DistMult", "nll", 1024, 4, 2, 1), ("ComplEx", "nll", 520, 4, 2, 1),
]


@pytest.mark.parametrize("sorted_partial", [False, True], ids=["gather", "sorted"])
@pytest.mark.parametrize("model,loss,k,eta,W,norm", CASES)
def test_dim_sharded_step_vs_oracle(engine, model, loss, k, eta, W, norm, sorted_partial):
    """scores, loss, summed row gradients of the merged slices against the oracle on the whole model (supplied corruptions)."""
    from emgraph_b200 import _lib
    rng = np.random.default_rng(31)
    E, R, n = 2500, 9, 203
    K = ko.internal_k(model, k)
    lim = 0.4 if model != "TransE" else 0.2
    ent = rng.uniform(-lim, lim, size=(E, K)).astype(np.float32)
    rel = rng.uniform(-lim, lim, size=(R, K)).astype(np.float32)
    pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
    keep = rng.integers(0, 2, n * eta).astype(np.uint8)
    repl = rng.integers(0, E, n * eta).astype(np.int32)
    margin = 5.0 if model == "DistMult" else 1.0
    r = virtual_step(engine, model, k, W, loss, eta, ent, rel, pos, keep, repl, norm=norm, margin=margin, flags=_lib.F_NO_UPDATE,
                     chunks=3 if W == 4 else 1, sorted_partial=sorted_partial)
    o = ko.train_step(model, k, loss, eta, ent, rel, pos, keep, repl, margin=margin, norm=norm, dtype=np.float64)
    _close(r["scores"][:n], o["scores_pos"])
    _close(r["scores"][n:], o["scores_neg"])
    for sc in r["scores_all"][1:]:  # every rank evaluates the same scores and the same loss from the same totals
        np.testing.assert_array_equal(sc, r["scores"])
    assert len(set(r["loss"])) == 1
    np.testing.assert_allclose(r["loss"][0], o["loss"], rtol=1e-5)
    rt = 1e-4 if (model == "TransE" and norm == 1) else 1e-5  # L1 sign gradients: exact integers unless a hinge sits on the boundary
    _close(r["g_ent"], o["grad_ent"], rtol=rt)
    _close(r["g_rel"], o["grad_rel"], rtol=rt)
    np.testing.assert_array_equal(r["ent"], ent)  # NO_UPDATE


@pytest.mark.parametrize("model,k,W,opt", [("DistMult", 64, 8, "adam"), ("ComplEx", 20, 4, "adagrad"), ("TransE", 48, 2, "momentum"),
                                           ("DistMult", 36, 4, "sgd")])
def test_dim_sharded_three_stateful_steps(engine, model, k, W, opt):
    """parameters and optimizer state after three steps equal the oracle's (stateful sparse optimizer on every slice)."""
    rng = np.random.default_rng(7)
    E, R, eta, n = 400, 5, 6, 150
    K = ko.internal_k(model, k)
    ent = (rng.normal(size=(E, K)) * 0.4).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.4).astype(np.float32)
    if opt == "adam":
        st = dict(ent_m=np.zeros_like(ent), ent_v=np.zeros_like(ent), rel_m=np.zeros_like(rel), rel_v=np.zeros_like(rel))
    elif opt == "adagrad":
        st = dict(ent_m=np.full_like(ent, 0.1), rel_m=np.full_like(rel, 0.1))
    elif opt == "momentum":
        st = dict(ent_m=np.zeros_like(ent), rel_m=np.zeros_like(rel))
    else:
        st = {}
    e_o, r_o, o_state = ent.copy(), rel.copy(), None
    for step in (1, 2, 3):
        pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
        pos[: n // 3, 0] = 7  # a hub entity: runs that span several chunks of the reduction
        keep = rng.integers(0, 2, n * eta).astype(np.uint8)
        repl = rng.integers(0, E, n * eta).astype(np.int32)
        r = virtual_step(engine, model, k, W, "nll", eta, ent, rel, pos, keep, repl, opt=opt, lr=5e-3, state=st or None, step=step, chunks=2,
                         sorted_partial=(W >= 4))
        o = ko.train_step(model, k, "nll", eta, e_o, r_o, pos, keep, repl, opt=opt, lr=5e-3, state=o_state, step=step)
        ent, rel = r["ent"], r["rel"]
        st = r.get("state", {})
        e_o, r_o, o_state = o["ent_new"], o["rel_new"], (o["state_ent"], o["state_rel"])
        np.testing.assert_allclose(ent, e_o, rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(rel, r_o, rtol=2e-5, atol=2e-6)


def test_dim_sharded_philox_stream_equals_single_gpu(engine):
    """With in-kernel corruptions the sharded step draws the single-GPU stream of the same (global) batch: the merged
    update equals the fused single-GPU step to rounding (different summation order of the score only)."""
    from emgraph_b200 import _lib
    from emgraph_b200.engine import model_id
    rng = np.random.default_rng(3)
    model, k, W, eta, E, R, n = "ComplEx", 24, 4, 8, 700, 6, 260
    K = ko.internal_k(model, k)
    ent = (rng.normal(size=(E, K)) * 0.3).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.3).astype(np.float32)
    pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
    st = dict(ent_m=np.zeros_like(ent), ent_v=np.zeros_like(ent), rel_m=np.zeros_like(rel), rel_v=np.zeros_like(rel))
    r = virtual_step(engine, model, k, W, "nll", eta, ent, rel, pos, opt="adam", lr=1e-2, state=st, step=5, seed=1234)
    e1, r1 = _dev(ent), _dev(rel)
    s1 = {nm: _dev(v) for nm, v in st.items()}
    loss1 = torch.zeros(1, device="cuda")
    sc1 = torch.zeros(n * (1 + eta), device="cuda")
    a = engine.train_args(model=model_id(model), loss=_lib.LOSS_IDS["nll"], opt=0, k=k, eta=eta, ent=e1, rel=r1, pos=_dev(pos, torch.int32),
                          loss_out=loss1, lr=1e-2, step=5, seed=1234, dbg_scores=sc1, **s1)
    engine.train_step(a)
    torch.cuda.synchronize()
    _close(r["scores"], sc1.cpu().numpy())
    np.testing.assert_allclose(r["loss"][0], float(loss1.item()), rtol=1e-5)
    touched = np.abs(e1.cpu().numpy() - ent).max(1) > 0
    np.testing.assert_array_equal(np.abs(r["ent"] - ent).max(1) > 0, touched)
    big = np.abs(s1["ent_m"].cpu().numpy()) > 1e-4  # first Adam step ~ lr*sign(g): compare where the gradient is not ~ 0
    np.testing.assert_allclose(r["ent"][big], e1.cpu().numpy()[big], rtol=1e-5, atol=1e-6)
    _close(r["state"]["ent_m"], s1["ent_m"].cpu().numpy())


@pytest.mark.parametrize("K", [8, 20, 32, 48, 64])
def test_group_reduction_equals_warp_reduction(engine, K, monkeypatch):
    """kge_reduce_apply_group_kernel (a group of lanes per chunk, narrow rows) against the oracle, hubs included; the
    warp-per-chunk kernel runs the same slots in the same order, so both are pinned by the same oracle check."""
    rng = np.random.default_rng(K)
    model, E, R, eta, n = "DistMult", 150, 3, 9, 400
    ent = (rng.normal(size=(E, K)) * 0.4).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.4).astype(np.float32)
    pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
    pos[:150, 2] = 3
    keep = rng.integers(0, 2, n * eta).astype(np.uint8)
    repl = rng.integers(0, E, n * eta).astype(np.int32)
    st = dict(ent_m=np.zeros_like(ent), ent_v=np.zeros_like(ent), rel_m=np.zeros_like(rel), rel_v=np.zeros_like(rel))
    r = virtual_step(engine, model, K, 1, "nll", eta, ent, rel, pos, keep, repl, opt="adam", lr=1e-2, state=st, step=1)
    o = ko.train_step(model, K, "nll", eta, ent, rel, pos, keep, repl, opt="adam", lr=1e-2, step=1,
                      state=((np.zeros_like(ent), np.zeros_like(ent)), (np.zeros_like(rel), np.zeros_like(rel))))
    _close(r["g_ent"], o["grad_ent"])
    _close(r["g_rel"], o["grad_rel"])
    big = np.abs(o["grad_ent"]) > 1e-3
    np.testing.assert_allclose(r["ent"][big], o["ent_new"][big], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(r["ent"][~o["touched_ent"]], ent[~o["touched_ent"]])


@pytest.mark.parametrize("K,opt", [(32, "adam"), (56, "adagrad"), (200, "adam")])
def test_dense_batch_spans_warp_per_run_and_hubs(engine, K, opt):
    """Many more slots than rows (the regime of a small table on many GPUs): nearly every run crosses chunk borders.  The
    warp-per-run span pass finishes the short spans, a hub entity with > 64 chunks of slots goes through hub_list to the
    CTA kernel; gradients and updated rows against the oracle."""
    rng = np.random.default_rng(100 + K)
    model, E, R, eta, n = "DistMult", 60, 3, 4, 3000
    ent = (rng.normal(size=(E, K)) * 0.3).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.3).astype(np.float32)
    pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
    pos[:1500, 0] = 5  # 1500+ slots of one key: ~100 chunks
    keep = rng.integers(0, 2, n * eta).astype(np.uint8)
    repl = rng.integers(0, E, n * eta).astype(np.int32)
    if opt == "adam":
        st = dict(ent_m=np.zeros_like(ent), ent_v=np.zeros_like(ent), rel_m=np.zeros_like(rel), rel_v=np.zeros_like(rel))
        o_state = ((np.zeros_like(ent), np.zeros_like(ent)), (np.zeros_like(rel), np.zeros_like(rel)))
    else:
        st = dict(ent_m=np.full_like(ent, 0.1), rel_m=np.full_like(rel, 0.1))
        o_state = ((np.full_like(ent, 0.1),), (np.full_like(rel, 0.1),))
    r = virtual_step(engine, model, K, 1, "nll", eta, ent, rel, pos, keep, repl, opt=opt, lr=1e-2, state=st, step=1)
    o = ko.train_step(model, K, "nll", eta, ent, rel, pos, keep, repl, opt=opt, lr=1e-2, step=1, state=o_state)
    _close(r["g_ent"], o["grad_ent"], rtol=2e-5)
    _close(r["g_rel"], o["grad_rel"], rtol=2e-5)
    big = np.abs(o["grad_ent"]) > 1e-3
    np.testing.assert_allclose(r["ent"][big], o["ent_new"][big], rtol=1e-5, atol=1e-6)
