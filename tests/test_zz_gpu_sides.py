"""GPU: a LIST-valued corrupt_side summed into ONE optimizer step (reference models/EmbeddingModel.py:780-816).

The math of the stacked batch is pinned by tests/golden/train_multiside_*.npz (reference-executed; consumed by
tests/test_gpu_parity.py::test_train_step_vs_reference_golden).  Here: the per-negative side codes of
kge_train_args.keep_subj WITHOUT supplied replacements (in-kernel Philox draws), and the host wiring in fit()."""
import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu


def _step(engine, ent, rel, pos, eta, k, side, keep=None, seed=7, step=3, loss="nll", model="ComplEx", opt="sgd", lr=1e-2):
    from emgraph_b200 import _lib
    from emgraph_b200.engine import model_id
    n = pos.shape[0]
    ent_d, rel_d = torch.from_numpy(ent).cuda(), torch.from_numpy(rel).cuda()
    out = dict(loss=torch.zeros(1, device="cuda"), scores=torch.zeros(n * (1 + eta), device="cuda"),
               g_ent=torch.zeros_like(ent_d), g_rel=torch.zeros_like(rel_d))
    a = engine.train_args(model=model_id(model), loss=_lib.LOSS_IDS[loss], opt=_lib.OPT_IDS[opt], k=k, eta=eta, ent=ent_d, rel=rel_d,
                          pos=torch.from_numpy(pos).cuda(), loss_out=out["loss"], side=_lib.TRAIN_SIDE_IDS[side], lr=lr,
                          seed=seed, step=step, keep_subj=None if keep is None else torch.from_numpy(keep).cuda(),
                          dbg_scores=out["scores"], dbg_grad_ent=out["g_ent"], dbg_grad_rel=out["g_rel"])
    engine.train_step(a)
    torch.cuda.synchronize()
    res = {k_: v.cpu().numpy() for k_, v in out.items()}
    res["ent"], res["rel"] = ent_d.cpu().numpy(), rel_d.cpu().numpy()
    return res


def _case(seed=0, E=300, R=5, k=8, n=96):
    rng = np.random.default_rng(seed)
    K = ko.internal_k("ComplEx", k)
    ent = (rng.normal(size=(E, K)) * 0.4).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.4).astype(np.float32)
    pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
    return ent, rel, pos, k


def test_keep_codes_fix_the_side_of_in_kernel_corruptions(engine):
    """keep_subj codes with Philox-drawn replacements: all 1 == side 'o', all 0 == side 's', all 2 == side 's,o';
    the replacement stream does not depend on the side, so the results are bit-identical."""
    ent, rel, pos, k = _case()
    eta, n = 5, pos.shape[0]
    for side, code in (("o", 1), ("s", 0), ("s,o", 2)):
        ref = _step(engine, ent, rel, pos, eta, k, side)
        got = _step(engine, ent, rel, pos, eta, k, "s,o", keep=np.full(n * eta, code, np.uint8))
        for key in ("loss", "scores", "g_ent", "g_rel", "ent", "rel"):
            np.testing.assert_array_equal(got[key], ref[key], err_msg="%s %s" % (side, key))
    # the sides really differ
    a, b = _step(engine, ent, rel, pos, eta, k, "o"), _step(engine, ent, rel, pos, eta, k, "s")
    assert not np.array_equal(a["scores"][n:], b["scores"][n:])


def test_in_kernel_corruption_stream_matches_the_oracle(engine):
    """The Philox stream of kge_emit_kernel restated in the oracle (oracle/kge_oracle.py:draw_corruptions): a step
    with in-kernel corruptions -- sides 's,o' / 's' / 'o' and mixed per-negative codes -- equals the oracle's step
    on the corruptions the oracle draws for the same (seed, step)."""
    ent, rel, pos, k = _case(seed=1)
    eta, n = 4, pos.shape[0]
    E = ent.shape[0]
    mixed = np.tile(np.concatenate([np.zeros(n // 3, np.uint8), np.ones(n // 3, np.uint8), np.full(n - 2 * (n // 3), 2, np.uint8)]), eta)
    for side, codes, seed, step in (("s,o", None, 7, 3), ("s", None, 8, 1), ("o", None, 2**40 + 5, 2**33 + 1), ("s,o", mixed, 9, 12)):
        got = _step(engine, ent, rel, pos, eta, k, side, keep=codes, seed=seed, step=step, loss="pairwise")
        repl, keep = ko.draw_corruptions(seed, step, n, eta, E, side, keep_codes=codes)
        o = ko.train_step("ComplEx", k, "pairwise", eta, ent, rel, pos, keep, repl)
        np.testing.assert_allclose(got["scores"][n:], o["scores_neg"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(got["loss"][0], o["loss"], rtol=1e-5)
        sc = max(1.0, np.abs(o["grad_ent"]).max())
        np.testing.assert_allclose(got["g_ent"], o["grad_ent"], rtol=1e-4, atol=1e-5 * sc)
        np.testing.assert_allclose(got["g_rel"], o["grad_rel"], rtol=1e-4, atol=1e-5 * max(1.0, np.abs(o["grad_rel"]).max()))


@pytest.mark.parametrize("host_batches", [False, True])
def test_fit_with_a_list_of_sides_is_one_step_per_batch(engine, host_batches):
    """fit(corrupt_side=['s','o']) == the engine stepped by hand on the stacked batch (same seed and counters)."""
    from emgraph_b200 import models
    ent, rel, pos, k = _case(seed=2, E=120, n=90)
    E, R = ent.shape[0], rel.shape[0]
    # make every entity / relation id appear so that the label index is the identity
    cover = np.stack([np.arange(E), np.arange(E) % R, (np.arange(E) + 1) % E], 1).astype(np.int32)
    pos = np.concatenate([pos, cover])
    X = np.empty(pos.shape, dtype=object)
    X[:, 0] = ["e%05d" % v for v in pos[:, 0]]
    X[:, 1] = ["r%03d" % v for v in pos[:, 1]]
    X[:, 2] = ["e%05d" % v for v in pos[:, 2]]
    eta, bc = 3, 2
    m = models.ComplEx(k=k, eta=eta, epochs=2, batches_count=bc, seed=11, optimizer="sgd", optimizer_params={"lr": 1e-2},
                       loss="multiclass_nll", initializer="constant", initializer_params={"entity": ent, "relation": rel},
                       embedding_model_params={"corrupt_side": ["s", "o"]}, engine_params={"host_batches": host_batches})
    m.fit(X.astype(str))
    assert m._opt_step == 2 * bc  # one optimizer step per batch, not one per side
    assert len(m.loss_history) == 2 and np.all(np.isfinite(m.loss_history))
    # by hand
    e_h, r_h = ent.copy(), rel.copy()
    N = pos.shape[0]
    bs = int(np.ceil(N / bc))
    step = 0
    for _ in range(2):
        for b in range(bc):
            p = pos[b * bs:min(N, (b + 1) * bs)]
            n = p.shape[0]
            step += 1
            keep = np.tile(np.repeat(np.array([0, 1], np.uint8), n), eta)
            r = _step(engine, e_h, r_h, np.tile(p, (2, 1)), eta, k, "s,o", keep=keep, seed=11, step=step, loss="multiclass_nll")
            e_h, r_h = r["ent"], r["rel"]
    np.testing.assert_array_equal(m.trained_model_params[0], e_h)
    np.testing.assert_array_equal(m.trained_model_params[1], r_h)


def test_select_best_model_ranking_on_the_engine(engine):
    """Grid search end to end on the GPU engine (reference tests/emgraph/evaluation/test_protocol.py:1046-1094):
    every configuration is trained and ranked, the best validation MRR wins."""
    from emgraph_b200 import models
    from emgraph_b200.evaluation import select_best_model_ranking
    E, R = 60, 3
    tri = ko.synthetic_triples(E, R, 900, seed=5)
    X = np.empty(tri.shape, dtype=object)
    X[:, 0] = ["e%03d" % v for v in tri[:, 0]]
    X[:, 1] = ["r%d" % v for v in tri[:, 1]]
    X[:, 2] = ["e%03d" % v for v in tri[:, 2]]
    X = X.astype(str)
    Xtr, Xva, Xte = X[:700], X[700:800], X[800:]
    grid = {"batches_count": [4], "seed": 0, "epochs": [30], "k": [4, 16], "eta": [2], "loss": ["nll"], "loss_params": {},
            "embedding_model_params": {}, "regularizer": [None], "regularizer_params": {}, "optimizer": ["adam"],
            "optimizer_params": {"lr": [1e-9, 5e-2]}}
    best, params, mrr_valid, ranks_test, res, hist = select_best_model_ranking(models.DistMult, Xtr, Xva, Xte, grid)
    assert len(hist) == 4 and all("mrr" in h["results"] for h in hist)
    assert params["optimizer_params"]["lr"] in (1e-9, 5e-2) and isinstance(best, models.DistMult) and best.is_fitted
    assert mrr_valid == max(h["results"]["mrr"] for h in hist)
    assert set(res) == {"mrr", "mr", "hits_1", "hits_3", "hits_10"} and all(np.isfinite(v) and v >= 0 for v in res.values())
    assert ranks_test.shape == (Xte.shape[0], 2) or ranks_test.shape[1] == 2  # unseen-entity triples are filtered out
    # random search with a callable
    grid2 = {"batches_count": [4], "epochs": [5], "k": [4, 8], "eta": [2], "optimizer_params": {"lr": lambda: float(np.random.uniform(1e-3, 1e-2))}}
    out = select_best_model_ranking(models.TransE, Xtr, Xva, Xte, grid2, max_combinations=3, use_filter=False, corrupt_side="o")
    assert len(out[5]) == 3 and out[3].ndim == 1


@pytest.mark.parametrize("model,loss,opt", [("ComplEx", "nll", "adam"), ("TransE", "pairwise", "adagrad"), ("DistMult", "multiclass_nll", "adam")])
@pytest.mark.parametrize("host_batches", [False, True])
def test_pipelined_steps_are_bit_identical_to_in_order_steps(engine, model, loss, opt, host_batches):
    """KGE_F_PIPELINE (emit + sort of step t+1 on the side stream beside step t, two buffer sets, main-part graphs
    for host batches) must not change a single bit: same kernels, same inputs, same summation order.  Hub-heavy
    Zipf batches of two sizes (the last batch is short) over several epochs, device and host batches."""
    from emgraph_b200 import models
    E, R, N = 500, 6, 5000
    tri = ko.synthetic_triples(E, R, N, seed=21, zipf=True)
    X = np.empty(tri.shape, dtype=object)
    X[:, 0] = ["e%04d" % v for v in tri[:, 0]]
    X[:, 1] = ["r%d" % v for v in tri[:, 1]]
    X[:, 2] = ["e%04d" % v for v in tri[:, 2]]
    X = X.astype(str)
    out = []
    for pipeline in (False, True):
        m = getattr(models, model)(k=12, eta=5, epochs=4, batches_count=7, seed=3, optimizer=opt, optimizer_params={"lr": 1e-2},
                                   loss=loss, engine_params={"pipeline": pipeline, "host_batches": host_batches})
        m.fit(X)
        out.append((m.trained_model_params[0].copy(), m.trained_model_params[1].copy(), list(m.loss_history)))
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])
    assert out[0][2] == out[1][2]
    assert np.all(np.isfinite(out[0][2])) and out[0][2][-1] < out[0][2][0]


def test_pipelined_and_in_order_steps_interleave(engine):
    """Alternating pipelined and in-order calls on one ctx (buffer-set bookkeeping) == all in-order."""
    from emgraph_b200 import _lib
    from emgraph_b200.engine import model_id
    ent, rel, pos, k = _case(seed=4, E=200, n=160)
    eta = 6
    res = []
    for pattern in ([0] * 8, [1, 0, 1, 1, 0, 0, 1, 1]):
        ent_d, rel_d = torch.from_numpy(ent).cuda(), torch.from_numpy(rel).cuda()
        m_d, v_d = torch.zeros_like(ent_d), torch.zeros_like(ent_d)
        rm, rv = torch.zeros_like(rel_d), torch.zeros_like(rel_d)
        pos_d = torch.from_numpy(pos).cuda()
        loss_d = torch.zeros(1, device="cuda")
        torch.cuda.synchronize()
        losses = []
        for step, pipe in enumerate(pattern, 1):
            lo = (step % 3) * 40
            a = engine.train_args(model=model_id("ComplEx"), loss=_lib.LOSS_IDS["nll"], opt=_lib.OPT_IDS["adam"], k=k, eta=eta,
                                  ent=ent_d, rel=rel_d, ent_m=m_d, ent_v=v_d, rel_m=rm, rel_v=rv, pos=pos_d[lo:lo + 80 + 8 * (step % 2)],
                                  loss_out=loss_d, lr=1e-2, seed=5, step=step, flags=_lib.F_PIPELINE if pipe else 0)
            engine.train_step(a)
            losses.append(loss_d.clone())
        torch.cuda.synchronize()
        res.append((ent_d.cpu().numpy(), rel_d.cpu().numpy(), [float(x) for x in losses]))
    np.testing.assert_array_equal(res[0][0], res[1][0])
    np.testing.assert_array_equal(res[0][1], res[1][1])
    assert res[0][2] == res[1][2]
