"""GPU: the reference-facing Python API on top of the C ABI -- fit/predict/evaluate_performance with
entities_subset, early stopping, LP regulariser, negative_corruption_entities, save/restore
 -- checked against the oracle on the same inputs."""
import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu


def _labels(tri):
    """int ids -> string labels whose sorted order is the id order (e0007 ...)."""
    X = np.empty(tri.shape, dtype=object)
    X[:, 0] = ["e%05d" % v for v in tri[:, 0]]
    X[:, 1] = ["r%03d" % v for v in tri[:, 1]]
    X[:, 2] = ["e%05d" % v for v in tri[:, 2]]
    return X.astype(str)


def _fitted(model_cls, k, E, R, ent, rel, X, **kw):
    """A model whose parameters are injected (initializer='constant') and training is a no-op (lr=0 sgd)."""
    m = model_cls(k=k, eta=2, epochs=1, batches_count=1, seed=0, optimizer="sgd", optimizer_params={"lr": 0.0},
                  loss="nll", initializer="constant", initializer_params={"entity": ent, "relation": rel}, **kw)
    m.fit(X)
    assert len(m.ent_to_idx) == E and len(m.rel_to_idx) == R
    return m


@pytest.mark.parametrize("model", ["DistMult", "TransE", "ComplEx"])
def test_entities_subset_ranking_matches_oracle(engine, model):
    from emgraph_b200 import models
    from emgraph_b200.evaluation import evaluate_performance
    rng = np.random.default_rng(31)
    E, R, k = 180, 4, 12
    K = ko.internal_k(model, k)
    ent = (rng.normal(size=(E, K)) * 0.6).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.6).astype(np.float32)
    tri = ko.synthetic_triples(E, R, 1200, seed=8)
    X = _labels(tri)
    m = _fitted(models.MODEL_REGISTRY[model], k, E, R, ent, rel, X)
    np.testing.assert_array_equal(m.trained_model_params[0], ent)  # lr = 0: nothing moved
    test = tri[rng.permutation(len(tri))[:50]]
    subset_ids = np.sort(rng.permutation(E)[:47])
    subset_labels = ["e%05d" % v for v in subset_ids] + ["not-an-entity"]
    for side in ("s,o", "s+o", "o", "s"):
        for strat in ("worst", "middle", "best"):
            for filt in (None, tri):
                got = evaluate_performance(_labels(test), m, filter_triples=None if filt is None else X, entities_subset=subset_labels,
                                           corrupt_side=side, ranking_strategy=strat)
                exp = ko.ranks(model, k, ent, rel, test, filt, side, strat, subset=subset_ids)
                assert got.shape == exp.shape
                assert (got != exp).sum() <= max(1, exp.size // 50), (side, strat, filt is not None, got[:5], exp[:5])
                assert got.max() <= 2 * len(subset_ids) + 1
    # and the full sweep afterwards still uses the un-permuted filter
    got = evaluate_performance(_labels(test), m, filter_triples=X, corrupt_side="s,o")
    exp = ko.ranks(model, k, ent, rel, test, tri, "s,o", "worst")
    assert (got != exp).sum() <= 2


def test_early_stopping_keeps_best_parameters(engine):
    from emgraph_b200.evaluation import evaluate_performance, mrr_score
    from emgraph_b200.models import ComplEx
    rng = np.random.default_rng(2)
    E, R = 120, 3
    tri = ko.synthetic_triples(E, R, 900, seed=3)
    X = _labels(tri)
    valid = X[rng.permutation(len(X))[:60]]
    # a large learning rate so that validation MRR peaks early and then degrades / plateaus
    m = ComplEx(k=16, eta=5, epochs=60, batches_count=3, seed=1, optimizer="adam", optimizer_params={"lr": 0.05}, loss="nll")
    m.fit(X, early_stopping=True,
          early_stopping_params={"x_valid": valid, "x_filter": X, "criteria": "mrr", "burn_in": 2, "check_interval": 2, "stop_interval": 2})
    hist = m.early_stopping_history
    assert len(hist) >= 2 and all(e % 2 == 0 and e >= 2 for e, _ in hist)
    if m.early_stopping_epoch is not None:
        assert m.early_stopping_epoch < 60 and len(m.loss_history) == m.early_stopping_epoch
        # stopped after `stop_interval` checks without improvement
        vals = [v for _, v in hist]
        assert max(vals) == pytest.approx(m.early_stopping_best_value)
        assert vals[-1] <= max(vals) and vals[-2] <= max(vals)
    # the kept parameters reproduce the best validation value
    ranks = evaluate_performance(valid, m, filter_triples=X)
    assert mrr_score(ranks) == pytest.approx(m.early_stopping_best_value, rel=1e-6)
    assert not m.is_filtered and m.eval_config == {}
    with pytest.raises(KeyError):
        ComplEx(k=4, epochs=1, batches_count=1).fit(X, early_stopping=True, early_stopping_params={})


def test_negative_corruption_entities_options(engine):
    from emgraph_b200 import _lib
    from emgraph_b200.models import DistMult
    rng = np.random.default_rng(5)
    E, R, n, eta, k = 400, 3, 512, 8, 8
    tri = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
    tri[:E, 0] = np.arange(E)  # every entity appears
    X = _labels(tri)
    S = (3 + eta) * n

    def emitted(m):
        f = m._fit
        pos = torch.from_numpy(tri).cuda()
        neg = f["neg"]
        if m._neg_batch:
            neg = dict(neg_entities=torch.unique(pos[:, [0, 2]]).to(torch.int32))
        a = f["eng"].train_args(ent=f["ent"], rel=f["rel"], pos=pos, loss_out=f["loss_dev"], step=1, **f["kw"], **f["st"], **neg)
        keys = torch.empty(S, dtype=torch.int32, device="cuda")
        f["eng"].train_emit(a, keys)
        torch.cuda.synchronize()
        return keys.cpu().numpy()[2 * n:2 * n + eta * n]

    base = dict(k=k, eta=eta, epochs=1, batches_count=1, seed=3, optimizer="sgd", optimizer_params={"lr": 1e-3}, loss="nll")
    m = DistMult(**base, embedding_model_params={"negative_corruption_entities": 37})
    m.fit(X)
    r = emitted(m)
    assert r.min() >= 0 and r.max() < 37 and len(np.unique(r)) == 37
    chosen = ["e%05d" % v for v in (5, 17, 333, 12)] + ["unknown-label"]
    m = DistMult(**base, embedding_model_params={"negative_corruption_entities": chosen})
    m.fit(X)
    assert set(np.unique(emitted(m)).tolist()) == {5, 12, 17, 333}
    sub = tri[:64].copy()
    m = DistMult(**base, embedding_model_params={"negative_corruption_entities": "batch"})
    m.fit(X)
    f = m._fit
    pos = torch.from_numpy(sub).cuda()
    a = f["eng"].train_args(ent=f["ent"], rel=f["rel"], pos=pos, loss_out=f["loss_dev"], step=1, **f["kw"], **f["st"],
                            neg_entities=torch.unique(pos[:, [0, 2]]).to(torch.int32))
    keys = torch.empty((3 + eta) * 64, dtype=torch.int32, device="cuda")
    f["eng"].train_emit(a, keys)
    torch.cuda.synchronize()
    r = keys.cpu().numpy()[2 * 64:2 * 64 + eta * 64]
    assert set(r.tolist()) <= set(sub[:, [0, 2]].reshape(-1).tolist())
    assert np.isfinite(m.loss_history[-1])
    for bad in (0, E + 1, "some", 1.5):
        with pytest.raises(ValueError):
            DistMult(**base, embedding_model_params={"negative_corruption_entities": bad}).fit(X)
    del _lib


def test_fit_with_lp_regulariser_and_new_losses_converges(engine):
    from emgraph_b200.models import ComplEx, TransE
    tri = ko.synthetic_triples(150, 4, 1500, seed=4)
    X = _labels(tri)
    for cls, loss, lp in ((ComplEx, "self_adversarial", {"margin": 3.0, "alpha": 0.5}), (TransE, "absolute_margin", {"margin": 2.0}),
                          (ComplEx, "nll", {})):
        m = cls(k=16, eta=5, epochs=12, batches_count=4, seed=0, optimizer="adam", optimizer_params={"lr": 0.02}, loss=loss,
                loss_params=lp, regularizer="LP", regularizer_params={"p": 2, "lambda": 1e-4})
        m.fit(X)
        assert np.all(np.isfinite(m.loss_history)) and m.loss_history[-1] < m.loss_history[0], (cls.__name__, loss, m.loss_history)
    # a strong penalty shrinks every row (touched rows in the reduction, the others in the dense pass)
    ent0 = np.random.default_rng(0).uniform(-0.5, 0.5, size=(150, 32)).astype(np.float32)
    rel0 = np.random.default_rng(1).uniform(-0.5, 0.5, size=(4, 32)).astype(np.float32)
    m = ComplEx(k=16, eta=2, epochs=3, batches_count=2, seed=0, optimizer="sgd", optimizer_params={"lr": 0.05}, loss="nll",
                regularizer="LP", regularizer_params={"p": 2, "lambda": 0.5}, initializer="constant",
                initializer_params={"entity": ent0, "relation": rel0})
    m.fit(X)
    assert np.all(np.linalg.norm(m.trained_model_params[0], axis=1) < np.linalg.norm(ent0, axis=1))


def test_save_restore_predict_and_resume(engine, tmp_path):
    from emgraph_b200 import restore_model, save_model
    from emgraph_b200.evaluation import evaluate_performance
    from emgraph_b200.models import HolE
    tri = ko.synthetic_triples(90, 3, 700, seed=6)
    X = _labels(tri)
    m = HolE(k=8, eta=3, epochs=4, batches_count=2, seed=2, optimizer="adam", optimizer_params={"lr": 0.01}, loss="multiclass_nll")
    m.fit(X)
    y0 = m.predict(X[:50])
    path = str(tmp_path / "hole.pkl")
    save_model(m, path, save_optimizer_state=True)
    r = restore_model(path)
    np.testing.assert_array_equal(r.predict(X[:50]), y0)
    np.testing.assert_array_equal(evaluate_performance(X[:30], r, filter_triples=X), evaluate_performance(X[:30], m, filter_triples=X))
    np.testing.assert_array_equal(r.predict(tri[:50], from_idx=True), y0)  # reference tests/.../test_models.py:967-992
    assert set(r._opt_state) == {"ent_m", "ent_v", "rel_m", "rel_v"} and r._opt_step == m._opt_step


@pytest.mark.parametrize("nl", ["tanh", "sigmoid", "softplus"])
def test_non_linearity_train_and_rank_match_oracle(engine, nl):
    """embedding_model_params['non_linearity'] (models/EmbeddingModel.py:679-689, :801-812, :1868-1881): the loss and
    the rank comparison see nl(score), and predict returns nl(score) as the reference does (:2135-2147)."""
    from emgraph_b200 import _lib
    from emgraph_b200.engine import model_id
    from emgraph_b200.evaluation import evaluate_performance
    from emgraph_b200.models import ComplEx
    rng = np.random.default_rng(12)
    E, R, k, eta, n, model = 160, 4, 20, 6, 200, "ComplEx"
    K = ko.internal_k(model, k)
    ent = (rng.normal(size=(E, K)) * 0.35).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.35).astype(np.float32)
    pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
    keep = rng.integers(0, 2, n * eta).astype(np.uint8)
    repl = rng.integers(0, E, n * eta).astype(np.int32)
    for loss in ("nll", "pairwise", "multiclass_nll", "self_adversarial", "absolute_margin"):
        ent_d, rel_d = torch.from_numpy(ent).cuda(), torch.from_numpy(rel).cuda()
        out = dict(loss=torch.zeros(1, device="cuda"), scores=torch.zeros(n * (1 + eta), device="cuda"),
                   g_ent=torch.zeros_like(ent_d), g_rel=torch.zeros_like(rel_d))
        a = engine.train_args(model=model_id(model), loss=_lib.LOSS_IDS[loss], opt=0, k=k, eta=eta, ent=ent_d, rel=rel_d,
                              pos=torch.from_numpy(pos).cuda(), loss_out=out["loss"], flags=_lib.F_NO_UPDATE, margin=0.7, alpha=0.8,
                              repl=torch.from_numpy(repl).cuda(), keep_subj=torch.from_numpy(keep).cuda(),
                              dbg_scores=out["scores"], dbg_grad_ent=out["g_ent"], dbg_grad_rel=out["g_rel"],
                              non_linearity=_lib.NL_IDS[nl])
        engine.train_step(a)
        torch.cuda.synchronize()
        o = ko.train_step(model, k, loss, eta, ent, rel, pos, keep, repl, margin=0.7, alpha=0.8, nl=nl, dtype=np.float64)
        np.testing.assert_allclose(out["scores"].cpu().numpy()[:n], o["scores_pos"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(out["scores"].cpu().numpy()[n:], o["scores_neg"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(out["loss"].item(), o["loss"], rtol=1e-5)
        sc = max(1.0, np.abs(o["grad_ent"]).max())
        np.testing.assert_allclose(out["g_ent"].cpu().numpy(), o["grad_ent"], rtol=1e-4, atol=1e-5 * sc)
        np.testing.assert_allclose(out["g_rel"].cpu().numpy(), o["grad_rel"], rtol=1e-4, atol=1e-5 * max(1.0, np.abs(o["grad_rel"]).max()))
    # ranking through the model API
    tri = ko.synthetic_triples(E, R, 1000, seed=9)
    X = _labels(tri)
    m = _fitted(ComplEx, k, E, R, ent, rel, X, embedding_model_params={"non_linearity": nl})
    test = tri[:40]
    for tc in (False, True):
        m.engine_params["rank_tensor_cores"] = tc
        got = evaluate_performance(_labels(test), m, filter_triples=X, corrupt_side="s,o")
        exp = ko.ranks(model, k, ent, rel, test, tri, "s,o", "worst", nl=nl)
        assert got.shape == exp.shape and (got != exp).sum() <= max(2, exp.size // 20), (nl, tc, (got != exp).sum())
    exp_pred = ko.non_linearity(nl, ko.score(model, k, ent, rel, tri[:20]))[0]
    np.testing.assert_allclose(m.predict(X[:20]), exp_pred, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(m.predict(tri[:20], from_idx=True), exp_pred, rtol=1e-5, atol=1e-6)
    for bad in ([[E, 0, 1]], [[0, R, 1]], [[0, 0, -1]]):  # the reference's gather raises on ids outside the tables
        with pytest.raises(ValueError):
            m.predict(np.asarray(bad), from_idx=True)
    with pytest.raises(ValueError):
        ComplEx(k=4, epochs=1, batches_count=1, embedding_model_params={"non_linearity": "relu"}).fit(X)
