"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and against the committed
outputs of the reference's own Python (tests/golden).  Tolerances: scores / losses / gradients
1e-5 relative (BASELINE.json north_star); ranks identical except where a candidate's score ties
the positive's within that tolerance (classified by ``_admissible``)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-5


def _ids(model, norm=1):
    from emgraph_b200.engine import model_id
    return model_id(model, norm)


def _dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _close(a, b, rtol=RTOL, scale=None):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    sc = float(np.abs(b).max()) if scale is None else scale  # tensor scale, no absolute floor
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * sc)


def _nl_id(g):
    from emgraph_b200 import _lib
    return _lib.NL_IDS[str(g["nl"])] if "nl" in g.files else 0


def _reg_kw(g):
    if "reg_p" not in g.files or int(g["reg_p"]) == 0:
        return {}
    return dict(reg_p=int(g["reg_p"]), reg_lambda_ent=float(g["reg_lambda_ent"]), reg_lambda_rel=float(g["reg_lambda_rel"]))


def run_step(engine, model, k, loss, eta, ent, rel, pos, keep, repl, margin=1.0, norm=1, opt="adam", lr=1e-3,
             flags=0, state=None, step=1, alpha=0.5, **reg):
    from emgraph_b200 import _lib
    n = pos.shape[0]
    ent_d, rel_d = _dev(ent), _dev(rel)
    out = dict(loss=torch.zeros(1, device="cuda"), scores=torch.zeros(n * (1 + eta), device="cuda"),
               g_ent=torch.zeros_like(ent_d), g_rel=torch.zeros_like(rel_d))
    st = {}
    if state is not None:
        st = {k_: _dev(v) for k_, v in state.items()}
    a = engine.train_args(model=_ids(model, norm), loss=_lib.LOSS_IDS[loss], opt=_lib.OPT_IDS[opt], k=k, eta=eta,
                          ent=ent_d, rel=rel_d, pos=_dev(pos, torch.int32), loss_out=out["loss"], flags=flags,
                          margin=margin, alpha=alpha, lr=lr, step=step, repl=_dev(repl, torch.int32), keep_subj=_dev(keep, torch.uint8),
                          dbg_scores=out["scores"], dbg_grad_ent=out["g_ent"], dbg_grad_rel=out["g_rel"], **st, **reg)
    engine.train_step(a)
    torch.cuda.synchronize()
    res = {k_: v.cpu().numpy() for k_, v in out.items()}
    res["ent"], res["rel"] = ent_d.cpu().numpy(), rel_d.cpu().numpy()
    res["state"] = {k_: v.cpu().numpy() for k_, v in st.items()}
    return res


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "train_*.npz"))), ids=os.path.basename)
def test_train_step_vs_reference_golden(engine, path):
    """scores, loss and summed row gradients against the reference's own _fn / loss.apply / autograd."""
    from emgraph_b200 import _lib
    g = np.load(path)
    model, k, eta = str(g["model"]), int(g["k"]), int(g["eta"])
    r = run_step(engine, model, k, str(g["loss_name"]), eta, g["ent"], g["rel"], g["pos"], g["keep_subj"], g["repl"],
                 margin=float(g["margin"]), norm=int(g["norm"]), flags=_lib.F_NO_UPDATE,
                 alpha=float(g["alpha"]) if "alpha" in g.files else 0.5, non_linearity=_nl_id(g), **_reg_kw(g))
    n = g["pos"].shape[0]
    _close(r["scores"][:n], g["scores_pos"])
    _close(r["scores"][n:], g["scores_neg"])
    np.testing.assert_allclose(r["loss"][0], g["loss"], rtol=RTOL)
    _close(r["g_ent"], g["grad_ent"])
    _close(r["g_rel"], g["grad_rel"])
    # NO_UPDATE leaves the parameters untouched
    np.testing.assert_array_equal(r["ent"], g["ent"])
    np.testing.assert_array_equal(r["rel"], g["rel"])
    # predict path agrees too (nl(score), models/EmbeddingModel.py:2135-2147; the goldens hold nl(score) as well)
    sc = engine.score(_ids(model, int(g["norm"])), k, _dev(g["ent"]), _dev(g["rel"]), _dev(g["pos"], torch.int32),
                      non_linearity=_nl_id(g)).cpu().numpy()
    _close(sc, g["scores_pos"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "wide_*.npz"))), ids=os.path.basename)
def test_wide_train_step_vs_reference_golden(engine, path):
    """K >= 256 rows on a 5000-entity table against the reference's own code (oracle/make_golden.py WIDE_CASES): the
    multi-chunk / bulk-copy kernel variants are pinned to the reference, not only to the NumPy oracle."""
    from emgraph_b200 import _lib
    from oracle.make_golden import wide_tables
    g = np.load(path)
    model, k, eta, E, R = str(g["model"]), int(g["k"]), int(g["eta"]), int(g["E"]), int(g["R"])
    ent, rel = wide_tables(int(g["table_seed"]), E, R, ko.internal_k(model, k))
    r = run_step(engine, model, k, str(g["loss_name"]), eta, ent, rel, g["pos"], g["keep_subj"], g["repl"], margin=float(g["margin"]),
                 norm=int(g["norm"]), flags=_lib.F_NO_UPDATE, alpha=float(g["alpha"]))
    n = g["pos"].shape[0]
    _close(r["scores"][:n], g["scores_pos"])
    _close(r["scores"][n:], g["scores_neg"])
    np.testing.assert_allclose(r["loss"][0], g["loss"], rtol=RTOL)
    rows = g["grad_rows"]
    rest = np.ones(E, bool)
    rest[rows] = False
    assert not r["g_ent"][rest].any()
    rt = 1e-4 if (model == "TransE" and int(g["norm"]) == 1) else RTOL
    _close(r["g_ent"][rows], g["grad_ent_rows"], rtol=rt)
    _close(r["g_rel"], g["grad_rel"], rtol=rt)


@pytest.mark.parametrize("opt", ["adam", "adagrad", "momentum", "sgd"])
def test_reset_state_update_matches_keras_first_step(engine, opt):
    """Reference-faithful mode: fresh optimizer state on every batch."""
    from emgraph_b200 import _lib
    g = np.load(os.path.join(GOLD, "train_complex_nll.npz"))
    model, k, eta = str(g["model"]), int(g["k"]), int(g["eta"])
    r = run_step(engine, model, k, "nll", eta, g["ent"], g["rel"], g["pos"], g["keep_subj"], g["repl"], opt=opt, lr=1e-2,
                 flags=_lib.F_RESET_STATE)
    o = ko.train_step(model, k, "nll", eta, g["ent"], g["rel"], g["pos"], g["keep_subj"], g["repl"], opt=opt, lr=1e-2)
    # rows without gradient are untouched, bit for bit
    np.testing.assert_array_equal(r["ent"][~o["touched_ent"]], g["ent"][~o["touched_ent"]])
    if opt == "adam":
        # fresh-state Adam ~ lr*sign(g): entries with |g| ~ eps are ill-conditioned; compare where |g| >> eps
        big = np.abs(o["grad_ent"]) > 1e-3
        np.testing.assert_allclose(r["ent"][big], o["ent_new"][big], rtol=1e-5, atol=1e-6)
    else:
        np.testing.assert_allclose(r["ent"], o["ent_new"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(r["rel"], o["rel_new"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("opt", ["adam", "adagrad", "momentum"])
def test_stateful_sparse_optimizer_three_steps(engine, opt):
    rng = np.random.default_rng(5)
    E, R, k, eta, n = 300, 7, 16, 4, 128
    model = "DistMult"
    ent = (rng.normal(size=(E, k)) * 0.5).astype(np.float32)
    rel = (rng.normal(size=(R, k)) * 0.5).astype(np.float32)
    if opt == "adam":
        st = dict(ent_m=np.zeros_like(ent), ent_v=np.zeros_like(ent), rel_m=np.zeros_like(rel), rel_v=np.zeros_like(rel))
    elif opt == "adagrad":
        st = dict(ent_m=np.full_like(ent, 0.1), rel_m=np.full_like(rel, 0.1))
    else:
        st = dict(ent_m=np.zeros_like(ent), rel_m=np.zeros_like(rel))
    o_state = None
    e_o, r_o = ent.copy(), rel.copy()
    for step in (1, 2, 3):
        pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
        keep = rng.integers(0, 2, n * eta).astype(np.uint8)
        repl = rng.integers(0, E, n * eta).astype(np.int32)
        r = run_step(engine, model, k, "pairwise", eta, ent, rel, pos, keep, repl, opt=opt, lr=5e-3, state=st, step=step)
        o = ko.train_step(model, k, "pairwise", eta, e_o, r_o, pos, keep, repl, opt=opt, lr=5e-3, state=o_state, step=step)
        if o_state is None and opt == "adagrad":
            pass
        ent, rel, st = r["ent"], r["rel"], r["state"]
        e_o, r_o, o_state = o["ent_new"], o["rel_new"], (o["state_ent"], o["state_rel"])
        np.testing.assert_allclose(ent, e_o, rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(rel, r_o, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("model,loss,k,eta", [
    ("TransE", "pairwise", 100, 20), ("DistMult", "pairwise", 200, 10), ("ComplEx", "nll", 200, 20),
    ("HolE", "multiclass_nll", 256, 20), ("DistMult", "nll", 256, 64), ("TransE", "multiclass_nll", 7, 33),
    ("ComplEx", "pairwise", 5, 3), ("DistMult", "multiclass_nll", 130, 40),
    ("ComplEx", "self_adversarial", 200, 20), ("DistMult", "absolute_margin", 200, 10), ("HolE", "self_adversarial", 256, 20),
    ("TransE", "self_adversarial", 100, 40), ("TransE", "absolute_margin", 9, 5), ("DistMult", "self_adversarial", 256, 64),
])
def test_train_step_vs_oracle_bench_shapes(engine, model, loss, k, eta):
    """The benchmark embedding sizes (and odd sizes that take the scalar path) against the oracle."""
    from emgraph_b200 import _lib
    rng = np.random.default_rng(11)
    E, R, n = 3000, 11, 257
    K = ko.internal_k(model, k)
    lim = 0.4 if model != "TransE" else 0.2
    ent = rng.uniform(-lim, lim, size=(E, K)).astype(np.float32)
    rel = rng.uniform(-lim, lim, size=(R, K)).astype(np.float32)
    pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
    keep = rng.integers(0, 2, n * eta).astype(np.uint8)
    repl = rng.integers(0, E, n * eta).astype(np.int32)
    margin = 5.0 if model == "DistMult" else 1.0
    r = run_step(engine, model, k, loss, eta, ent, rel, pos, keep, repl, margin=margin, flags=_lib.F_NO_UPDATE)
    o = ko.train_step(model, k, loss, eta, ent, rel, pos, keep, repl, margin=margin, dtype=np.float64)
    _close(r["scores"][:n], o["scores_pos"])
    _close(r["scores"][n:], o["scores_neg"])
    np.testing.assert_allclose(r["loss"][0], o["loss"], rtol=RTOL)
    if model == "TransE" and loss == "pairwise":
        # L1 sign gradients are exact integers unless a hinge sits exactly on the boundary
        _close(r["g_ent"], o["grad_ent"], rtol=1e-4)
    else:
        _close(r["g_ent"], o["grad_ent"])
        _close(r["g_rel"], o["grad_rel"])


@pytest.mark.parametrize("opt,p", [("sgd", 2), ("adam", 3), ("adagrad", 1), ("momentum", 2)])
def test_lp_regulariser_updates_every_row(engine, opt, p):
    """LP penalty (regularizers/lp.py:81-113): rows the batch touches get the penalty gradient added in the
    reduction, all other rows are updated by the dense pass; loss includes the penalty of the whole tables."""
    from emgraph_b200 import _lib
    rng = np.random.default_rng(17 + p)
    E, R, k, eta, n, model = 900, 6, 20, 5, 100, "ComplEx"
    K = ko.internal_k(model, k)
    ent = rng.uniform(-0.4, 0.4, size=(E, K)).astype(np.float32)
    rel = rng.uniform(-0.4, 0.4, size=(R, K)).astype(np.float32)
    pos = np.stack([rng.integers(0, E, n), rng.integers(0, R - 1, n), rng.integers(0, E, n)], 1).astype(np.int32)
    keep = rng.integers(0, 2, n * eta).astype(np.uint8)
    repl = rng.integers(0, E, n * eta).astype(np.int32)
    reg = dict(reg_p=p, reg_lambda_ent=3e-2, reg_lambda_rel=1e-1)
    r = run_step(engine, model, k, "nll", eta, ent, rel, pos, keep, repl, opt=opt, lr=1e-2, flags=_lib.F_RESET_STATE, **reg)
    o = ko.train_step(model, k, "nll", eta, ent, rel, pos, keep, repl, opt=opt, lr=1e-2, **reg)
    np.testing.assert_allclose(r["loss"][0], o["loss"], rtol=RTOL)
    _close(r["g_ent"], o["grad_ent"])
    _close(r["g_rel"], o["grad_rel"])
    if opt == "adam":
        big = np.abs(o["grad_ent"]) > 1e-4
        np.testing.assert_allclose(r["ent"][big], o["ent_new"][big], rtol=1e-5, atol=1e-6)
        assert (r["ent"] != ent).mean() > 0.99  # every row moved, touched or not
    else:
        np.testing.assert_allclose(r["ent"], o["ent_new"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(r["rel"], o["rel_new"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("E,n,eta", [(12, 300, 6), (40, 257, 20), (700, 64, 3)])
def test_hub_rows_long_and_short_runs(engine, E, n, eta):
    """Few entities, many slots: sorted runs of one row span many 16-slot chunks (hub path), two chunks
    (finished by the run-head warp) or one; the summed gradients and the update must match the oracle."""
    from emgraph_b200 import _lib
    rng = np.random.default_rng(E)
    R, k, model = 3, 24, "ComplEx"
    K = ko.internal_k(model, k)
    ent = rng.uniform(-0.4, 0.4, size=(E, K)).astype(np.float32)
    rel = rng.uniform(-0.4, 0.4, size=(R, K)).astype(np.float32)
    pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
    keep = rng.integers(0, 2, n * eta).astype(np.uint8)
    repl = rng.integers(0, E, n * eta).astype(np.int32)
    r = run_step(engine, model, k, "nll", eta, ent, rel, pos, keep, repl, opt="sgd", lr=1e-2, flags=_lib.F_RESET_STATE)
    o = ko.train_step(model, k, "nll", eta, ent, rel, pos, keep, repl, opt="sgd", lr=1e-2, dtype=np.float64)
    _close(r["g_ent"], o["grad_ent"])
    _close(r["g_rel"], o["grad_rel"])
    np.testing.assert_allclose(r["ent"], o["ent_new"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r["rel"], o["rel_new"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("model,k,n,eta,E", [("DistMult", 4, 1, 1, 3), ("ComplEx", 2, 3, 2, 5), ("TransE", 1, 5, 3, 4), ("HolE", 3, 2, 33, 40),
                                             ("DistMult", 1000, 7, 2, 50), ("ComplEx", 256, 9, 3, 60), ("TransE", 516, 4, 5, 30)])
def test_train_step_edge_shapes(engine, model, k, n, eta, E):
    """Tiny and ragged shapes: a single positive, eta = 1, k below one vector, rows that need more than one pass
    of a warp (K = 1000, 1024 floats), fewer entities than negatives (every row a duplicate run)."""
    from emgraph_b200 import _lib
    rng = np.random.default_rng(n * 31 + eta)
    R = 2
    K = ko.internal_k(model, k)
    ent = rng.uniform(-0.5, 0.5, size=(E, K)).astype(np.float32)
    rel = rng.uniform(-0.5, 0.5, size=(R, K)).astype(np.float32)
    pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
    keep = rng.integers(0, 2, n * eta).astype(np.uint8)
    repl = rng.integers(0, E, n * eta).astype(np.int32)
    for loss in ("nll", "multiclass_nll"):
        r = run_step(engine, model, k, loss, eta, ent, rel, pos, keep, repl, opt="sgd", lr=1e-2, flags=_lib.F_RESET_STATE)
        o = ko.train_step(model, k, loss, eta, ent, rel, pos, keep, repl, opt="sgd", lr=1e-2, dtype=np.float64)
        np.testing.assert_allclose(r["loss"][0], o["loss"], rtol=RTOL)
        _close(r["scores"][:n], o["scores_pos"])
        _close(r["scores"][n:], o["scores_neg"])
        if not (model == "TransE" and k == 1):  # |x| has no gradient at exact ties of 1-d L1 distances
            _close(r["g_ent"], o["grad_ent"], rtol=1e-4)
            _close(r["g_rel"], o["grad_rel"], rtol=1e-4)
            np.testing.assert_allclose(r["ent"], o["ent_new"], rtol=1e-5, atol=1e-5)
        np.testing.assert_array_equal(r["ent"][~o["touched_ent"]], ent[~o["touched_ent"]])


def test_train_step_rejects_bad_arguments(engine):
    """Error conventions of the C ABI: status < 0 + kge_last_error, surfaced as KgeError (never a CPU fallback)."""
    from emgraph_b200 import _lib
    ent, rel = torch.zeros((10, 8), device="cuda"), torch.zeros((2, 8), device="cuda")
    pos = torch.zeros((4, 3), dtype=torch.int32, device="cuda")
    loss = torch.zeros(1, device="cuda")
    base = dict(model=2, loss=1, opt=0, k=8, eta=2, ent=ent, rel=rel, pos=pos, loss_out=loss)
    for bad in (dict(model=9), dict(loss=7), dict(opt=5), dict(k=16), dict(eta=0), dict(non_linearity=4), dict(neg_entities_n=11)):
        with pytest.raises(_lib.KgeError):
            engine.train_step(engine.train_args(**{**base, **bad}))
    # an embedding too wide for the fused kernel is refused loudly
    wide_e, wide_r = torch.zeros((4, 8192), device="cuda"), torch.zeros((2, 8192), device="cuda")
    with pytest.raises(_lib.KgeError):
        engine.train_step(engine.train_args(**{**base, "k": 8192, "ent": wide_e, "rel": wide_r, "flags": _lib.F_NO_UPDATE}))
    # an empty batch is a no-op
    engine.train_step(engine.train_args(**{**base, "pos": torch.zeros((0, 3), dtype=torch.int32, device="cuda")}))
    torch.cuda.synchronize()


def test_host_step_graph_replay_matches_device_step(engine):
    """kge_train_step_host (eager first call, then a captured CUDA graph replayed with a fresh step
    counter) must leave the same parameters and losses as the device-resident step, bit for bit."""
    from emgraph_b200 import models
    rng = np.random.default_rng(3)
    E, R, k, eta, n = 900, 9, 32, 8, 512
    X = np.stack([rng.integers(0, E, 4 * n), rng.integers(0, R, 4 * n), rng.integers(0, E, 4 * n)], 1).astype(np.int32)
    ent0 = rng.uniform(-0.3, 0.3, size=(E, 2 * k)).astype(np.float32)
    rel0 = rng.uniform(-0.3, 0.3, size=(R, 2 * k)).astype(np.float32)
    out = {}
    for mode in ("device", "host", "pipelined"):
        m = models.ComplEx(k=k, eta=eta, epochs=1, batches_count=4, seed=5, optimizer="adam", optimizer_params={"lr": 1e-2},
                           loss="nll", initializer="constant", initializer_params={"entity": ent0, "relation": rel0})
        f = m._fit_prepare(E, R)
        losses = []
        Xd = torch.from_numpy(X).cuda()
        Xh = torch.from_numpy(X).pin_memory()
        for step in range(9):
            lo, hi = (step % 4) * n, (step % 4 + 1) * n
            if mode == "device":
                m._fit_step_device(Xd[lo:hi])
                torch.cuda.synchronize()
                losses.append(float(f["loss_dev"].item()))
            elif mode == "host":
                losses.append(m._fit_step_host(Xh[lo:hi]))
            else:  # the loss of step t comes back from the call that submits step t+1, the last one from the flush
                lv = m._fit_step_host_pipelined(Xh[lo:hi])
                assert (lv is None) == (step == 0)
                if lv is not None:
                    losses.append(lv)
        if mode == "pipelined":
            losses.append(m._fit_host_flush())
            assert m._fit_host_flush() is None
        torch.cuda.synchronize()
        out[mode] = (f["ent"].cpu().numpy(), f["rel"].cpu().numpy(), np.asarray(losses))
    np.testing.assert_array_equal(out["device"][0], out["host"][0])
    np.testing.assert_array_equal(out["device"][1], out["host"][1])
    np.testing.assert_array_equal(out["device"][2], out["host"][2])
    for j in range(3):
        np.testing.assert_array_equal(out["host"][j], out["pipelined"][j])
    assert np.all(np.isfinite(out["host"][2])) and out["host"][2][-1] < out["host"][2][0]


def test_in_kernel_corruptions_are_uniform_and_reproducible(engine):
    from emgraph_b200 import _lib
    E, R, k, eta, n = 1000, 5, 8, 16, 4096
    rng = np.random.default_rng(0)
    ent = _dev(rng.normal(size=(E, k)).astype(np.float32))
    rel = _dev(rng.normal(size=(R, k)).astype(np.float32))
    pos = _dev(np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1), torch.int32)
    S = (3 + eta) * n
    outs = []
    for seed, step in ((7, 1), (7, 1), (7, 2), (8, 1)):
        keys = torch.empty(S, dtype=torch.int32, device="cuda")
        a = engine.train_args(model=2, loss=0, opt=0, k=k, eta=eta, ent=ent, rel=rel, pos=pos,
                              loss_out=torch.zeros(1, device="cuda"), seed=seed, step=step)
        engine.train_emit(a, keys)
        torch.cuda.synchronize()
        outs.append(keys.cpu().numpy())
    np.testing.assert_array_equal(outs[0], outs[1])
    assert (outs[0] != outs[2]).mean() > 0.5 and (outs[0] != outs[3]).mean() > 0.5
    p = pos.cpu().numpy()
    np.testing.assert_array_equal(outs[0][:n], p[:, 0])
    np.testing.assert_array_equal(outs[0][n:2 * n], p[:, 2])
    np.testing.assert_array_equal(outs[0][2 * n + eta * n:], E + p[:, 1])
    repl = outs[0][2 * n:2 * n + eta * n]
    assert repl.min() >= 0 and repl.max() < E
    hist = np.bincount(repl, minlength=E)
    exp = eta * n / E
    assert abs(hist.mean() - exp) < 1e-9 and hist.std() < 2.0 * np.sqrt(exp)


# ------------------------------------------------------------------------------------------------
# ranking
# ------------------------------------------------------------------------------------------------
def _admissible(model, k, ent, rel, x, side_col, got, exp, norm, nl="linear"):
    """A rank may differ from the oracle's only by the number of candidates whose score sits within
    fp32 noise of the positive's x1e5 quantisation boundary."""
    so, ss, sp = ko.sweep_scores(model, k, ent, rel, x, norm)
    so, ss, sp = (ko.non_linearity(nl, np.asarray(v, np.float64), np.float64)[0] for v in (so, ss, sp))
    sc = so if side_col == 1 else ss
    tol = 1.0 + 1e-6 * 1e5 * max(1.0, np.abs(sc).max())  # quanta
    near = np.sum(np.abs(sc * 1e5 - sp * 1e5) <= tol)
    return abs(int(got) - int(exp)) <= near


def _rank_gpu(engine, model, k, ent, rel, test, filt, side, strat, norm=1, tc=False, nl=0):
    from emgraph_b200 import _lib
    ent_d, rel_d = _dev(ent), _dev(rel)
    if filt is not None:
        engine.filter_build(_dev(filt, torch.int32), ent.shape[0], rel.shape[0])
    r = engine.rank(_ids(model, norm), k, ent_d, rel_d, _dev(test, torch.int32), side=_lib.RANK_SIDE_IDS[side],
                    strategy=_lib.STRATEGY_IDS[strat], filtered=filt is not None, use_tensor_cores=tc, non_linearity=nl)
    torch.cuda.synchronize()
    return r.cpu().numpy()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "rank_*.npz"))), ids=os.path.basename)
def test_ranks_vs_reference_golden(engine, path):
    g = np.load(path)
    model, k, norm = str(g["model"]), int(g["k"]), int(g["norm"])
    for side in ("s,o", "s+o", "s", "o"):
        for strat in ("worst", "best", "middle"):
            for fl in (0, 1):
                key = "ranks_%s_%s_%d" % (side.replace(",", "c").replace("+", "p"), strat, fl)
                got = _rank_gpu(engine, model, k, g["ent"], g["rel"], g["test"], g["filt"] if fl else None, side, strat, norm,
                                nl=_nl_id(g))
                exp = g[key]
                assert got.shape == exp.shape, key
                bad = np.argwhere(got != exp)
                for idx in bad:
                    t = idx[0]
                    col = idx[1] if exp.ndim == 2 else (1 if side == "o" else 0)
                    assert _admissible(model, k, g["ent"], g["rel"], g["test"][t], col, got[tuple(idx)], exp[tuple(idx)], norm,
                                       str(g["nl"]) if "nl" in g.files else "linear"), (key, t, got[tuple(idx)], exp[tuple(idx)])
                assert len(bad) <= max(1, exp.size // 50), (key, len(bad))


@pytest.mark.parametrize("model,k,E,norm", [("TransE", 100, 4100, 1), ("TransE", 33, 700, 2), ("DistMult", 200, 1500, 1),
                                             ("ComplEx", 200, 1500, 1), ("HolE", 35, 900, 1)])
def test_ranks_vs_oracle_larger(engine, model, k, E, norm):
    rng = np.random.default_rng(3)
    R, F, T = 9, 6000, 150
    K = ko.internal_k(model, k)
    ent = (rng.normal(size=(E, K)) * 0.3).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.3).astype(np.float32)
    filt = ko.synthetic_triples(E, R, F, seed=9, zipf=True)
    test = filt[rng.permutation(F)[:T]]
    for side, fl in (("s,o", True), ("s,o", False), ("s+o", True)):
        got = _rank_gpu(engine, model, k, ent, rel, test, filt if fl else None, side, "worst", norm)
        exp = ko.ranks(model, k, ent, rel, test, filt if fl else None, side, "worst", norm)
        bad = np.argwhere(got != exp)
        for idx in bad:
            t = idx[0]
            col = idx[1] if exp.ndim == 2 else 1
            if exp.ndim == 2:
                assert _admissible(model, k, ent, rel, test[t], col, got[tuple(idx)], exp[tuple(idx)], norm)
        assert len(bad) <= max(2, exp.size // 25), (side, fl, len(bad))
    # size-independent properties (reference tests/emgraph/evaluation/test_protocol.py:174-301,
    # tests/emgraph/models/test_models.py:203-215)
    so_f = _rank_gpu(engine, model, k, ent, rel, test, filt, "s,o", "worst", norm)
    s_f = _rank_gpu(engine, model, k, ent, rel, test, filt, "s", "worst", norm)
    o_f = _rank_gpu(engine, model, k, ent, rel, test, filt, "o", "worst", norm)
    so_u = _rank_gpu(engine, model, k, ent, rel, test, None, "s,o", "worst", norm)
    np.testing.assert_array_equal(so_f[:, 0], s_f)
    np.testing.assert_array_equal(so_f[:, 1], o_f)
    assert np.all(so_f <= so_u) and so_f.min() >= 1 and so_u.min() >= 2 and so_u.max() <= E + 1
    best = _rank_gpu(engine, model, k, ent, rel, test, filt, "s,o", "best", norm)
    mid = _rank_gpu(engine, model, k, ent, rel, test, filt, "s,o", "middle", norm)
    assert np.all(best <= mid) and np.all(mid <= so_f)


def test_rank_edge_cases(engine):
    rng = np.random.default_rng(2)
    E, R, k = 70, 2, 8
    ent = rng.normal(size=(E, k)).astype(np.float32)
    rel = rng.normal(size=(R, k)).astype(np.float32)
    # empty test set
    got = _rank_gpu(engine, "DistMult", k, ent, rel, np.zeros((0, 3), np.int32), None, "s,o", "worst")
    assert got.shape == (0, 2)
    # empty filter: only self is filtered
    test = np.array([[1, 0, 2], [3, 1, 3]], np.int32)
    got = _rank_gpu(engine, "DistMult", k, ent, rel, test, np.zeros((0, 3), np.int32), "s,o", "worst")
    exp = ko.ranks("DistMult", k, ent, rel, test, np.zeros((0, 3), np.int32), "s,o", "worst")
    np.testing.assert_array_equal(got, exp)
    # every candidate known: rank 1 on the object side
    filt = np.array([[1, 0, e] for e in range(E)], np.int32)
    got = _rank_gpu(engine, "DistMult", k, ent, rel, test[:1], filt, "o", "worst")
    assert got.tolist() == [1]
    # all-equal embeddings: every score ties (worst / best / middle differ maximally)
    ent1 = np.ones((E, k), np.float32)
    rel1 = np.ones((R, k), np.float32)
    for strat in ("worst", "best", "middle"):
        got = _rank_gpu(engine, "TransE", k, ent1, rel1, test, None, "s,o", strat)
        exp = ko.ranks("TransE", k, ent1, rel1, test, None, "s,o", strat)
        np.testing.assert_array_equal(got, exp)


# ------------------------------------------------------------------------------------------------
# tensor-core (tcgen05 3xTF32) ranking sweep
# ------------------------------------------------------------------------------------------------
def _assert_ranks_close(model, k, ent, rel, test, got, exp, norm=1, max_frac=0.04, nl="linear"):
    assert got.shape == exp.shape
    bad = np.argwhere(got != exp)
    for idx in bad:
        t = idx[0]
        col = idx[1] if exp.ndim == 2 else 1
        if exp.ndim == 2:
            assert _admissible(model, k, ent, rel, test[t], col, got[tuple(idx)], exp[tuple(idx)], norm, nl), \
                (t, col, got[tuple(idx)], exp[tuple(idx)])
    assert len(bad) <= max(2, int(exp.size * max_frac)), len(bad)


@pytest.mark.parametrize("path", sorted(p for p in glob.glob(os.path.join(GOLD, "rank_*.npz")) if "transe" not in p),
                         ids=os.path.basename)
def test_tc_ranks_vs_reference_golden(engine, path):
    if not engine.has_tensor_core_rank():
        pytest.skip("library built without the tcgen05 sweep")
    g = np.load(path)
    model, k, norm = str(g["model"]), int(g["k"]), int(g["norm"])
    for side in ("s,o", "s+o", "s", "o"):
        for strat in ("worst", "middle"):
            for fl in (0, 1):
                key = "ranks_%s_%s_%d" % (side.replace(",", "c").replace("+", "p"), strat, fl)
                got = _rank_gpu(engine, model, k, g["ent"], g["rel"], g["test"], g["filt"] if fl else None, side, strat, norm, tc=True,
                                nl=_nl_id(g))
                exp = g[key]
                if exp.ndim == 2:
                    _assert_ranks_close(model, k, g["ent"], g["rel"], g["test"], got, exp, norm,
                                        nl=str(g["nl"]) if "nl" in g.files else "linear")
                else:
                    assert got.shape == exp.shape and (got != exp).sum() <= max(1, exp.size // 25), key


@pytest.mark.parametrize("model,k,E,T", [("DistMult", 200, 1500, 150), ("ComplEx", 200, 1500, 150), ("HolE", 35, 900, 70),
                                          ("DistMult", 64, 5000, 700), ("ComplEx", 100, 14541, 300), ("ComplEx", 200, 14541, 200)])
def test_tc_ranks_vs_fp32_sweep_and_oracle(engine, model, k, E, T):
    """The tcgen05 path against the fp32 CUDA-core sweep (same quantisation rule) and the oracle."""
    if not engine.has_tensor_core_rank():
        pytest.skip("library built without the tcgen05 sweep")
    rng = np.random.default_rng(7)
    R, F = 9, 8000
    K = ko.internal_k(model, k)
    ent = (rng.normal(size=(E, K)) * 0.3).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.3).astype(np.float32)
    filt = ko.synthetic_triples(E, R, F, seed=11, zipf=True)
    test = filt[rng.permutation(F)[:T]]
    for side, fl in (("s,o", True), ("s,o", False), ("s+o", True), ("o", True), ("s", True)):
        tc = _rank_gpu(engine, model, k, ent, rel, test, filt if fl else None, side, "worst", tc=True)
        ref = _rank_gpu(engine, model, k, ent, rel, test, filt if fl else None, side, "worst", tc=False)
        assert tc.shape == ref.shape
        if side == "s,o":
            _assert_ranks_close(model, k, ent, rel, test, tc, ref)
        else:
            assert (tc != ref).mean() <= 0.05
        assert np.abs(tc.astype(np.int64) - ref).max() <= 3
    # ... and against the ORACLE itself: all test triples on the small tables, a sample of them at E = 14 541 (the
    # FB15k-237 size the benchmark ranks), so that the tensor-core path is pinned to the oracle at full width as well
    n_chk = T if E <= 2000 else 48
    sub = test[:n_chk]
    exp = ko.ranks(model, k, ent, rel, sub, filt, "s,o", "worst")
    got = _rank_gpu(engine, model, k, ent, rel, sub, filt, "s,o", "worst", tc=True)
    _assert_ranks_close(model, k, ent, rel, sub, got, exp)
    # size-independent properties on the tensor-core path
    so_f = _rank_gpu(engine, model, k, ent, rel, test, filt, "s,o", "worst", tc=True)
    so_u = _rank_gpu(engine, model, k, ent, rel, test, None, "s,o", "worst", tc=True)
    assert np.all(so_f <= so_u) and so_f.min() >= 1 and so_u.min() >= 2 and so_u.max() <= E + 1


# ------------------------------------------------------------------ the sort in front of the segmented reduction
def _entries(rng, n, n_keys, zipf):
    if zipf:
        w = 1.0 / np.arange(1, n_keys + 1) ** 1.1
        keys = rng.choice(n_keys, size=n, p=w / w.sum())
    else:
        keys = rng.integers(0, n_keys, n)
    slots = np.arange(n, dtype=np.int64) | (rng.integers(0, 4, n).astype(np.int64) << 30)  # slot id + the two side bits
    return (keys.astype(np.int64) << 32) | slots


@pytest.mark.parametrize("n,n_keys,zipf", [(1, 1, False), (31, 7, False), (512, 100, True), (513, 100, True), (97796, 14778, True),
                                           (97796, 14778, False), (55276, 14778, True), (262144, 2000, True), (40000, 32768, False),
                                           (5000, 1, False), (2048, 65536, False), (2049, 300, True), (200000, 65536, True)])
def test_counting_sort_equals_stable_radix_sort(engine, n, n_keys, zipf):
    """kge_sort_small.cu (single-launch two-pass radix sort of small batches over small key ranges) against cub's stable
    radix sort and against a stable host sort: equal keys keep their input order, bit for bit.  Run twice: its grid barrier
    must be left ready for the next launch."""
    rng = np.random.default_rng(n + n_keys)
    for rep in range(2):
        e = _entries(rng, n, n_keys, zipf)
        d = torch.from_numpy(e).cuda()
        want = e[np.argsort(e >> 32, kind="stable")]
        got_radix = engine.sort_entries(d, n_keys, algo=0).cpu().numpy()
        got_count = engine.sort_entries(d, n_keys, algo=1).cpu().numpy()
        np.testing.assert_array_equal(got_radix, want)
        np.testing.assert_array_equal(got_count, want)


def test_counting_sort_sizes_in_sequence_and_out_of_range(engine):
    """different shapes through the same context (different grids on the same barrier words), then a size outside the
    small sort's range: refused loudly by the hook, routed to the radix sort by the train step."""
    rng = np.random.default_rng(5)
    for n, n_keys in [(3000, 900), (70000, 14000), (100, 32768), (262144, 1500), (3000, 900)]:
        e = _entries(rng, n, n_keys, True)
        got = engine.sort_entries(torch.from_numpy(e).cuda(), n_keys, algo=1).cpu().numpy()
        np.testing.assert_array_equal(got, e[np.argsort(e >> 32, kind="stable")])
    e = _entries(rng, 1000, 70000, False)
    from emgraph_b200._lib import KgeError
    with pytest.raises(KgeError):
        engine.sort_entries(torch.from_numpy(e).cuda(), 70000, algo=1)
    with pytest.raises(KgeError):
        engine.sort_entries(torch.from_numpy(_entries(rng, 300000, 900, True)).cuda(), 900, algo=1)
    got = engine.sort_entries(torch.from_numpy(e).cuda(), 70000, algo=0).cpu().numpy()
    np.testing.assert_array_equal(got, e[np.argsort(e >> 32, kind="stable")])
