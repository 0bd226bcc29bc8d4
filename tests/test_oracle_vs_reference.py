"""The oracle against the reference's OWN Python executed live (oracle/ref_shim.py: the reference's modules imported
unmodified from /root/reference over a torch-backed `tensorflow` shim).  Builder-container only: skipped wherever
/root/reference is absent (the GPU box, the driver's CPU tier outside the builder image).  The committed goldens
under tests/golden/ are outputs of the same shim, so this test is what shows they still regenerate."""
import os

import numpy as np
import pytest

from oracle import kge_oracle as ko
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference is not present")

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("model,loss,k,norm,nl", [("TransE", "pairwise", 10, 1, "linear"), ("TransE", "nll", 8, 2, "linear"),
                                                  ("DistMult", "multiclass_nll", 12, 1, "tanh"), ("ComplEx", "nll", 6, 1, "linear"),
                                                  ("HolE", "self_adversarial", 8, 1, "sigmoid"), ("ComplEx", "absolute_margin", 5, 1, "linear")])
def test_oracle_train_step_equals_reference_code(model, loss, k, norm, nl):
    rng = np.random.default_rng(hash((model, loss)) % 1000)
    E, R, eta, n = 70, 4, 5, 37
    K = ko.internal_k(model, k)
    ent = (rng.normal(size=(E, K)) * 0.5).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.5).astype(np.float32)
    pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
    keep = rng.integers(0, 2, n * eta).astype(np.uint8)
    repl = rng.integers(0, E, n * eta).astype(np.int32)
    ep = {"norm": norm, "non_linearity": nl}
    ref = ref_shim.ref_train_forward_backward(model, k, eta, loss, ent, rel, pos, keep, repl, {}, ep, "s,o", None, None)
    margin = 3.0 if loss == "self_adversarial" else 1.0
    o = ko.train_step(model, k, loss, eta, ent, rel, pos, keep, repl, margin=margin, norm=norm, nl=nl)
    np.testing.assert_array_equal(o["neg"], ref["neg"])
    np.testing.assert_allclose(o["scores_pos"], ref["scores_pos"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(o["scores_neg"], ref["scores_neg"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(o["loss"], ref["loss"], rtol=1e-5)
    np.testing.assert_allclose(o["grad_ent"], ref["grad_ent"], rtol=1e-4, atol=1e-5 * np.abs(ref["grad_ent"]).max())
    np.testing.assert_allclose(o["grad_rel"], ref["grad_rel"], rtol=1e-4, atol=1e-5 * np.abs(ref["grad_rel"]).max())


@pytest.mark.parametrize("model,k,norm", [("TransE", 8, 1), ("DistMult", 10, 1), ("ComplEx", 6, 1), ("HolE", 8, 1)])
def test_oracle_ranks_equal_reference_code(model, k, norm):
    rng = np.random.default_rng(k)
    E, R = 60, 3
    K = ko.internal_k(model, k)
    ent = (rng.normal(size=(E, K)) * 0.6).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.6).astype(np.float32)
    filt = ko.synthetic_triples(E, R, 400, seed=k)
    test = filt[:15]
    for side in ("s,o", "s+o", "o"):
        for strat in ("worst", "middle"):
            r = ref_shim.ref_ranks(model, k, ent, rel, test, filt, side, strat, {"norm": norm})
            np.testing.assert_array_equal(ko.ranks(model, k, ent, rel, test, filt, side, strat, norm), r)


def test_committed_golden_regenerates():
    """one committed fixture, regenerated live, equals the file bit for bit"""
    g = np.load(os.path.join(GOLD, "train_complex_nll.npz"))
    ref = ref_shim.ref_train_forward_backward(str(g["model"]), int(g["k"]), int(g["eta"]), str(g["loss_name"]), g["ent"], g["rel"], g["pos"],
                                              g["keep_subj"], g["repl"], {}, {}, str(g["side"]), None, None)
    np.testing.assert_array_equal(ref["scores_neg"], g["scores_neg"])
    np.testing.assert_array_equal(ref["grad_ent"], g["grad_ent"])
    assert np.float32(ref["loss"]) == g["loss"]
