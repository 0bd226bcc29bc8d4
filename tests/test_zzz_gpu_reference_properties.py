"""GPU: the reference's own end-to-end model tests, run through emgraph_b200 with the reference's constructor
arguments (tests/emgraph/models/test_models.py:218-335 fit/predict ordinal property on the toy graph, :338-367 refit
determinism), and fit() as a whole against the oracle's restatement of the loop (oracle/kge_oracle.py:fit_emulation)."""
import numpy as np
import pytest

from toy_graph import TOY_CASES, TOY_QUERY, TOY_X, toy_fit_emulation

pytestmark = pytest.mark.gpu


def _toy_model(model, bc, margin, reg, **kw):
    from emgraph_b200 import models
    extra = dict(regularizer="LP", regularizer_params={"lambda": reg[0], "p": reg[1]}) if reg else {}
    return getattr(models, model)(batches_count=bc, seed=555, epochs=20, k=10, loss="pairwise", loss_params={"margin": margin},
                                  optimizer="adagrad", optimizer_params={"lr": 0.1}, **extra, **kw)


@pytest.mark.parametrize("model,bc,margin,reg", TOY_CASES)
def test_fit_predict_toy_graph(engine, model, bc, margin, reg):
    m = _toy_model(model, bc, margin, reg)
    m.fit(TOY_X)
    y = m.predict(TOY_QUERY)
    assert y[0] > y[1]  # the reference's assertion
    # the whole fit loop against the oracle's restatement of it (fp32 on both sides; 20 to 40 optimizer steps)
    y_ref, losses_ref = toy_fit_emulation(model, bc, margin, reg)
    np.testing.assert_allclose(y, y_ref, rtol=2e-2, atol=2e-2)
    denom = 8 * 2  # batch_size * batches_count with eta-tiled positives (models/EmbeddingModel.py:1343-1344, :1453-1457)
    np.testing.assert_allclose(m.loss_history[-1], losses_ref[-1] / denom, rtol=2e-2)


def test_refit_is_deterministic(engine):
    """reference tests/emgraph/models/test_models.py:338-367: fitting the same model twice gives the same predictions."""
    m = _toy_model("ComplEx", 1, 1.0, (0.1, 2))
    m.fit(TOY_X)
    y1 = m.predict(TOY_QUERY)
    m.fit(TOY_X)
    y2 = m.predict(TOY_QUERY)
    np.testing.assert_array_equal(y1, y2)
    # host batches take the graph-replay path: same numbers
    mh = _toy_model("ComplEx", 1, 1.0, (0.1, 2), engine_params={"host_batches": True})
    mh.fit(TOY_X)
    np.testing.assert_array_equal(mh.predict(TOY_QUERY), y1)


def test_resume_equals_training_in_one_go(engine, tmp_path):
    """engine_params['resume']: 2 epochs + save (with the sparse optimizer's state) + restore + 2 epochs == 4 epochs,
    bit for bit -- the kernels are deterministic and parameters, per-row Adam state and the global step carry over
    (host logic pinned on CPU by tests/test_fit_host_logic.py::test_resume_continues_from_saved_optimizer_state)."""
    from emgraph_b200 import models, restore_model, save_model
    from oracle import kge_oracle as ko
    tri = ko.synthetic_triples(200, 4, 3000, seed=12, zipf=True)
    X = np.empty(tri.shape, dtype=object)
    X[:, 0] = ["e%04d" % v for v in tri[:, 0]]
    X[:, 1] = ["r%d" % v for v in tri[:, 1]]
    X[:, 2] = ["e%04d" % v for v in tri[:, 2]]
    X = X.astype(str)
    kw = dict(k=16, eta=4, batches_count=5, seed=4, optimizer="adam", optimizer_params={"lr": 0.01}, loss="nll")
    straight = models.ComplEx(epochs=4, **kw)
    straight.fit(X)
    first = models.ComplEx(epochs=2, **kw)
    first.fit(X)
    second = restore_model(save_model(first, str(tmp_path / "m.pkl"), save_optimizer_state=True))
    second.engine_params["resume"] = True
    second.fit(X)
    assert second._opt_step == straight._opt_step == 20
    np.testing.assert_array_equal(second.trained_model_params[0], straight.trained_model_params[0])
    np.testing.assert_array_equal(second.trained_model_params[1], straight.trained_model_params[1])
