"""Run the reference's OWN Python (bi-graph/Emgraph, /root/reference) over a `tensorflow` shim.

TEST INFRASTRUCTURE ONLY, builder-container only: /root/reference does not exist on the GPU box,
so nothing in `-m gpu` tests, smoke() or bench.py imports this.  Its single consumer is
``oracle/make_golden.py`` (golden-vector generation) and ``tests/test_oracle_vs_reference.py``
(skipped when /root/reference is absent).

TensorFlow 2.2 (requirements/default.txt:11) is not installable here, so a ~40-op `tensorflow`
module backed by torch-CPU fp32 ops is placed in ``sys.modules``; the reference's modules are then
imported unmodified from ``/root/reference``.  Random draws (`tf.random.uniform`) pop pre-drawn
arrays from a FIFO so that corruption indices are an INPUT while the reference's own code
assembles the negatives.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

REF_ROOT = os.environ.get("EMGRAPH_REF", "/root/reference")

RANDOM_FIFO: list = []


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "emgraph"))


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


class _Shape(list):
    """tf.shape(x): entries behave as python ints (the reference multiplies them)."""


def _build_tf():
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.float64 = torch.float32, torch.float64
    tf.int32, tf.int64, tf.bool = torch.int32, torch.int64, torch.bool
    tf.Tensor = torch.Tensor

    def constant(v, dtype=None, name=None, shape=None):
        return _t(v, dtype)

    def Variable(v, dtype=None, trainable=True, name=None, **kw):
        t = _t(v, dtype).clone()
        if trainable and t.is_floating_point():
            t.requires_grad_(True)
        return t

    def gather(p, ids, *a, **k):
        return p[_t(ids).long()]

    tf.constant, tf.Variable, tf.gather = constant, Variable, gather
    tf.convert_to_tensor = constant
    nn = types.ModuleType("tensorflow.nn")
    nn.embedding_lookup = lambda p, ids, **k: p[_t(ids).long()]
    nn.softmax = lambda x, axis=-1: torch.softmax(_t(x), dim=axis)
    tf.nn = nn
    tf.reduce_sum = lambda x, axis=None, **k: torch.sum(_t(x)) if axis is None else torch.sum(_t(x), dim=axis)
    tf.reduce_mean = lambda x, axis=None, **k: torch.mean(_t(x)) if axis is None else torch.mean(_t(x), dim=axis)
    tf.split = lambda x, n, axis=0: torch.chunk(x, n, dim=axis)
    tf.norm = lambda x, ord=2, axis=None, **k: torch.linalg.vector_norm(x, ord=ord, dim=axis)
    tf.negative = torch.neg
    tf.maximum = lambda a, b: torch.maximum(_t(a), _t(b, _t(a).dtype))
    tf.exp, tf.tanh, tf.sigmoid = torch.exp, torch.tanh, torch.sigmoid
    tf.concat = lambda xs, axis=0: torch.cat(list(xs), dim=axis)
    tf.stack = lambda xs, axis=0: torch.stack([_t(x) for x in xs], dim=axis)
    tf.transpose = lambda x, perm=None: x.permute(*perm) if perm is not None else x.t()
    tf.squeeze = lambda x, axis=None: x.squeeze() if axis is None else x.squeeze(axis)
    tf.expand_dims = lambda x, axis: _t(x).unsqueeze(axis)
    tf.logical_not = torch.logical_not
    tf.equal = lambda a, b: _t(a) == _t(b)
    tf.clip_by_value = lambda v, clip_value_min, clip_value_max: torch.clamp(v, clip_value_min, clip_value_max)
    tf.reshape = lambda x, shape: _t(x).reshape(*[int(s) for s in shape])
    tf.shape = lambda x: _Shape(int(s) for s in _t(x).shape)
    tf.tile = lambda x, reps: _t(x).repeat(*[int(r) for r in reps])
    tf.ones = lambda n, dtype=torch.float32: torch.ones(int(n) if not isinstance(n, (list, tuple)) else n, dtype=dtype)
    tf.range = lambda *a, dtype=None, **k: torch.arange(*a, dtype=dtype)

    def cast(x, dtype):
        return _t(x).to(dtype)  # float -> int truncates toward zero, like tf.cast

    tf.cast = cast

    def slice_(x, begin, size):
        idx = tuple(slice(int(b), None if int(s) == -1 else int(b) + int(s)) for b, s in zip(begin, size))
        return x[idx]

    tf.slice = slice_
    tf.boolean_mask = lambda x, m: x[m]
    tf.unique = lambda x: (torch.unique(x), None)
    tf.Assert = lambda *a, **k: None
    tf.control_dependencies = lambda deps: contextlib.nullcontext()
    def custom_gradient(f):
        """tf.custom_gradient: f(x) -> (y, grad_fn); backward calls grad_fn(dy)."""
        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x):
                y, g = f(x.detach())
                ctx.grad_fn_ = g
                return y

            @staticmethod
            def backward(ctx, dy):
                return ctx.grad_fn_(dy)
        return _Fn.apply

    tf.custom_gradient = custom_gradient
    tf.device = lambda name: contextlib.nullcontext()

    m = types.ModuleType("tensorflow.math")
    m.log, m.add, m.multiply, m.ceil = torch.log, torch.add, torch.mul, torch.ceil
    m.log = lambda x: torch.log(_t(x))
    m.ceil = lambda x: torch.ceil(_t(x, torch.float32))
    m.log_sigmoid = lambda x: torch.nn.functional.logsigmoid(_t(x))
    tf.multiply = lambda a, b: torch.mul(_t(a), _t(b))
    tf.pow = lambda a, b: torch.pow(_t(a), b)
    tf.abs = lambda a: torch.abs(_t(a))
    tf.math = m

    rnd = types.ModuleType("tensorflow.random")

    def uniform(shape, minval=0, maxval=None, dtype=torch.float32, seed=None, **k):
        arr = RANDOM_FIFO.pop(0)
        n = int(shape[0])
        assert arr.shape[0] == n, "shim RNG FIFO: expected %d draws, got %d" % (n, arr.shape[0])
        assert arr.min() >= minval and arr.max() < maxval
        return _t(arr, dtype)

    rnd.uniform = uniform
    rnd.set_seed = lambda s: None
    tf.random = rnd

    keras = types.ModuleType("tensorflow.keras")
    backend = types.ModuleType("tensorflow.keras.backend")
    backend.repeat = lambda x, n: x.unsqueeze(1).repeat(1, int(n), 1)
    backend.expand_dims = lambda x, axis=-1: x.unsqueeze(axis)
    keras.backend = backend
    keras.__getattr__ = lambda name: mock.MagicMock(name="tf.keras." + name)
    tf.keras = keras

    for name in ("data", "compat", "config", "optimizers", "summary", "lookup", "initializers",
                 "contrib", "train", "errors", "io", "linalg", "sparse", "debugging", "python"):
        setattr(tf, name, mock.MagicMock(name="tf." + name))
    tf.while_loop = mock.MagicMock()
    tf.__getattr__ = lambda name: mock.MagicMock(name="tf." + name)
    tf.__version__ = "2.2.3-shim"
    return tf


_LOADED = {}


def load():
    """Import the reference package (unmodified) and return a namespace of what the oracle drives."""
    if _LOADED:
        return _LOADED["ns"]
    if not available():
        raise RuntimeError("reference tree not found at " + REF_ROOT)
    tf = _build_tf()
    sys.modules["tensorflow"] = tf
    for sub in ("nn", "math", "random", "keras"):
        sys.modules["tensorflow." + sub] = getattr(tf, sub)
    sys.modules["tensorflow.keras.backend"] = tf.keras.backend
    # importlib.util.find_spec("tensorflow") (torch._dynamo probes optional packages that way) raises on a module
    # without a spec: give the shim modules one
    import importlib.machinery
    for name, mod in list(sys.modules.items()):
        if (name == "tensorflow" or name.startswith("tensorflow.")) and isinstance(mod, types.ModuleType) and getattr(mod, "__spec__", None) is None:
            mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    for sub in ("python", "python.ops", "python.framework", "keras.layers", "keras.models"):
        sys.modules.setdefault("tensorflow." + sub, mock.MagicMock())
    import pydantic.v1 as pv1  # the reference's dataset classes use the pydantic-1 API

    saved = sys.modules.get("pydantic")
    sys.modules["pydantic"] = pv1
    sys.path.insert(0, REF_ROOT)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    try:
        ns = types.SimpleNamespace()
        ns.tf = tf
        ns.models = importlib.import_module("emgraph.models")
        ns.protocol = importlib.import_module("emgraph.evaluation.protocol")
        ns.metrics = importlib.import_module("emgraph.evaluation.metrics")
        ns.numpy_adapter = importlib.import_module("emgraph.datasets.numpy_adapter")
        ns.sqlite_adapter = importlib.import_module("emgraph.datasets.sqlite_adapter")
        ns.losses = importlib.import_module("emgraph.losses")
        ns.embedding_model = importlib.import_module("emgraph.models.EmbeddingModel")
    finally:
        sys.path.remove(REF_ROOT)
        if saved is not None:
            sys.modules["pydantic"] = saved
    _LOADED["ns"] = ns
    return ns


# ------------------------------------------------------------------------------------------------
# harness: drives the reference's own methods the way _get_model_loss / the eval graph do
# ------------------------------------------------------------------------------------------------
def make_model(name, k, eta, loss, loss_params=None, embedding_model_params=None, regularizer=None, regularizer_params=None):
    ns = load()
    cls = getattr(ns.models, name)
    return cls(k=k, eta=eta, loss=loss, loss_params=loss_params or {},
               embedding_model_params=embedding_model_params or {}, regularizer=regularizer,
               regularizer_params=dict(regularizer_params or {}))


def _ref_non_linearity(ns, model, scores):
    """models/EmbeddingModel.py:679-689 / :801-812 / :1868-1881 with the reference's own ops."""
    nl = model.embedding_model_params.get("non_linearity", "linear")
    if nl == "linear":
        return scores
    if nl == "tanh":
        return ns.tf.tanh(scores)
    if nl == "sigmoid":
        return ns.tf.sigmoid(scores)
    if nl == "softplus":
        return ns.embedding_model.custom_softplus(scores)
    raise ValueError("Invalid non-linearity")


def ref_train_forward_backward(name, k, eta, loss, ent, rel, pos, keep_subj, repl,
                               loss_params=None, embedding_model_params=None, side="s,o", regularizer=None,
                               regularizer_params=None):
    """models/EmbeddingModel.py:675-677, :724-729, :788-816 executed with the reference's own
    _lookup_embeddings/_fn/generate_corruptions_for_fit/loss.apply; backward by torch autograd
    (stands in for tf.GradientTape).  Returns loss, scores, dense row gradients."""
    ns = load()
    model = make_model(name, k, eta, loss, loss_params, embedding_model_params, regularizer, regularizer_params)
    model.ent_emb = torch.tensor(ent, dtype=torch.float32, requires_grad=True)
    model.rel_emb = torch.tensor(rel, dtype=torch.float32, requires_grad=True)
    x_pos = torch.as_tensor(np.asarray(pos), dtype=torch.int32)
    e_s, e_p, e_o = model._lookup_embeddings(x_pos)
    scores_pos = _ref_non_linearity(ns, model, model._fn(e_s, e_p, e_o))
    sp_out = scores_pos.detach().numpy().copy()
    if model.loss.get_state("require_same_size_pos_neg"):
        scores_pos = ns.tf.reshape(ns.tf.tile(scores_pos, [eta]), [ns.tf.shape(scores_pos)[0] * eta])
    RANDOM_FIFO.clear()
    if side in ("s,o", "s+o"):
        RANDOM_FIFO.append(np.asarray(keep_subj, np.int64))
    RANDOM_FIFO.append(np.asarray(repl, np.int64))
    x_neg = ns.protocol.generate_corruptions_for_fit(
        x_pos, entities_list=None, eta=eta, corrupt_side=side, entities_size=ent.shape[0], rnd=0)
    assert not RANDOM_FIFO
    e_s, e_p, e_o = model._lookup_embeddings(x_neg)
    scores_neg = _ref_non_linearity(ns, model, model._fn(e_s, e_p, e_o))
    loss_t = model.loss.apply(scores_pos, scores_neg)
    if model.regularizer is not None:  # models/EmbeddingModel.py:818-820
        loss_t = loss_t + model.regularizer.apply([model.ent_emb, model.rel_emb])
    loss_t.backward()
    return dict(loss=float(loss_t.detach()), scores_pos=sp_out,
                scores_neg=scores_neg.detach().numpy().copy(), neg=x_neg.numpy().copy(),
                grad_ent=model.ent_emb.grad.numpy().copy(), grad_rel=model.rel_emb.grad.numpy().copy())


def ref_train_forward_backward_sides(name, k, eta, loss, ent, rel, pos, sides, keep_subj_list, repl_list,
                                     loss_params=None, embedding_model_params=None, regularizer=None,
                                     regularizer_params=None):
    """A LIST-valued corrupt_side (models/EmbeddingModel.py:780-816): one loss term per side, each with its own
    corruptions, summed into one loss before the single backward / optimizer step.  keep_subj_list[i] is only
    consumed for 's,o' / 's+o' entries.  Returns the summed loss, the positives' scores, per-side negative
    scores / triples and the dense row gradients of the sum."""
    ns = load()
    model = make_model(name, k, eta, loss, loss_params, embedding_model_params, regularizer, regularizer_params)
    model.ent_emb = torch.tensor(ent, dtype=torch.float32, requires_grad=True)
    model.rel_emb = torch.tensor(rel, dtype=torch.float32, requires_grad=True)
    x_pos = torch.as_tensor(np.asarray(pos), dtype=torch.int32)
    e_s, e_p, e_o = model._lookup_embeddings(x_pos)
    scores_pos = _ref_non_linearity(ns, model, model._fn(e_s, e_p, e_o))
    sp_out = scores_pos.detach().numpy().copy()
    if model.loss.get_state("require_same_size_pos_neg"):
        scores_pos = ns.tf.reshape(ns.tf.tile(scores_pos, [eta]), [ns.tf.shape(scores_pos)[0] * eta])
    loss_t = 0
    negs, sneg = [], []
    for side, keep_subj, repl in zip(sides, keep_subj_list, repl_list):
        RANDOM_FIFO.clear()
        if side in ("s,o", "s+o"):
            RANDOM_FIFO.append(np.asarray(keep_subj, np.int64))
        RANDOM_FIFO.append(np.asarray(repl, np.int64))
        x_neg = ns.protocol.generate_corruptions_for_fit(
            x_pos, entities_list=None, eta=eta, corrupt_side=side, entities_size=ent.shape[0], rnd=0)
        assert not RANDOM_FIFO
        e_s, e_p, e_o = model._lookup_embeddings(x_neg)
        scores_neg = _ref_non_linearity(ns, model, model._fn(e_s, e_p, e_o))
        loss_t = loss_t + model.loss.apply(scores_pos, scores_neg)
        negs.append(x_neg.numpy().copy())
        sneg.append(scores_neg.detach().numpy().copy())
    if model.regularizer is not None:  # models/EmbeddingModel.py:818-820 (once, after the side loop)
        loss_t = loss_t + model.regularizer.apply([model.ent_emb, model.rel_emb])
    loss_t.backward()
    return dict(loss=float(loss_t.detach()), scores_pos=sp_out, scores_neg=sneg, neg=negs,
                grad_ent=model.ent_emb.grad.numpy().copy(), grad_rel=model.rel_emb.grad.numpy().copy())


def ref_ranks(name, k, ent, rel, test, filter_triples=None, side="s,o", strategy="worst",
              embedding_model_params=None):
    """Per test triple: the body of _initialize_eval_graph small-graph branch
    (models/EmbeddingModel.py:1856-1866, :1883-1892, :1942-1986) using the reference's
    generate_corruptions_for_eval, _fn, perform_comparision and the SQLite filter adapter."""
    ns = load()
    tf = ns.tf
    model = make_model(name, k, 1, "nll", None, embedding_model_params)
    model.ent_emb = torch.tensor(ent, dtype=torch.float32)
    model.rel_emb = torch.tensor(rel, dtype=torch.float32)
    model.eval_config = {"corrupt_side": side, "ranking_strategy": strategy}
    E = ent.shape[0]
    all_ent = torch.arange(E, dtype=torch.int32)
    adapter = None
    if filter_triples is not None:
        adapter = ns.sqlite_adapter.SQLiteAdapter()
        ent_map = {i: i for i in range(E)}
        rel_map = {i: i for i in range(rel.shape[0])}
        adapter.use_mappings(rel_map, ent_map)
        adapter.set_data(np.asarray(filter_triples, dtype=np.int64), "filter", mapped_status=True)
    out = []
    try:
        for x in np.asarray(test).reshape(-1, 3):
            xt = torch.as_tensor(x[None, :].astype(np.int32))
            corr = ns.protocol.generate_corruptions_for_eval(xt, all_ent, side)
            e_s, e_p, e_o = model._lookup_embeddings(corr)
            scores_predict = _ref_non_linearity(ns, model, model._fn(e_s, e_p, e_o))
            e_s, e_p, e_o = model._lookup_embeddings(xt)
            score_positive = _ref_non_linearity(ns, model, tf.squeeze(model._fn(e_s, e_p, e_o)))
            hi_o = hi_s = 0
            if side == "s,o":
                half = scores_predict.shape[0] // 2
                obj_sc, sub_sc = scores_predict[:half], scores_predict[half:]
            if adapter is not None:
                idx_o, idx_s = adapter.get_participating_entities(x)
                idx_o = torch.as_tensor(idx_o.reshape(-1).astype(np.int64))
                idx_s = torch.as_tensor(idx_s.reshape(-1).astype(np.int64))
                if side == "s,o":
                    sp_o, sp_s = obj_sc[idx_o], sub_sc[idx_s]
                else:
                    sp_o = scores_predict[idx_o]
                    sp_s = scores_predict[idx_s + E] if side == "s+o" else scores_predict[idx_s]
                if "o" in side:
                    hi_o = int(model.perform_comparision(sp_o, score_positive))
                if "s" in side:
                    hi_s = int(model.perform_comparision(sp_s, score_positive))
            if side == "s,o":
                out.append([int(model.perform_comparision(sub_sc, score_positive)) + 1 - hi_s,
                            int(model.perform_comparision(obj_sc, score_positive)) + 1 - hi_o])
            else:
                out.append(int(model.perform_comparision(scores_predict, score_positive)) + 1 - hi_s - hi_o)
    finally:
        if adapter is not None:
            adapter.cleanup()
    return np.asarray(out)
