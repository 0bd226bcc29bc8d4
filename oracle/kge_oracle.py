"""CPU oracle for the KGE hot path (train step + filtered ranking).

TEST INFRASTRUCTURE ONLY.  Nothing under ``emgraph_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it, and there only as the checker / the CPU arm.

It is a NumPy restatement of the reference's algorithm (bi-graph/Emgraph 1.0.0-rc1).  Every
function cites the reference ``file:line`` it follows (paths relative to ``/root/reference``).
The arithmetic lives in TensorFlow 2.2 in the reference, which is absent from
``/root/reference`` and not installable here; the restatement therefore follows the reference's
*call sites* op by op.

Parity pin status
-----------------
* id mapping, eval-corruption layout, metrics, rank_score: pinned against the reference's own
  golden vectors (tests/emgraph/evaluation/test_protocol.py:418-455, :490-496;
  tests/emgraph/evaluation/test_metrics.py:6-39) in ``tests/test_oracle_golden.py``.
* scores / losses / gradients / filtered ranks: the reference holds no golden vectors for these
  (SURVEY.md section 8c).  They are pinned instead against outputs of the reference's OWN Python
  (``_fn``, ``Loss._apply``, ``generate_corruptions_for_fit/_for_eval``, ``perform_comparision``,
  ``SQLiteAdapter.get_participating_entities``) executed in the builder container through the
  ``tensorflow`` shim in ``oracle/ref_shim.py``; the resulting vectors are committed under
  ``tests/golden/`` together with the generating script ``oracle/make_golden.py``.
* Keras optimizer first-step formulas (Adam/Adagrad/SGD sparse apply) come from TF's documented
  semantics, the source being absent: "parity unpinned" for the optimizer update beyond that.
"""
from __future__ import annotations

import numpy as np

MODELS = ("TransE", "DistMult", "ComplEx", "HolE")
LOSSES = ("pairwise", "nll", "multiclass_nll")

CLIP_LO, CLIP_HI = -75.0, 75.0  # losses/_loss_constants.py:14,16
SCORE_COMPARISON_PRECISION = 1e5  # utils/constants.py:87


# --------------------------------------------------------------------------------------------
# id mapping  (evaluation/protocol.py:410-445, :662-723)
# --------------------------------------------------------------------------------------------
def create_mappings(X):
    """np.unique (sorted) ids for entities (subjects+objects) and relations.
    evaluation/protocol.py:429-445."""
    X = np.asarray(X)
    unique_ent = np.unique(np.concatenate((X[:, 0], X[:, 2])))
    unique_rel = np.unique(X[:, 1])
    rel_to_idx = dict(zip(unique_rel, range(len(unique_rel))))
    ent_to_idx = dict(zip(unique_ent, range(len(unique_ent))))
    return rel_to_idx, ent_to_idx


def to_idx(X, ent_to_idx, rel_to_idx):
    """evaluation/protocol.py:662-723 -- unseen label -> ValueError."""
    X = np.asarray(X)
    if X.ndim == 1:
        X = X[np.newaxis, :]
    out = np.empty(X.shape, dtype=np.int64)
    for c, m in ((0, ent_to_idx), (1, rel_to_idx), (2, ent_to_idx)):
        for r in range(X.shape[0]):
            v = m.get(X[r, c])
            if v is None:
                raise ValueError("Input triples include one or more concepts not present in the training set.")
            out[r, c] = v
    return out


# --------------------------------------------------------------------------------------------
# scoring functions  (models/TransE.py:208-216, DistMult.py:201, ComplEx.py:288-298, HolE.py:189)
# --------------------------------------------------------------------------------------------
def internal_k(model, k):
    """ComplEx/HolE rows are [re(k) | im(k)]  (models/ComplEx.py:224)."""
    return 2 * k if model in ("ComplEx", "HolE") else k


def score_rows(model, k, e_s, e_p, e_o, norm=1, dtype=np.float32):
    """_fn on already gathered rows [n, K] -> [n]."""
    e_s = np.asarray(e_s, dtype=dtype)
    e_p = np.asarray(e_p, dtype=dtype)
    e_o = np.asarray(e_o, dtype=dtype)
    if model == "TransE":
        u = e_s + e_p - e_o
        if norm == 1:
            return -np.sum(np.abs(u), axis=1, dtype=dtype)
        return -np.sqrt(np.sum(u * u, axis=1, dtype=dtype))
    if model == "DistMult":
        return np.sum(e_s * e_p * e_o, axis=1, dtype=dtype)
    if model in ("ComplEx", "HolE"):
        s_r, s_i = e_s[:, :k], e_s[:, k:]
        p_r, p_i = e_p[:, :k], e_p[:, k:]
        o_r, o_i = e_o[:, :k], e_o[:, k:]
        f = (
            np.sum(p_r * s_r * o_r, axis=1, dtype=dtype)
            + np.sum(p_r * s_i * o_i, axis=1, dtype=dtype)
            + np.sum(p_i * s_r * o_i, axis=1, dtype=dtype)
            - np.sum(p_i * s_i * o_r, axis=1, dtype=dtype)
        )
        if model == "HolE":
            f = dtype(2.0 / k) * f  # models/HolE.py:189
        return f
    raise ValueError(model)


def score(model, k, ent, rel, triples, norm=1, dtype=np.float32):
    """_lookup_embeddings + _fn  (models/EmbeddingModel.py:490-533, :675-677)."""
    t = np.asarray(triples).reshape(-1, 3)
    return score_rows(model, k, ent[t[:, 0]], rel[t[:, 1]], ent[t[:, 2]], norm, dtype)


def score_grad_rows(model, k, e_s, e_p, e_o, norm=1, dtype=np.float64):
    """d f / d(e_s, e_p, e_o) per triple."""
    e_s = np.asarray(e_s, dtype=dtype)
    e_p = np.asarray(e_p, dtype=dtype)
    e_o = np.asarray(e_o, dtype=dtype)
    if model == "TransE":
        u = e_s + e_p - e_o
        if norm == 1:
            g = -np.sign(u)
        else:
            nrm = np.sqrt(np.sum(u * u, axis=1, keepdims=True))
            g = -u / nrm
        return g, g, -g
    if model == "DistMult":
        return e_p * e_o, e_s * e_o, e_s * e_p
    s_r, s_i = e_s[:, :k], e_s[:, k:]
    p_r, p_i = e_p[:, :k], e_p[:, k:]
    o_r, o_i = e_o[:, :k], e_o[:, k:]
    g_s = np.concatenate([p_r * o_r + p_i * o_i, p_r * o_i - p_i * o_r], axis=1)
    g_o = np.concatenate([p_r * s_r - p_i * s_i, p_r * s_i + p_i * s_r], axis=1)
    g_p = np.concatenate([s_r * o_r + s_i * o_i, s_r * o_i - s_i * o_r], axis=1)
    if model == "HolE":
        c = 2.0 / k
        g_s, g_p, g_o = c * g_s, c * g_p, c * g_o
    return g_s, g_p, g_o


# --------------------------------------------------------------------------------------------
# training corruptions  (evaluation/protocol.py:586-659)
# --------------------------------------------------------------------------------------------
def corruptions_for_fit(X, eta, keep_subj, repl):
    """Row j*n+i corrupts positive i (tile, :598); keep_subj=1 -> object replaced (:643-653)."""
    X = np.asarray(X)
    n = X.shape[0]
    ds = np.tile(X.reshape(-1), eta).reshape(n * eta, 3)
    ks = np.asarray(keep_subj).astype(ds.dtype)
    ko = 1 - ks
    repl = np.asarray(repl).astype(ds.dtype)
    subj = ks * ds[:, 0] + ko * repl
    obj = ko * ds[:, 2] + ks * repl
    return np.stack([subj, ds[:, 1], obj], axis=1)


def side_mask(side, n_eta, rng=None):
    """'s,o' == 's+o' -> Bernoulli(1/2); 'o' keeps the subject; 's' keeps the object (:587-608)."""
    if side in ("s,o", "s+o"):
        return rng.integers(0, 2, size=n_eta).astype(np.uint8)
    if side == "o":
        return np.ones(n_eta, np.uint8)
    if side == "s":
        return np.zeros(n_eta, np.uint8)
    raise ValueError("Invalid argument value {} for corruption side passed for evaluation.".format(side))


# --------------------------------------------------------------------------------------------
# losses  (losses/pairwise.py:66-70, nll.py:55-59, nll_multiclass.py:70-81, utils.py:44-53)
# --------------------------------------------------------------------------------------------
def non_linearity(name, x, dtype=np.float32):
    """(g(x), g'(x)) of embedding_model_params['non_linearity'] (models/EmbeddingModel.py:679-689, :801-812);
    'softplus' is the reference's custom_softplus log(1 + 9999*exp(x)), gradient 1 - 1/(1 + 9999*exp(x)) (:89-96)."""
    x = np.asarray(x, dtype=dtype)
    if name in (None, "linear"):
        return x, np.ones_like(x)
    if name == "tanh":
        t = np.tanh(x)
        return t, 1 - t * t
    if name == "sigmoid":
        g = 1 / (1 + np.exp(-x))
        return g, g * (1 - g)
    if name == "softplus":
        e = dtype(9999) * np.exp(x)
        return np.log(1 + e), 1 - 1 / (1 + e)
    raise ValueError("Invalid non-linearity")


def _clip(x):
    return np.clip(x, CLIP_LO, CLIP_HI)


def loss_and_dscore(loss, pos, neg, eta, margin=1.0, dtype=np.float32, alpha=0.5):
    """Loss value and dL/dpos [n], dL/dneg [eta*n].  pos is NOT tiled on entry; the eta-tiling of
    models/EmbeddingModel.py:724-729 is applied here for pairwise / nll / absolute_margin
    (require_same_size_pos_neg, losses/utils.py:24); multiclass_nll and self_adversarial take the
    positives untiled (losses/nll_multiclass.py:7, self_adversarial.py:12)."""
    pos = np.asarray(pos, dtype=dtype)
    neg = np.asarray(neg, dtype=dtype)
    n = pos.shape[0]
    negm = neg.reshape(eta, n)
    if loss == "pairwise":
        t = dtype(margin) - pos[None, :] + negm
        val = np.sum(np.maximum(t, 0), dtype=dtype)
        act = (t >= 0).astype(dtype)  # tf.maximum: gradient to first arg when x >= y
        return val, -act.sum(axis=0), act.reshape(-1)
    if loss == "nll":
        cp, cn = _clip(pos), _clip(negm)
        val = dtype(eta) * np.sum(np.log(1 + np.exp(-cp)), dtype=dtype) + np.sum(np.log(1 + np.exp(cn)), dtype=dtype)
        in_p = ((pos >= CLIP_LO) & (pos <= CLIP_HI)).astype(dtype)
        in_n = ((negm >= CLIP_LO) & (negm <= CLIP_HI)).astype(dtype)
        dpos = -dtype(eta) * (1 / (1 + np.exp(cp))) * in_p
        dneg = (1 / (1 + np.exp(-cn))) * in_n
        return val, dpos, dneg.reshape(-1)
    if loss == "multiclass_nll":
        cp, cn = _clip(pos), _clip(negm)
        pe, ne = np.exp(cp), np.exp(cn)
        z = ne.sum(axis=0, dtype=dtype) + pe
        val = -np.sum(np.log(pe / z), dtype=dtype)
        in_p = ((pos >= CLIP_LO) & (pos <= CLIP_HI)).astype(dtype)
        in_n = ((negm >= CLIP_LO) & (negm <= CLIP_HI)).astype(dtype)
        dpos = -(1 - pe / z) * in_p
        dneg = (ne / z[None, :]) * in_n
        return val, dpos, dneg.reshape(-1)
    if loss == "absolute_margin":
        # losses/absolute_margin.py:69 : sum(max(margin + neg, 0) - pos_tiled)
        t = dtype(margin) + negm
        val = np.sum(np.maximum(t, 0) - pos[None, :], dtype=dtype)
        act = (t >= 0).astype(dtype)
        return val, np.full(n, -dtype(eta), dtype=dtype), act.reshape(-1)
    if loss == "self_adversarial":
        # losses/self_adversarial.py:97-110 : p = softmax(alpha*neg) over the eta axis (gradient flows
        # through p); loss = sum -logsigmoid(margin + pos) - sum p * logsigmoid(-neg - margin)
        a, mg = dtype(alpha), dtype(margin)
        zmax = (a * negm).max(axis=0, keepdims=True)
        ex = np.exp(a * negm - zmax)
        p = ex / ex.sum(axis=0, keepdims=True, dtype=dtype)

        def logsig(x):
            return np.minimum(x, 0) - np.log1p(np.exp(-np.abs(x)))

        def sig(x):
            return 1 / (1 + np.exp(-x))
        ln = logsig(-negm - mg)
        val = np.sum(-logsig(mg + pos), dtype=dtype) - np.sum(p * ln, dtype=dtype)
        lbar = np.sum(p * ln, axis=0, keepdims=True, dtype=dtype)
        dneg = p * sig(negm + mg) - a * p * (ln - lbar)
        dpos = -sig(-(mg + pos))
        return val, dpos.astype(dtype), dneg.reshape(-1).astype(dtype)
    raise ValueError("Unsupported loss function: {}".format(loss))


# --------------------------------------------------------------------------------------------
# optimizer step as the reference executes it: fresh Keras optimizer per batch (F5)
# (training/adam.py:45-46, adagrad.py:42-45, momentum.py:63-68, sgd.py:97-124)
# --------------------------------------------------------------------------------------------
ADAM_B1, ADAM_B2, KERAS_EPS = 0.9, 0.999, 1e-7
ADAGRAD_ACC0 = 0.1


def optimizer_step(opt, w, g, touched, lr, state=None, step=1, momentum=0.9, dtype=np.float32):
    """Update rows `touched` (bool [rows]) of w in place-free style; returns (w_new, state_new).
    state=None => the reference's fresh-state step (F5).  Persistent state => the engine's lazy
    stateful mode (rows without gradient skipped)."""
    w = np.array(w, dtype=dtype)
    g = np.asarray(g, dtype=dtype)
    t = np.asarray(touched, dtype=bool)
    lr = dtype(lr)
    if opt == "adam":
        m, v = (np.zeros_like(w), np.zeros_like(w)) if state is None else (state[0].copy(), state[1].copy())
        b1, b2 = dtype(ADAM_B1), dtype(ADAM_B2)
        m[t] = b1 * m[t] + (1 - b1) * g[t]
        v[t] = b2 * v[t] + (1 - b2) * g[t] * g[t]
        lr_t = lr * dtype(np.sqrt(1.0 - ADAM_B2 ** step) / (1.0 - ADAM_B1 ** step))
        w[t] = w[t] - lr_t * m[t] / (np.sqrt(v[t]) + dtype(KERAS_EPS))
        return w, (m, v)
    if opt == "adagrad":
        a = np.full_like(w, ADAGRAD_ACC0) if state is None else state[0].copy()
        a[t] = a[t] + g[t] * g[t]
        w[t] = w[t] - lr * g[t] / (np.sqrt(a[t]) + dtype(KERAS_EPS))
        return w, (a,)
    if opt == "momentum":
        vel = np.zeros_like(w) if state is None else state[0].copy()
        vel[t] = dtype(momentum) * vel[t] - lr * g[t]
        w[t] = w[t] + vel[t]
        return w, (vel,)
    if opt == "sgd":
        w[t] = w[t] - lr * g[t]
        return w, ()
    raise ValueError("Unsupported optimizer: {}".format(opt))


# --------------------------------------------------------------------------------------------
# one training step  (models/EmbeddingModel.py:614-822 + optimizer.minimize :1415-1418)
# --------------------------------------------------------------------------------------------
def train_step(model, k, loss, eta, ent, rel, pos, keep_subj, repl, margin=1.0, norm=1,
               opt=None, lr=5e-4, state=None, step=1, dtype=np.float32, grad_dtype=np.float64, alpha=0.5,
               reg_p=0, reg_lambda_ent=0.0, reg_lambda_rel=0.0, nl="linear"):
    """Forward in `dtype` (the reference is fp32), gradients in `grad_dtype`.
    Returns dict(loss, scores_pos, scores_neg, grad_ent[E,K], grad_rel[R,K], touched_ent, touched_rel
    [, ent_new, rel_new, state_ent, state_rel])."""
    pos = np.asarray(pos).reshape(-1, 3)
    n = pos.shape[0]
    neg = corruptions_for_fit(pos, eta, keep_subj, repl)
    sp = score(model, k, ent, rel, pos, norm, dtype)
    sn = score(model, k, ent, rel, neg, norm, dtype)
    sp, gsp = non_linearity(nl, sp, dtype)
    sn, gsn = non_linearity(nl, sn, dtype)
    val, dpos, dneg = loss_and_dscore(loss, sp, sn, eta, margin, dtype, alpha)
    dpos, dneg = dpos * gsp, dneg * gsn
    gd = grad_dtype
    g_ent = np.zeros(ent.shape, dtype=gd)
    g_rel = np.zeros(rel.shape, dtype=gd)
    for trip, dsc in ((pos, dpos), (neg, dneg)):
        gs, gp, go = score_grad_rows(model, k, ent[trip[:, 0]], rel[trip[:, 1]], ent[trip[:, 2]], norm, gd)
        w = np.asarray(dsc, dtype=gd)[:, None]
        np.add.at(g_ent, trip[:, 0], w * gs)
        np.add.at(g_ent, trip[:, 2], w * go)
        np.add.at(g_rel, trip[:, 1], w * gp)
    t_ent = np.zeros(ent.shape[0], bool)
    t_ent[pos[:, 0]] = True
    t_ent[pos[:, 2]] = True
    t_ent[neg[:, 0]] = True
    t_ent[neg[:, 2]] = True
    t_rel = np.zeros(rel.shape[0], bool)
    t_rel[pos[:, 1]] = True
    if reg_p > 0 and (reg_lambda_ent != 0 or reg_lambda_rel != 0):
        # regularizers/lp.py:106-111 : loss += lambda_i * sum(|param_i|^p) over the whole tables
        # (models/EmbeddingModel.py:818-820) => every row has a gradient
        for w, gacc, lam in ((ent, g_ent, reg_lambda_ent), (rel, g_rel, reg_lambda_rel)):
            w64 = np.asarray(w, dtype=gd)
            val = dtype(val + dtype(lam) * np.sum(np.abs(np.asarray(w, dtype=dtype)) ** reg_p, dtype=dtype))
            gacc += lam * reg_p * np.abs(w64) ** (reg_p - 1) * np.sign(w64)
        t_ent[:] = True
        t_rel[:] = True
    out = dict(loss=val, scores_pos=sp, scores_neg=sn, grad_ent=g_ent, grad_rel=g_rel,
               touched_ent=t_ent, touched_rel=t_rel, neg=neg)
    if opt is not None:
        st_e = None if state is None else state[0]
        st_r = None if state is None else state[1]
        out["ent_new"], out["state_ent"] = optimizer_step(opt, ent, g_ent, t_ent, lr, st_e, step, dtype=dtype)
        out["rel_new"], out["state_rel"] = optimizer_step(opt, rel, g_rel, t_rel, lr, st_r, step, dtype=dtype)
    return out


# --------------------------------------------------------------------------------------------
# the engine's own corruption stream (no reference counterpart: the reference draws from TF's RNG, whose stream
# is a parity INPUT -- SURVEY 8a row a11).  Philox4x32-10 (Salmon et al., SC'11; Random123), keyed as
# emgraph_b200/csrc/kge_train.cu:kge_emit_kernel keys it, so that throughput runs are checkable too.
# --------------------------------------------------------------------------------------------
def philox4x32_10(counter, key):
    """counter [..., 4], key [..., 2] uint32 -> [..., 4] uint32."""
    c = [np.asarray(counter)[..., i].astype(np.uint64) for i in range(4)]
    k0 = np.asarray(key)[..., 0].astype(np.uint64)
    k1 = np.asarray(key)[..., 1].astype(np.uint64)
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(MASK)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & np.uint64(MASK)
        k1 = (k1 + np.uint64(W1)) & np.uint64(MASK)
    return np.stack(c, axis=-1).astype(np.uint32)


def draw_corruptions(seed, step, n, eta, E, side="s,o", neg_index_base=0, neg_entities=None, neg_entities_n=0,
                     keep_codes=None):
    """(repl[eta*n] int32, keep_subj[eta*n] uint8) as the engine draws them for negative q = j*n + i:
    counter = (g_lo, g_hi, step_lo, step_hi) with g = neg_index_base + q, key = (seed_lo, seed_hi);
    replacement = mulhi(word0, #candidates) (uniform over [0,E), the first neg_entities_n ids, or the list
    neg_entities, evaluation/protocol.py:610-641); side coin = top bit of word1 for 's,o' (:600-604), fixed for
    's' / 'o' (:605-608); keep_codes[q] in {0,1} overrides the side of negative q (>= 2: by `side`)."""
    q = np.arange(n * eta, dtype=np.uint64) + np.uint64(neg_index_base)
    ctr = np.stack([q & np.uint64(0xFFFFFFFF), q >> np.uint64(32),
                    np.full_like(q, int(step) & 0xFFFFFFFF), np.full_like(q, int(step) >> 32)], -1)
    key = np.stack([np.full_like(q, int(seed) & 0xFFFFFFFF), np.full_like(q, int(seed) >> 32)], -1)
    o = philox4x32_10(ctr, key).astype(np.uint64)
    ncand = len(neg_entities) if neg_entities is not None else (int(neg_entities_n) if neg_entities_n else int(E))
    repl = ((o[:, 0] * np.uint64(ncand)) >> np.uint64(32)).astype(np.int64)
    if neg_entities is not None:
        repl = np.asarray(neg_entities)[repl]
    if side in ("s,o", "s+o"):
        keep = (o[:, 1] >> np.uint64(31)).astype(np.uint8)
    else:
        keep = side_mask(side, n * eta)
    if keep_codes is not None:
        kc = np.asarray(keep_codes).astype(np.uint8)
        keep = np.where(kc < 2, kc, keep).astype(np.uint8)
    return repl.astype(np.int32), keep


def stack_sides(pos, eta, sides, keep_subj_list, repl_list):
    """A LIST-valued corrupt_side (models/EmbeddingModel.py:780-816) sums one loss term per side, each over its
    own corruptions of the SAME positives, before the single optimizer step.  Every loss is a sum of
    per-positive terms over that positive's eta negatives, so the summed loss equals the loss of ONE batch in
    which the positives appear once per side: positives [pos ; pos ; ...] (n' = S*n) and negative row
    j*n' + s*n + i = side s's negative (j, i).  Returns (pos', keep_subj', repl') in that layout."""
    pos = np.asarray(pos).reshape(-1, 3)
    n, S = pos.shape[0], len(sides)
    keep = np.zeros((eta, S, n), np.uint8)
    repl = np.zeros((eta, S, n), np.int32)
    for s, side in enumerate(sides):
        ks = np.asarray(keep_subj_list[s]) if side in ("s,o", "s+o") else side_mask(side, n * eta)
        keep[:, s, :] = ks.reshape(eta, n)
        repl[:, s, :] = np.asarray(repl_list[s]).reshape(eta, n)
    return np.tile(pos, (S, 1)), keep.reshape(-1), repl.reshape(-1)


def train_step_sides(model, k, loss, eta, ent, rel, pos, sides, keep_subj_list, repl_list, **kw):
    """train_step for a list of corruption sides (one summed loss, one optimizer step)."""
    pos2, keep2, repl2 = stack_sides(pos, eta, sides, keep_subj_list, repl_list)
    return train_step(model, k, loss, eta, ent, rel, pos2, keep2, repl2, **kw)


def fit_emulation(model, k, eta, epochs, batches_count, seed, loss, opt, lr, Xi, E, R, margin=1.0, norm=1, alpha=0.5,
                  reg_p=0, reg_lambda_ent=0.0, reg_lambda_rel=0.0, nl="linear", side="s,o"):
    """The engine's fit() loop restated on the oracle: Glorot-uniform tables from RandomState(seed) (entities first,
    initializers/glorot_uniform.py:59-99 as emgraph_b200/models.py:_init_table draws them), sequential unshuffled
    batches of ceil(N / batches_count) positives (datasets/numpy_adapter.py:105-111), the engine's own corruption
    stream per (seed, step) (draw_corruptions), one train_step per batch with PERSISTENT optimizer state (the
    engine's default).  Returns (ent, rel, per-epoch summed loss)."""
    K = internal_k(model, k)
    rnd = np.random.RandomState(seed)
    lim_e, lim_r = np.sqrt(6.0 / (E + K)), np.sqrt(6.0 / (R + K))
    ent = rnd.uniform(-lim_e, lim_e, size=(E, K)).astype(np.float32)
    rel = rnd.uniform(-lim_r, lim_r, size=(R, K)).astype(np.float32)

    def init_state(w):
        if opt == "adam":
            return (np.zeros_like(w), np.zeros_like(w))
        if opt == "adagrad":
            return (np.full_like(w, ADAGRAD_ACC0),)
        if opt == "momentum":
            return (np.zeros_like(w),)
        return ()

    state = (init_state(ent), init_state(rel))
    Xi = np.asarray(Xi).reshape(-1, 3)
    N = Xi.shape[0]
    bs = int(np.ceil(N / batches_count))
    step, losses = 0, []
    for _ in range(epochs):
        tot = 0.0
        for b in range(batches_count):
            pos = Xi[b * bs:min(N, (b + 1) * bs)]
            if pos.shape[0] == 0:
                continue
            step += 1
            repl, keep = draw_corruptions(seed, step, pos.shape[0], eta, E, side)
            o = train_step(model, k, loss, eta, ent, rel, pos, keep, repl, margin=margin, norm=norm, opt=opt, lr=lr,
                           state=state, step=step, alpha=alpha, reg_p=reg_p, reg_lambda_ent=reg_lambda_ent,
                           reg_lambda_rel=reg_lambda_rel, nl=nl)
            ent, rel = o["ent_new"], o["rel_new"]
            state = (o["state_ent"], o["state_rel"])
            tot += float(o["loss"])
        losses.append(tot)
    return ent, rel, losses


# --------------------------------------------------------------------------------------------
# evaluation corruptions + filter sets + ranking
# --------------------------------------------------------------------------------------------
def corruptions_for_eval(x, entities, side="s,o"):
    """[ (s,p,e) for all e ; (e,p,o) for all e ]  (evaluation/protocol.py:448-528)."""
    x = np.asarray(x).reshape(3)
    ents = np.asarray(entities).reshape(-1)
    if side == "s,o":
        side = "s+o"
    if side not in ("s+o", "s", "o"):
        raise ValueError("Invalid argument value for corruption side passed for evaluation")
    n = ents.shape[0]
    obj_sweep = np.stack([np.full(n, x[0]), np.full(n, x[1]), ents], axis=1)
    sub_sweep = np.stack([ents, np.full(n, x[1]), np.full(n, x[2])], axis=1)
    if side == "s+o":
        return np.concatenate([obj_sweep, sub_sweep], axis=0)
    return obj_sweep if side == "o" else sub_sweep


class FilterIndex:
    """(s,p)->objects and (p,o)->subjects, deduplicated, self always included
    (datasets/sqlite_adapter.py:472-489: ``select o UNION select distinct object ...``)."""

    def __init__(self, filter_triples):
        self.sp = {}
        self.po = {}
        for s, p, o in np.asarray(filter_triples).reshape(-1, 3):
            self.sp.setdefault((int(s), int(p)), set()).add(int(o))
            self.po.setdefault((int(p), int(o)), set()).add(int(s))

    def participating(self, x):
        s, p, o = (int(v) for v in np.asarray(x).reshape(3))
        objs = sorted(self.sp.get((s, p), set()) | {o})
        subs = sorted(self.po.get((p, o), set()) | {s})
        return np.asarray(objs, np.int64), np.asarray(subs, np.int64)


def quantise(x):
    """tf.cast(x * 1e5, tf.int32): fp32 multiply then truncate toward zero
    (models/EmbeddingModel.py:2010-2014)."""
    return (np.asarray(x, np.float32) * np.float32(SCORE_COMPARISON_PRECISION)).astype(np.int32)


def compare(score_corr, score_pos, strategy="worst"):
    """perform_comparision  (models/EmbeddingModel.py:1989-2033)."""
    assert strategy in ("worst", "best", "middle"), "Invalid score comparision type!"
    c, p = quantise(score_corr), quantise(score_pos)
    if strategy == "best":
        return int(np.sum(c > p))
    if strategy == "middle":
        return int(np.sum(c > p)) + int(np.ceil(np.sum(c == p) / 2))
    return int(np.sum(c >= p))


def rank_one(model, k, ent, rel, x, filt=None, side="s,o", strategy="worst", norm=1, dtype=np.float32, subset=None,
             nl="linear"):
    """Per-test-triple rank (models/EmbeddingModel.py:1856-1866, :1883-1892, :1942-1986).

    subset: optional entity ids used to generate the corruptions (eval_config['corruption_entities'],
    :1845-1857); the filter indices are then re-mapped to positions in that list and the entities outside it
    masked out (:1898-1940).  That branch of the reference is dead code under TF2 (tf.contrib hash table,
    SURVEY F3), so it is restated from its body and NOT pinned against reference-executed outputs."""
    E = ent.shape[0]
    x = np.asarray(x).reshape(3)
    cand = np.arange(E) if subset is None else np.asarray(subset, dtype=np.int64).reshape(-1)
    C = cand.shape[0]
    corr = corruptions_for_eval(x, cand, side)
    sc = non_linearity(nl, score(model, k, ent, rel, corr, norm, dtype), dtype)[0]  # models/EmbeddingModel.py:1868-1881
    sp = non_linearity(nl, score(model, k, ent, rel, x, norm, dtype), dtype)[0][0]
    hi_o = hi_s = 0
    if filt is not None:
        idx_o, idx_s = filt.participating(x)
        if subset is not None:
            pos_of = {int(e): i for i, e in enumerate(cand)}
            idx_o = np.asarray([pos_of[int(e)] for e in idx_o if int(e) in pos_of], np.int64)
            idx_s = np.asarray([pos_of[int(e)] for e in idx_s if int(e) in pos_of], np.int64)
    if side == "s,o":
        obj_sc, sub_sc = sc[:C], sc[C:]
        if filt is not None:
            hi_o = compare(obj_sc[idx_o], sp, strategy)
            hi_s = compare(sub_sc[idx_s], sp, strategy)
        return [compare(sub_sc, sp, strategy) + 1 - hi_s, compare(obj_sc, sp, strategy) + 1 - hi_o]
    if filt is not None:
        if side in ("o", "s+o"):
            hi_o = compare(sc[idx_o], sp, strategy)
        if side == "s+o":
            hi_s = compare(sc[idx_s + C], sp, strategy)
        elif side == "s":
            hi_s = compare(sc[idx_s], sp, strategy)
    return compare(sc, sp, strategy) + 1 - hi_s - hi_o


def ranks(model, k, ent, rel, test, filter_triples=None, side="s,o", strategy="worst", norm=1, dtype=np.float32,
          subset=None, nl="linear"):
    """Intended semantics of get_ranks: the per-triple graph evaluated for EVERY test triple
    (models/EmbeddingModel.py:2046-2099)."""
    filt = FilterIndex(filter_triples) if filter_triples is not None else None
    return np.asarray([rank_one(model, k, ent, rel, x, filt, side, strategy, norm, dtype, subset, nl)
                       for x in np.asarray(test).reshape(-1, 3)])


def sweep_scores(model, k, ent, rel, x, norm=1, dtype=np.float64):
    """All-entity object-side and subject-side scores + positive score for one triple, in `dtype`.
    Used by the tie classifier: a rank mismatch is admissible only if some candidate score sits
    within tolerance of the positive's quantisation boundary."""
    E = ent.shape[0]
    corr = corruptions_for_eval(x, np.arange(E), "s+o")
    sc = score(model, k, ent.astype(dtype), rel.astype(dtype), corr, norm, dtype)
    sp = score(model, k, ent.astype(dtype), rel.astype(dtype), x, norm, dtype)[0]
    return sc[:E], sc[E:], sp


# --------------------------------------------------------------------------------------------
# metrics  (evaluation/metrics.py:66-67, :129-130, :161-164, :221-222)
# --------------------------------------------------------------------------------------------
def hits_at_n_score(ranks_, n):
    r = np.asarray(ranks_).reshape(-1)
    return np.sum(r <= n) / len(r)


def mrr_score(ranks_):
    r = np.asarray(ranks_).reshape(-1)
    return np.sum(1 / r) / len(r)


def mr_score(ranks_):
    r = np.asarray(ranks_).reshape(-1)
    return np.sum(r) / len(r)


def rank_score(y_true, y_pred, pos_lab=1):
    idx = np.argsort(y_pred)[::-1]
    y_ord = np.asarray(y_true)[idx]
    return int(np.where(y_ord == pos_lab)[0][0] + 1)


# --------------------------------------------------------------------------------------------
# synthetic graphs of the benchmark shapes -- shared by tests and bench
# --------------------------------------------------------------------------------------------
def synthetic_triples(E, R, N, seed=0, zipf=False, cover=True):
    """Unique (s,p,o) int32 triples, no self loops, every entity appearing at least once."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if zipf:
        w = 1.0 / np.arange(1, E + 1)
        w /= w.sum()
        perm = rng.permutation(E)
        cdf = np.cumsum(w)

        def draw(m):
            return perm[np.minimum(np.searchsorted(cdf, rng.random(m)), E - 1)]
    else:
        def draw(m):
            return rng.integers(0, E, size=m)
    parts = []
    have = 0
    if cover:
        chain = np.stack([np.arange(E), rng.integers(0, R, size=E), (np.arange(E) + 1) % E], axis=1)
        parts.append(chain[: min(E, N)])
        have = parts[0].shape[0]
    seen = None
    out = np.concatenate(parts, axis=0) if parts else np.zeros((0, 3), np.int64)
    while True:
        key = (out[:, 0].astype(np.int64) * R + out[:, 1]) * E + out[:, 2]
        _, first = np.unique(key, return_index=True)
        out = out[np.sort(first)]
        if out.shape[0] >= N:
            break
        m = int((N - out.shape[0]) * 1.2) + 16
        s, o, p = draw(m), draw(m), rng.integers(0, R, size=m)
        keep = s != o
        out = np.concatenate([out, np.stack([s[keep], p[keep], o[keep]], axis=1)], axis=0)
    del seen, have
    return out[:N].astype(np.int32)


def glorot_uniform(rows, cols, seed):
    """U(+-sqrt(6/(rows+cols)))  (initializers/glorot_uniform.py:59-99)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lim = np.sqrt(6.0 / (rows + cols))
    return rng.uniform(-lim, lim, size=(rows, cols)).astype(np.float32)
