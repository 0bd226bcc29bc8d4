"""TEST INFRASTRUCTURE (never imported by emgraph_b200/): NumPy emulation of the row-sharded training step in its
OWNER-COMPUTE form -- the exchange DESIGN.md section 7 plans for round 2 -- with every byte that would cross
NVLink counted.

W simulated ranks own contiguous row ranges of the entity table (emgraph_b200/distributed.py:shard_range); every
rank brings its own batch.  Instead of fetching the rows of its negatives' replacement entities (the current
"owner push": (2+eta) rows per positive), a rank ships the two folded query rows of each positive to the owners,
the owners score their own candidates and later reduce `c * dF/dr` locally and send back one partial sum of
`c * dF/dQ` per (positive, side, owner).  The functions below prove that this decomposition reproduces the
single-process oracle step (oracle/kge_oracle.py:train_step, which restates the reference's
models/EmbeddingModel.py:614-822) and give the per-leg volumes the design is sized with.

Candidate scores from a folded query:
  trilinear models  F(Q, r) = Q . r            (DistMult / ComplEx / HolE; HolE's 2/k is folded into Q)
  TransE            F(Q, r) = -||Q - r||_p     (object side Q = s + p; subject side Q = o - p)
"""
from __future__ import annotations

import numpy as np

from . import kge_oracle as ko


def shard_range(E, W, r):
    rps = (E + W - 1) // W
    return min(E, r * rps), min(E, (r + 1) * rps)


def fold_queries(model, k, e_s, e_p, e_o, dtype=np.float64):
    """(Qo, Qs): S_o[e] = F(Qo, E_e), S_s[e] = F(Qs, E_e)  (csrc/kge_rank.cu:kge_rank_prepare_kernel)."""
    e_s, e_p, e_o = (np.asarray(x, dtype=dtype) for x in (e_s, e_p, e_o))
    if model == "TransE":
        return e_s + e_p, e_o - e_p
    if model == "DistMult":
        return e_s * e_p, e_p * e_o
    s_r, s_i, p_r, p_i, o_r, o_i = e_s[:, :k], e_s[:, k:], e_p[:, :k], e_p[:, k:], e_o[:, :k], e_o[:, k:]
    c = 2.0 / k if model == "HolE" else 1.0
    qo = np.concatenate([p_r * s_r - p_i * s_i, p_r * s_i + p_i * s_r], axis=1)
    qs = np.concatenate([p_r * o_r + p_i * o_i, p_r * o_i - p_i * o_r], axis=1)
    return c * qo, c * qs


def query_score(model, Q, r, norm=1):
    if model == "TransE":
        u = Q - r
        return -np.sum(np.abs(u), axis=1) if norm == 1 else -np.sqrt(np.sum(u * u, axis=1))
    return np.sum(Q * r, axis=1)


def query_score_grads(model, Q, r, norm=1):
    """(dF/dQ, dF/dr) rows."""
    if model == "TransE":
        u = Q - r
        g = -np.sign(u) if norm == 1 else -u / np.sqrt(np.sum(u * u, axis=1, keepdims=True))
        return g, -g
    return r, Q


def unfold_query_grads(model, k, e_s, e_p, e_o, gQo, gQs, dtype=np.float64):
    """Chain dL/dQo, dL/dQs back to the positive's own rows: (g_s, g_p, g_o)."""
    e_s, e_p, e_o = (np.asarray(x, dtype=dtype) for x in (e_s, e_p, e_o))
    if model == "TransE":
        return gQo, gQo - gQs, gQs
    if model == "DistMult":
        return gQo * e_p, gQo * e_s + gQs * e_o, gQs * e_p
    c = 2.0 / k if model == "HolE" else 1.0
    s_r, s_i, p_r, p_i, o_r, o_i = e_s[:, :k], e_s[:, k:], e_p[:, :k], e_p[:, k:], e_o[:, :k], e_o[:, k:]
    a_r, a_i = c * gQo[:, :k], c * gQo[:, k:]  # Qo = c * [p_r s_r - p_i s_i | p_r s_i + p_i s_r]
    b_r, b_i = c * gQs[:, :k], c * gQs[:, k:]  # Qs = c * [p_r o_r + p_i o_i | p_r o_i - p_i o_r]
    g_s = np.concatenate([a_r * p_r + a_i * p_i, -a_r * p_i + a_i * p_r], axis=1)
    g_o = np.concatenate([b_r * p_r - b_i * p_i, b_r * p_i + b_i * p_r], axis=1)
    g_p = np.concatenate([a_r * s_r + a_i * s_i + b_r * o_r + b_i * o_i, -a_r * s_i + a_i * s_r + b_r * o_i - b_i * o_r], axis=1)
    return g_s, g_p, g_o


def owner_compute_step(model, k, loss, eta, ent, rel, batches, W, margin=1.0, norm=1, alpha=0.5, nl="linear"):
    """One data-parallel step over W ranks.  batches[w] = (pos [n,3], keep_subj [eta*n], repl [eta*n]) of rank w (all
    ranks the same n).  Returns dict(loss, grad_ent, grad_rel, bytes={leg: bytes received per rank, worst rank}).

    Legs (row = 4K bytes):
      rows_in      subject / object rows of the positives fetched from their owners          2n rows      * (remote share)
      queries_in   all-gather of every rank's [Qo | Qs]                                      2n(W-1) rows
      scores_in    scores of this rank's negatives computed by remote owners                 4 B each
      coefs_in     all-gather of dL/dscore of every negative                                 4 eta n (W-1) B
      partials_in  per (positive, side, remote owner) partial sums of c * dF/dQ              <= 2n(W-1) rows
      posgrad_in   gradient rows of remote batches' positives whose s / o rows this rank owns  ~2n rows * (remote share)
    """
    E, K = ent.shape
    gd = np.float64
    ent64, rel64 = np.asarray(ent, gd), np.asarray(rel, gd)
    n = batches[0][0].shape[0]
    owner = lambda ids: np.asarray(ids) // ((E + W - 1) // W)  # noqa: E731
    row_b = 4 * K
    by = {leg: np.zeros(W, np.int64) for leg in ("rows_in", "queries_in", "scores_in", "coefs_in", "partials_in", "posgrad_in")}

    # 1. positives' rows -> folded queries, positive scores (batch owner)
    st = []
    for w, (pos, keep, repl) in enumerate(batches):
        pos = np.asarray(pos).reshape(-1, 3)
        e_s, e_p, e_o = ent64[pos[:, 0]], rel64[pos[:, 1]], ent64[pos[:, 2]]
        by["rows_in"][w] += row_b * (np.count_nonzero(owner(pos[:, 0]) != w) + np.count_nonzero(owner(pos[:, 2]) != w))
        Qo, Qs = fold_queries(model, k, e_s, e_p, e_o)
        sp = ko.score_rows(model, k, e_s, e_p, e_o, norm, gd)
        st.append(dict(pos=pos, keep=np.asarray(keep).astype(bool), repl=np.asarray(repl), e=(e_s, e_p, e_o), Qo=Qo, Qs=Qs, sp=sp))
    # 2. all-gather of the queries
    by["queries_in"][:] = 2 * n * row_b * (W - 1)
    # 3. owners score their candidates; scores travel back to the batch owner
    for w, b in enumerate(st):
        i = np.tile(np.arange(n), eta)
        Q = np.where(b["keep"][:, None], b["Qo"][i], b["Qs"][i])  # keep_subj -> the object is replaced -> object-side query
        b["Q"], b["i"] = Q, i
        b["sn"] = query_score(model, Q, ent64[b["repl"]], norm)  # evaluated by owner(repl), row by row
        by["scores_in"][w] += 4 * np.count_nonzero(owner(b["repl"]) != w)
    # 4. loss and dL/dscore at the batch owner (losses/*.py through the oracle), all-gather of the coefficients
    loss_total = 0.0
    for b in st:
        sp, gsp = ko.non_linearity(nl, b["sp"], gd)
        sn, gsn = ko.non_linearity(nl, b["sn"], gd)
        val, dpos, dneg = ko.loss_and_dscore(loss, sp, sn, eta, margin, gd, alpha)
        b["dpos"], b["c"] = np.asarray(dpos, gd) * gsp, np.asarray(dneg, gd) * gsn
        loss_total += float(val)
    by["coefs_in"][:] = 4 * eta * n * (W - 1)
    # 5. owners: gradient of their candidate rows + partial sums of c * dF/dQ per (rank, positive, side)
    g_ent = np.zeros((E, K), gd)
    g_rel = np.zeros(rel.shape, gd)
    for w, b in enumerate(st):
        dQ, dr = query_score_grads(model, b["Q"], ent64[b["repl"]], norm)
        np.add.at(g_ent, b["repl"], b["c"][:, None] * dr)  # reduced on owner(repl): local rows, local queries
        gQo, gQs = np.zeros((n, K), gd), np.zeros((n, K), gd)
        own_r = owner(b["repl"])
        for o in range(W):  # one partial row per (positive, side) and owner that holds at least one of its candidates
            for side, acc in ((True, gQo), (False, gQs)):
                m = (own_r == o) & (b["keep"] == side)
                part = np.zeros((n, K), gd)
                np.add.at(part, b["i"][m], b["c"][m, None] * dQ[m])
                acc += part
                if o != w:
                    by["partials_in"][w] += row_b * np.unique(b["i"][m]).size
        # 6. batch owner: chain to the positive's rows, add the positive's own term, hand s / o rows to their owners
        e_s, e_p, e_o = b["e"]
        gs, gp, go = unfold_query_grads(model, k, e_s, e_p, e_o, gQo, gQs)
        ps, pp, po = ko.score_grad_rows(model, k, e_s, e_p, e_o, norm, gd)
        d = b["dpos"][:, None]
        np.add.at(g_ent, b["pos"][:, 0], gs + d * ps)
        np.add.at(g_ent, b["pos"][:, 2], go + d * po)
        np.add.at(g_rel, b["pos"][:, 1], gp + d * pp)
        for col in (0, 2):
            ow = owner(b["pos"][:, col])
            for o in range(W):
                if o != w:
                    by["posgrad_in"][o] += row_b * np.count_nonzero(ow == o)
    return dict(loss=loss_total, grad_ent=g_ent, grad_rel=g_rel, bytes={k_: int(v.max()) for k_, v in by.items()},
                bytes_total=int(sum(v.max() for v in by.values())))


def push_step_bytes(eta, n, K, W, E, batches):
    """Bytes received per rank (worst rank) by the CURRENT exchange: rows of every entity slot
    of the batch whose owner is remote + the all-gathered [Qo | Qs | coef | keep] tails."""
    own = lambda ids: np.asarray(ids) // ((E + W - 1) // W)  # noqa: E731
    worst = 0
    for w, (pos, keep, repl) in enumerate(batches):
        pos = np.asarray(pos).reshape(-1, 3)
        slots = np.concatenate([pos[:, 0], pos[:, 2], np.asarray(repl)])
        worst = max(worst, 4 * K * int(np.count_nonzero(own(slots) != w)))
    tail = (2 * n * K + eta * n) * 4 + eta * n
    return dict(rows_in=worst, tails_in=tail * (W - 1), total=worst + tail * (W - 1))
