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


# ------------------------------------------------------------------------------------------------
# Dimension-sharded step (the exchange the product runs, emgraph_b200/csrc/kge_dim.cuh): rank r holds a COLUMN
# slice of every row and sees the whole global batch; the ranks exchange one partial sum per scored triple.
# Every reference scoring function is a sum over columns -- models/TransE.py:208-216 (norm 2: its square),
# DistMult.py:201, ComplEx.py:288-298, HolE.py:189 -- so score = finish(sum_r raw_r) and
# d score / d(local columns) = finish'(total) * d raw_r / d(local columns).
# ------------------------------------------------------------------------------------------------
def raw_partial(model, kc, e_s, e_p, e_o, norm=1, dtype=np.float64):
    """Raw column-slice sum of one rank: sum|u| (TransE-1), sum u^2 (TransE-2), the trilinear form without HolE's 2/k."""
    e_s, e_p, e_o = (np.asarray(x, dtype=dtype) for x in (e_s, e_p, e_o))
    if model == "TransE":
        u = e_s + e_p - e_o
        return np.sum(np.abs(u), axis=1, dtype=dtype) if norm == 1 else np.sum(u * u, axis=1, dtype=dtype)
    if model == "DistMult":
        return np.sum(e_s * e_p * e_o, axis=1, dtype=dtype)
    return ko.score_rows("ComplEx", kc, e_s, e_p, e_o, 1, dtype)


def finish_total(model, k, total, norm=1):
    """(score, d score / d total) from the all-reduced raw sum; k = the WHOLE model's k (HolE's scale)."""
    total = np.asarray(total)
    if model == "TransE":
        if norm == 1:
            return -total, -np.ones_like(total)
        nrm = np.sqrt(total)
        with np.errstate(divide="ignore"):
            return -nrm, np.where(nrm != 0, -0.5 / nrm, 0.0)
    c = 2.0 / k if model == "HolE" else 1.0
    return c * total, np.full_like(total, c)


def raw_partial_grads(model, kc, e_s, e_p, e_o, norm=1, dtype=np.float64):
    """d raw_partial / d(e_s, e_p, e_o) on one rank's slice."""
    e_s, e_p, e_o = (np.asarray(x, dtype=dtype) for x in (e_s, e_p, e_o))
    if model == "TransE":
        u = e_s + e_p - e_o
        g = np.sign(u) if norm == 1 else 2.0 * u
        return g, g, -g
    return ko.score_grad_rows("DistMult" if model == "DistMult" else "ComplEx", kc, e_s, e_p, e_o, 1, dtype)


def dim_slice(full, model, k, W, r):
    """Rank r's column slice (emgraph_b200/distributed.py:slice_columns restated: ceil(k/W) rounded up to 4 columns per
    half, zero columns past the end of the model)."""
    kc = (-(-k // W) + 3) // 4 * 4
    c0, c1 = min(k, r * kc), min(k, (r + 1) * kc)
    halves = 2 if model in ("ComplEx", "HolE") else 1
    out = np.zeros((np.asarray(full).shape[0], halves * kc), np.asarray(full).dtype)
    for h in range(halves):
        out[:, h * kc:h * kc + (c1 - c0)] = np.asarray(full)[:, h * k + c0:h * k + c1]
    return out, kc, (c0, c1)


def dim_sharded_step(model, k, loss, eta, ent, rel, pos, keep_subj, repl, W, margin=1.0, norm=1, alpha=0.5, nl="linear",
                     dtype=np.float64):
    """One step of the global batch on W column-sharded ranks.  Returns dict(loss, scores_pos, scores_neg, grad_ent,
    grad_rel (merged to the model's column order), bytes_per_rank = all-reduce payload)."""
    pos = np.asarray(pos).reshape(-1, 3)
    neg = ko.corruptions_for_fit(pos, eta, keep_subj, repl)
    halves = 2 if model in ("ComplEx", "HolE") else 1
    sl = []
    tot_p = np.zeros(pos.shape[0], dtype)
    tot_n = np.zeros(neg.shape[0], dtype)
    for r in range(W):  # phase 1 on every rank, then the all-reduce (sum in rank order)
        e, kc, cr = dim_slice(ent, model, k, W, r)
        rl, _, _ = dim_slice(rel, model, k, W, r)
        sl.append((e, rl, kc, cr))
        tot_p = tot_p + raw_partial(model, kc, e[pos[:, 0]], rl[pos[:, 1]], e[pos[:, 2]], norm, dtype)
        tot_n = tot_n + raw_partial(model, kc, e[neg[:, 0]], rl[neg[:, 1]], e[neg[:, 2]], norm, dtype)
    val, sp, sn, wp, wn = dim_loss_weights(model, k, loss, eta, tot_p, tot_n, margin, norm, alpha, nl, dtype)
    g_ent = np.zeros(np.asarray(ent).shape, dtype)
    g_rel = np.zeros(np.asarray(rel).shape, dtype)
    for r, (e, rl, kc, (c0, c1)) in enumerate(sl):  # phase 2: gradient rows of the slice, no exchange
        ge, gr = dim_slice_grads(model, kc, e, rl, pos, neg, wp, wn, norm, dtype)
        for h in range(halves):
            g_ent[:, h * k + c0:h * k + c1] = ge[:, h * kc:h * kc + (c1 - c0)]
            g_rel[:, h * k + c0:h * k + c1] = gr[:, h * kc:h * kc + (c1 - c0)]
    return dict(loss=val, scores_pos=sp, scores_neg=sn, grad_ent=g_ent, grad_rel=g_rel,
                bytes_per_rank=4 * (1 + eta) * pos.shape[0])


def dim_loss_weights(model, k, loss, eta, tot_p, tot_n, margin=1.0, norm=1, alpha=0.5, nl="linear", dtype=np.float64):
    """From the all-reduced raw sums: (loss, scores_pos, scores_neg, dL/d total_pos, dL/d total_neg).  Evaluated
    identically on every rank (losses/*.py through the oracle)."""
    sp_raw, fp = finish_total(model, k, np.asarray(tot_p, dtype), norm)
    sn_raw, fn = finish_total(model, k, np.asarray(tot_n, dtype), norm)
    sp, gsp = ko.non_linearity(nl, sp_raw, dtype)
    sn, gsn = ko.non_linearity(nl, sn_raw, dtype)
    val, dpos, dneg = ko.loss_and_dscore(loss, sp, sn, eta, margin, dtype, alpha)
    return val, sp, sn, np.asarray(dpos, np.float64) * gsp * fp, np.asarray(dneg, np.float64) * gsn * fn


def dim_slice_grads(model, kc, e, rl, pos, neg, wp, wn, norm=1, dtype=np.float64):
    """Summed gradient rows of ONE rank's slice ([E,Kc], [R,Kc]) given dL/d total of every scored triple."""
    ge, gr = np.zeros(e.shape, dtype), np.zeros(rl.shape, dtype)
    for trip, w in ((pos, wp), (neg, wn)):
        gs, gp, go = raw_partial_grads(model, kc, e[trip[:, 0]], rl[trip[:, 1]], e[trip[:, 2]], norm, dtype)
        np.add.at(ge, trip[:, 0], w[:, None] * gs)
        np.add.at(ge, trip[:, 2], w[:, None] * go)
        np.add.at(gr, trip[:, 1], w[:, None] * gp)
    return ge, gr
