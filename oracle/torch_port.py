"""Reference-equivalent CPU arm: the op graph the reference executes per batch / per test triple,
restated op for op on torch-CPU (multi-threaded), for ``bench.py``'s ``cpu_baseline`` and
``--impl reference`` legs.

TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by ``emgraph_b200``.

The reference (bi-graph/Emgraph) is Python over TensorFlow 2.2, which is not installable in this
image, and /root/reference does not exist on the GPU box, so the CPU baseline is a *port*
(``cpu_baseline.kind == "port"``).  It deliberately keeps the reference's structure:

* training (models/EmbeddingModel.py:614-822, :1415-1418): three materialised ``[n,K]`` gathers
  for the positives and three ``[n*eta,K]`` gathers for the corruptions
  (``_lookup_embeddings`` :490-533), un-fused elementwise/reduce chains (``_fn``), the loss op
  chain (losses/*.py), reverse-mode autodiff through all of it (tf.GradientTape -> torch
  autograd), duplicate-summed row gradients, and the Keras (non-lazy) Adam: a freshly constructed
  optimizer per batch (training/adam.py:45-46) whose sparse apply runs dense m/v decay and a dense
  var update over the whole [E,K] table.
* ranking (models/EmbeddingModel.py:1845-1986): one test triple at a time, all 2E corruptions
  materialised and scored, filter sets looked up per triple (dict lookups stand in for the two
  SQLite queries of datasets/sqlite_adapter.py:472-489 -- cheaper than the reference's), x1e5 int
  truncation and the comparison sums of perform_comparision (:1989-2033).

Validated against oracle/kge_oracle.py (and through it against the reference's own code) in
tests/test_oracle_golden.py::test_torch_port_matches_oracle.
"""
from __future__ import annotations

import numpy as np
import torch


def _fn(model, k, e_s, e_p, e_o, norm=1):
    if model == "TransE":
        return torch.neg(torch.linalg.vector_norm(e_s + e_p - e_o, ord=norm, dim=1))
    if model == "DistMult":
        return torch.sum(e_s * e_p * e_o, dim=1)
    s_r, s_i = torch.chunk(e_s, 2, dim=1)
    p_r, p_i = torch.chunk(e_p, 2, dim=1)
    o_r, o_i = torch.chunk(e_o, 2, dim=1)
    f = (torch.sum(p_r * s_r * o_r, dim=1) + torch.sum(p_r * s_i * o_i, dim=1)
         + torch.sum(p_i * s_r * o_i, dim=1) - torch.sum(p_i * s_i * o_r, dim=1))
    return (2.0 / k) * f if model == "HolE" else f


def _lookup(ent, rel, x):
    return ent[x[:, 0]], rel[x[:, 1]], ent[x[:, 2]]


def _loss(loss, scores_pos, scores_neg, eta, margin):
    if loss == "pairwise":
        sp = scores_pos.repeat(eta)
        return torch.sum(torch.maximum(margin - sp + scores_neg, torch.tensor(0.0)))
    if loss == "nll":
        sp = torch.clamp(scores_pos.repeat(eta), -75.0, 75.0)
        sn = torch.clamp(scores_neg, -75.0, 75.0)
        return torch.sum(torch.log(1 + torch.exp(torch.cat([-sp, sn], 0))))
    sp = torch.clamp(scores_pos, -75.0, 75.0)
    sn = torch.clamp(scores_neg, -75.0, 75.0).reshape(eta, scores_pos.shape[0])
    neg_exp, pos_exp = torch.exp(sn), torch.exp(sp)
    return -torch.sum(torch.log(pos_exp / (torch.sum(neg_exp, dim=0) + pos_exp)))


class CpuTrainer:
    """Holds ent/rel as leaf tensors; ``step`` == one ``optimizer.minimize`` of the reference."""

    def __init__(self, model, k, loss, eta, ent, rel, margin=1.0, norm=1, lr=5e-4, optimizer="adam"):
        self.model, self.k, self.loss, self.eta = model, k, loss, eta
        self.margin, self.norm, self.lr, self.optimizer = margin, norm, lr, optimizer
        self.ent = torch.tensor(np.asarray(ent), dtype=torch.float32, requires_grad=True)
        self.rel = torch.tensor(np.asarray(rel), dtype=torch.float32, requires_grad=True)

    def corruptions(self, x_pos, keep_subj, repl):
        n = x_pos.shape[0]
        ds = x_pos.reshape(-1).repeat(self.eta).reshape(n * self.eta, 3)
        ks = keep_subj.to(torch.int64)
        ko = 1 - ks
        subj = ks * ds[:, 0] + ko * repl
        obj = ko * ds[:, 2] + ks * repl
        return torch.stack([subj, ds[:, 1], obj], dim=1)

    def loss_value(self, x_pos, keep_subj, repl):
        e_s, e_p, e_o = _lookup(self.ent, self.rel, x_pos)
        scores_pos = _fn(self.model, self.k, e_s, e_p, e_o, self.norm)
        x_neg = self.corruptions(x_pos, keep_subj, repl)
        e_s, e_p, e_o = _lookup(self.ent, self.rel, x_neg)
        scores_neg = _fn(self.model, self.k, e_s, e_p, e_o, self.norm)
        return _loss(self.loss, scores_pos, scores_neg, self.eta, self.margin)

    def step(self, x_pos, keep_subj=None, repl=None, rng=None):
        """x_pos int64 [n,3].  Draws the corruptions like generate_corruptions_for_fit when not given."""
        n = x_pos.shape[0]
        E = self.ent.shape[0]
        if repl is None:
            keep_subj = torch.randint(0, 2, (n * self.eta,), generator=rng)
            repl = torch.randint(0, E, (n * self.eta,), generator=rng)
        self.ent.grad = None
        self.rel.grad = None
        loss = self.loss_value(x_pos, keep_subj, repl)
        loss.backward()  # dense [E,K] grad == IndexedSlices with duplicates summed
        with torch.no_grad():
            for w in (self.ent, self.rel):
                g = w.grad
                if self.optimizer == "adam":
                    # fresh Keras Adam, first step, dense apply over the whole table
                    m = (1 - 0.9) * g
                    v = (1 - 0.999) * g * g
                    lr_t = self.lr * np.sqrt(1 - 0.999) / (1 - 0.9)
                    w -= lr_t * m / (torch.sqrt(v) + 1e-7)
                elif self.optimizer == "adagrad":
                    a = 0.1 + g * g
                    w -= self.lr * g / (torch.sqrt(a) + 1e-7)
                else:
                    w -= self.lr * g
        return float(loss.detach())


class CpuRanker:
    """Per-test-triple all-entity ranking, the way the reference's eval graph does it."""

    def __init__(self, model, k, ent, rel, filter_triples=None, norm=1):
        self.model, self.k, self.norm = model, k, norm
        self.ent = torch.as_tensor(np.asarray(ent), dtype=torch.float32)
        self.rel = torch.as_tensor(np.asarray(rel), dtype=torch.float32)
        self.E = self.ent.shape[0]
        self.all_ent = torch.arange(self.E, dtype=torch.int64)
        self.sp, self.po = None, None
        if filter_triples is not None:
            self.sp, self.po = {}, {}
            for s, p, o in np.asarray(filter_triples).reshape(-1, 3).tolist():
                self.sp.setdefault((s, p), set()).add(o)
                self.po.setdefault((p, o), set()).add(s)

    @staticmethod
    def _cmp(sc, sp):
        return int(torch.sum((sc * 1e5).to(torch.int32) >= (sp * 1e5).to(torch.int32)))

    def rank(self, x):
        s, p, o = (int(v) for v in x)
        E = self.E
        xs = torch.full((E,), s, dtype=torch.int64)
        xp = torch.full((E,), p, dtype=torch.int64)
        xo = torch.full((E,), o, dtype=torch.int64)
        corr = torch.cat([torch.stack([xs, xp, self.all_ent], 1), torch.stack([self.all_ent, xp, xo], 1)], 0)
        e_s, e_p, e_o = _lookup(self.ent, self.rel, corr)
        sc = _fn(self.model, self.k, e_s, e_p, e_o, self.norm)
        xt = torch.tensor([[s, p, o]], dtype=torch.int64)
        e_s, e_p, e_o = _lookup(self.ent, self.rel, xt)
        sp = _fn(self.model, self.k, e_s, e_p, e_o, self.norm).squeeze()
        obj_sc, sub_sc = sc[:E], sc[E:]
        hi_o = hi_s = 0
        if self.sp is not None:
            idx_o = torch.tensor(sorted(self.sp.get((s, p), set()) | {o}), dtype=torch.int64)
            idx_s = torch.tensor(sorted(self.po.get((p, o), set()) | {s}), dtype=torch.int64)
            hi_o = self._cmp(obj_sc[idx_o], sp)
            hi_s = self._cmp(sub_sc[idx_s], sp)
        return [self._cmp(sub_sc, sp) + 1 - hi_s, self._cmp(obj_sc, sp) + 1 - hi_o]

    def ranks(self, test):
        return np.asarray([self.rank(x) for x in np.asarray(test).reshape(-1, 3)])
