"""TEST-ONLY stand-in for emgraph_b200.engine.Engine: the same method surface, CPU tensors, every step computed by
the oracle.  It exists so that the HOST logic of fit() / predict() -- batch slicing, step and seed counters, flags,
side stacking, loss bookkeeping, re-fit seeding -- runs end to end in the CPU test tier (`-m "not gpu"`).  It is
installed by monkeypatching inside tests/test_fit_host_logic.py only; the product never sees it and still raises
without the CUDA library (tests/test_abi_and_host.py::test_no_cpu_fallback)."""
from types import SimpleNamespace

import numpy as np
import torch

from emgraph_b200 import _lib
from oracle import kge_oracle as ko
from oracle import sharded_oracle as so

_MODEL = {0: ("TransE", 1), 1: ("TransE", 2), 2: ("DistMult", 1), 3: ("ComplEx", 1), 4: ("HolE", 1)}
_LOSS = {v: k for k, v in _lib.LOSS_IDS.items()}
_OPT = {v: k for k, v in _lib.OPT_IDS.items()}
_SIDE = {0: "s,o", 1: "s", 2: "o"}
_NL = {v: k for k, v in _lib.NL_IDS.items()}


class FakeEngine:
    tdev = torch.device("cpu")

    def __init__(self):
        self.launches = 0
        self.calls = []  # one record per step: what the host asked for

    # ---- argument block: keep what the host passed
    def train_args(self, **kw):
        kw.setdefault("side", 0)
        kw.setdefault("flags", 0)
        kw.setdefault("step", 1)
        kw.setdefault("seed", 0)
        return SimpleNamespace(**kw)

    def train_args_update(self, a, *, pos, step, lr, loss_out, flags):
        a.pos, a.step, a.lr, a.loss_out, a.flags = pos, step, lr, loss_out, flags
        return a

    def _state(self, a, which):
        m, v = getattr(a, which + "_m", None), getattr(a, which + "_v", None)
        opt = _OPT[a.opt]
        if (a.flags & _lib.F_RESET_STATE) or opt == "sgd":
            return None, ()
        if opt == "adam":
            return (m.numpy().copy(), v.numpy().copy()), (m, v)
        return (m.numpy().copy(),), (m,)

    def _step(self, a, pos):
        model, norm = _MODEL[a.model]
        n, eta, E = pos.shape[0], a.eta, a.ent.shape[0]
        keep_codes = None if getattr(a, "keep_subj", None) is None else a.keep_subj.numpy()
        if getattr(a, "repl", None) is not None:
            repl = a.repl.numpy()
            keep = keep_codes if keep_codes is not None else np.full(n * eta, 1 if a.side == 2 else 0, np.uint8)
        else:
            neg_list = getattr(a, "neg_entities", None)
            repl, keep = ko.draw_corruptions(a.seed, a.step, n, eta, E, _SIDE[a.side], neg_index_base=getattr(a, "neg_index_base", 0),
                                             neg_entities=None if neg_list is None else neg_list.numpy(),
                                             neg_entities_n=getattr(a, "neg_entities_n", 0), keep_codes=keep_codes)
        st_e, t_e = self._state(a, "ent")
        st_r, t_r = self._state(a, "rel")
        state = None if st_e is None else (st_e, st_r)
        o = ko.train_step(model, a.k, _LOSS[a.loss], eta, a.ent.numpy(), a.rel.numpy(), pos, keep, repl, margin=getattr(a, "margin", 1.0),
                          norm=norm, opt=_OPT[a.opt], lr=a.lr, state=state, step=a.step, alpha=getattr(a, "alpha", 0.5),
                          reg_p=getattr(a, "reg_p", 0), reg_lambda_ent=getattr(a, "reg_lambda_ent", 0.0),
                          reg_lambda_rel=getattr(a, "reg_lambda_rel", 0.0), nl=_NL[getattr(a, "non_linearity", 0)])
        if not (a.flags & _lib.F_NO_UPDATE):
            a.ent.copy_(torch.from_numpy(o["ent_new"]))
            a.rel.copy_(torch.from_numpy(o["rel_new"]))
            for tens, new in ((t_e, o["state_ent"]), (t_r, o["state_rel"])):
                for t, x in zip(tens, new):
                    t.copy_(torch.from_numpy(x))
        a.loss_out[0] = float(o["loss"])
        self.calls.append(dict(step=a.step, seed=a.seed, n=n, flags=a.flags, side=a.side, lr=a.lr, keep_codes=keep_codes is not None,
                               neg_entities=getattr(a, "neg_entities", None) is not None, pos=pos.copy()))
        self.launches += 1

    def train_step(self, a):
        self._step(a, a.pos.numpy())

    def train_step_host(self, a, pos_host, loss_host):
        self._step(a, pos_host.numpy())
        loss_host[0] = a.loss_out[0]

    def train_step_host_async(self, a, pos_host, loss_slot):
        self._step(a, pos_host.numpy())
        loss_slot[0] = a.loss_out[0]
        return len(self.calls) % 4

    def train_host_wait(self, ticket):
        assert 0 <= ticket < 4

    # ---- dimension-sharded step (include/kge_b200.h: kge_train_partial / _backward / _reduce) on one rank's column slice
    def _dim_corruptions(self, a):
        n, eta = a.pos.shape[0], a.eta
        keep_codes = None if getattr(a, "keep_subj", None) is None else a.keep_subj.numpy()
        if getattr(a, "repl", None) is not None:
            return a.repl.numpy(), (keep_codes if keep_codes is not None else np.full(n * eta, 1 if a.side == 2 else 0, np.uint8))
        return ko.draw_corruptions(a.seed, a.step, n, eta, a.ent.shape[0], _SIDE[a.side], keep_codes=keep_codes)

    def train_partial(self, a, sums, i_begin=0, i_end=None):
        model, norm = _MODEL[a.model]
        pos = a.pos.numpy()
        n, eta = pos.shape[0], a.eta
        i_end = n if i_end is None else i_end
        if i_begin == 0:
            repl, keep = self._dim_corruptions(a)
            self._dim = dict(neg=ko.corruptions_for_fit(pos, eta, keep, repl), totals={}, step=a.step)
        e, rl, neg = a.ent.numpy(), a.rel.numpy(), self._dim["neg"]
        nc = i_end - i_begin
        out = np.zeros((1 + eta) * nc, np.float32)
        p = pos[i_begin:i_end]
        out[:nc] = so.raw_partial(model, a.k, e[p[:, 0]], rl[p[:, 1]], e[p[:, 2]], norm, np.float32)
        for j in range(eta):
            t = neg[j * n + i_begin:j * n + i_end]
            out[nc + j * nc:nc + (j + 1) * nc] = so.raw_partial(model, a.k, e[t[:, 0]], rl[t[:, 1]], e[t[:, 2]], norm, np.float32)
        sums[:out.size].copy_(torch.from_numpy(out))
        self.launches += 1

    def train_partial_sorted(self, a, sums, n_chunks=1):
        """kge_train_partial_sorted: every piece of the chunk-major layout, computed piece by piece (the order in which the
        rows are read does not change the sums)."""
        n, eta = a.pos.shape[0], a.eta
        chunks = max(1, min(int(n_chunks), n))
        base, extra = divmod(n, chunks)
        lo = 0
        for c in range(chunks):
            hi = lo + base + (1 if c < extra else 0)
            self.train_partial(a, sums[(1 + eta) * lo:(1 + eta) * hi], lo, hi)
            lo = hi

    def train_backward(self, a, sums, i_begin=0, i_end=None):
        i_end = a.pos.shape[0] if i_end is None else i_end
        self._dim["totals"][(i_begin, i_end)] = sums.numpy()[:(1 + a.eta) * (i_end - i_begin)].copy()
        self.launches += 1

    def train_reduce(self, a):
        model, norm = _MODEL[a.model]
        pos = a.pos.numpy()
        n, eta = pos.shape[0], a.eta
        tot_p, tot_n = np.zeros(n, np.float32), np.zeros(n * eta, np.float32)
        covered = 0
        for (lo, hi), t in sorted(self._dim["totals"].items()):
            nc = hi - lo
            tot_p[lo:hi] = t[:nc]
            for j in range(eta):
                tot_n[j * n + lo:j * n + hi] = t[nc + j * nc:nc + (j + 1) * nc]
            covered += nc
        assert covered == n, "the chunks of a step must cover the global batch"
        k_model = getattr(a, "k_model", 0) or a.k
        val, _, _, wp, wn = so.dim_loss_weights(model, k_model, _LOSS[a.loss], eta, tot_p, tot_n, getattr(a, "margin", 1.0), norm,
                                                getattr(a, "alpha", 0.5), _NL[getattr(a, "non_linearity", 0)], np.float32)
        e, rl, neg = a.ent.numpy(), a.rel.numpy(), self._dim["neg"]
        ge, gr = so.dim_slice_grads(model, a.k, e, rl, pos, neg, wp, wn, norm)
        t_ent = np.zeros(e.shape[0], bool)
        for trip in (pos, neg):
            t_ent[trip[:, 0]] = True
            t_ent[trip[:, 2]] = True
        t_rel = np.zeros(rl.shape[0], bool)
        t_rel[pos[:, 1]] = True
        st_e, ten_e = self._state(a, "ent")
        st_r, ten_r = self._state(a, "rel")
        if not (a.flags & _lib.F_NO_UPDATE):
            e_new, se = ko.optimizer_step(_OPT[a.opt], e, ge, t_ent, a.lr, st_e, a.step)
            r_new, sr = ko.optimizer_step(_OPT[a.opt], rl, gr, t_rel, a.lr, st_r, a.step)
            a.ent.copy_(torch.from_numpy(e_new))
            a.rel.copy_(torch.from_numpy(r_new))
            for tens, new in ((ten_e, se), (ten_r, sr)):
                for t, x in zip(tens, new):
                    t.copy_(torch.from_numpy(x))
        a.loss_out[0] = float(val)
        self.calls.append(dict(step=a.step, seed=a.seed, n=n, flags=a.flags, side=a.side, lr=a.lr, pos=pos.copy(), dim=True))
        self.launches += 1

    def rank_counts_rows(self, model, k, E, rel, s_rows, o_rows, ent_local, test, *, row_begin, row_end, side=0, filtered=False,
                         use_tensor_cores=False, non_linearity=0, counts=None):
        """kge_rank_counts_rows: candidates are the rows of ent_local (global ids row_begin..row_end-1), the test triples'
        own subject / object rows come from the caller."""
        name, norm = _MODEL[model]
        cand = np.arange(row_begin, row_end)
        loc, reln, tst = ent_local.numpy()[:row_end - row_begin], rel.numpy(), test.numpy()
        sr, orr = s_rows.numpy(), o_rows.numpy()
        filt = ko.FilterIndex(self._filter) if filtered else None
        out = np.zeros((tst.shape[0], 2, 4), np.int32)
        nl = _NL[non_linearity]
        for t, x in enumerate(tst):
            p = reln[x[1]][None]
            qp = ko.quantise(ko.non_linearity(nl, ko.score_rows(name, k, sr[t][None], p, orr[t][None], norm))[0])[0]
            idx_o, idx_s = filt.participating(x) if filt is not None else ((), ())
            for sd, (col, known) in enumerate(((2, idx_o), (0, idx_s))):
                if (side == 3 and sd == 1) or (side == 2 and sd == 0):
                    continue
                m = cand != x[col]
                c, rows = cand[m], loc[m]
                if col == 2:
                    sc = ko.score_rows(name, k, np.repeat(sr[t][None], len(c), 0), np.repeat(p, len(c), 0), rows, norm)
                else:
                    sc = ko.score_rows(name, k, rows, np.repeat(p, len(c), 0), np.repeat(orr[t][None], len(c), 0), norm)
                q = ko.quantise(ko.non_linearity(nl, sc)[0])
                f = np.isin(c, np.asarray(list(known), dtype=np.int64))
                out[t, sd] = [(q > qp).sum(), (q == qp).sum(), ((q > qp) & f).sum(), ((q == qp) & f).sum()]
        return torch.from_numpy(out)

    def normalize_rows(self, emb):
        nrm = emb.norm(dim=1, keepdim=True).clamp(min=1.0)
        emb.div_(nrm)

    def score(self, model, k, ent, rel, triples, non_linearity=0):
        name, norm = _MODEL[model]
        return torch.from_numpy(ko.non_linearity(_NL[non_linearity], ko.score(name, k, ent.numpy(), rel.numpy(), triples.numpy(), norm))[0])

    # ---- evaluation: the filter is the raw triple array, ranks come from the oracle
    def filter_build(self, triples, E, R):
        self._filter = triples.numpy().copy()
        self._filter_shape = (E, R)

    def filter_clear(self):
        self._filter = None

    def rank_host(self, model, k, ent, rel, test_host, ranks_host, *, side=0, strategy=0, filtered=False, use_tensor_cores=False,
                  non_linearity=0):
        name, norm = _MODEL[model]
        side_s = {v: s for s, v in _lib.RANK_SIDE_IDS.items()}[side]
        strat_s = {v: s for s, v in _lib.STRATEGY_IDS.items()}[strategy]
        if test_host.shape[0] == 0:
            return
        filt = getattr(self, "_filter", None) if filtered else None
        if filtered:
            assert filt is not None and self._filter_shape == (ent.shape[0], rel.shape[0])
        r = ko.ranks(name, k, ent.numpy(), rel.numpy(), test_host.numpy(), filt, side_s, strat_s, norm, nl=_NL[non_linearity])
        ranks_host.copy_(torch.from_numpy(np.asarray(r, np.int32)).reshape(ranks_host.shape))
        self.rank_calls = getattr(self, "rank_calls", 0) + 1

    def rank_counts(self, model, k, ent, rel, test, *, side=0, filtered=False, use_tensor_cores=False, ent_local=None, row_begin=0,
                    row_end=None, counts=None, non_linearity=0):
        """counts[T,2,4]: {object sweep, subject sweep} x {gt, eq, gt & filtered, eq & filtered} over candidate rows
        [row_begin,row_end), the test triple's own entity skipped (include/kge_b200.h:kge_rank_counts)."""
        name, norm = _MODEL[model]
        E = ent.shape[0]
        row_end = E if row_end is None else row_end
        cand = np.arange(row_begin, row_end)
        entn, reln, tst = ent.numpy(), rel.numpy(), test.numpy()
        filt = ko.FilterIndex(self._filter) if filtered else None
        out = np.zeros((tst.shape[0], 2, 4), np.int32)
        nl = _NL[non_linearity]
        for t, x in enumerate(tst):
            qp = ko.quantise(ko.non_linearity(nl, ko.score(name, k, entn, reln, x, norm))[0])[0]
            idx_o, idx_s = filt.participating(x) if filt is not None else ((), ())
            for sd, (col, known) in enumerate(((2, idx_o), (0, idx_s))):
                if (side == 3 and sd == 1) or (side == 2 and sd == 0):  # KGE_RANK_O / KGE_RANK_S sweep one side
                    continue
                c = cand[cand != x[col]]
                tri = np.tile(x, (c.shape[0], 1))
                tri[:, col] = c
                q = ko.quantise(ko.non_linearity(nl, ko.score(name, k, entn, reln, tri, norm))[0])
                f = np.isin(c, np.asarray(list(known), dtype=np.int64))
                out[t, sd] = [(q > qp).sum(), (q == qp).sum(), ((q > qp) & f).sum(), ((q == qp) & f).sum()]
        return torch.from_numpy(out)

    def rank_finalize(self, counts, *, side=0, strategy=0, filtered=False, self_is_candidate=None):
        c = counts.numpy().astype(np.int64)
        T = c.shape[0]
        sc = np.ones((T, 2), np.int64) if self_is_candidate is None else (self_is_candidate.numpy() != 0).astype(np.int64)

        def cmpc(gt, eq):
            return gt if strategy == 1 else (gt + (eq + 1) // 2 if strategy == 2 else gt + eq)

        go, eo, gs, es = c[:, 0, 0], c[:, 0, 1] + sc[:, 1], c[:, 1, 0], c[:, 1, 1] + sc[:, 0]
        fo = cmpc(c[:, 0, 2], c[:, 0, 3] + sc[:, 1]) if filtered else 0
        fs = cmpc(c[:, 1, 2], c[:, 1, 3] + sc[:, 0]) if filtered else 0
        if side == 0:
            r = np.stack([cmpc(gs, es) + 1 - fs, cmpc(go, eo) + 1 - fo], 1)
        elif side == 1:
            r = cmpc(go + gs, eo + es) + 1 - fs - fo
        elif side == 2:
            r = cmpc(gs, es) + 1 - fs
        else:
            r = cmpc(go, eo) + 1 - fo
        return torch.from_numpy(np.asarray(r, np.int32))
