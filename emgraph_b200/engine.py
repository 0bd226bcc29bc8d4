"""Thin host driver over the C ABI: owns the kge_ctx, turns torch CUDA tensors into the plain
pointers/sizes of include/kge_b200.h, and enqueues on torch's current stream.  No arithmetic
happens here; without the CUDA library every entry point raises (no CPU fallback)."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import KgeTable, KgeTrainArgs, check


def model_id(name: str, norm: int = 1) -> int:
    if name == "TransE":
        if norm not in (1, 2):
            raise ValueError("TransE norm must be 1 or 2 on the CUDA path, got %r" % (norm,))
        return _lib.MODEL_IDS["TransE"] if norm == 1 else _lib.MODEL_IDS["TransE_L2"]
    return _lib.MODEL_IDS[name]


def internal_k(name: str, k: int) -> int:
    return 2 * k if name in ("ComplEx", "HolE") else k


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk_f32(t, name):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise ValueError("%s must be a contiguous float32 CUDA tensor" % name)


def _chk_i32(t, name):
    if not (t.is_cuda and t.dtype == torch.int32 and t.is_contiguous()):
        raise ValueError("%s must be a contiguous int32 CUDA tensor" % name)


def make_table(shards, rows=None, rows_per_shard=None, K=None) -> KgeTable:
    """shards: one [rows,K] fp32 CUDA tensor, or a list of raw device pointers / tensors (one per
    rank, peer-mapped) for a row-range-sharded table."""
    tb = KgeTable()
    if isinstance(shards, torch.Tensor):
        _chk_f32(shards, "table")
        tb.shard[0] = shards.data_ptr()
        tb.rows = shards.shape[0]
        tb.rows_per_shard = shards.shape[0]
        tb.n_shards = 1
        tb.K = shards.shape[1]
        return tb
    if shards is None:
        tb.n_shards = 1
        tb.rows = rows or 0
        tb.rows_per_shard = rows_per_shard or (rows or 0)
        tb.K = K or 0
        return tb
    assert len(shards) <= _lib.KGE_MAX_SHARDS
    for i, s in enumerate(shards):
        tb.shard[i] = s.data_ptr() if isinstance(s, torch.Tensor) else int(s)
    tb.rows = rows
    tb.rows_per_shard = rows_per_shard
    tb.n_shards = len(shards)
    tb.K = K
    return tb


class Engine:
    """One per process/GPU.  Wraps kge_ctx."""

    def __init__(self, device: int | None = None):
        if not torch.cuda.is_available():
            raise _lib.KgeError("emgraph_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        torch.cuda.set_device(self.device)
        h = C.c_void_p()
        check(self.lib.kge_ctx_create(self.device, C.byref(h)))
        self._h = h
        self.launches = 0  # kernels of ours launched through this engine (bench bookkeeping)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.kge_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def tdev(self):
        return torch.device("cuda", self.device)

    def has_tensor_core_rank(self) -> bool:
        return bool(self.lib.kge_has_tensor_core_rank())

    def train_grad_floats(self, eta: int, n_pos: int, K: int) -> int:
        """floats of the caller-owned gradient buffer for the phased (multi-GPU) calls"""
        return int(self.lib.kge_train_grad_floats(eta, n_pos, K))

    def set_timing(self, on: bool):
        check(self.lib.kge_ctx_set_timing(self._h, int(bool(on))))

    def get_timing(self):
        """average ms per step of the train-step phases since set_timing(True)"""
        out = (C.c_float * 5)()
        n = C.c_int()
        check(self.lib.kge_ctx_get_timing(self._h, out, C.byref(n)))
        return dict(zip(("emit", "fwd_bwd", "reduce_apply", "spans", "sort_after_emit"), [float(x) for x in out])), int(n.value)

    def get_timing_ex(self):
        """like get_timing with the wait for the sort split off the level-1 reduction kernel"""
        out = (C.c_float * 6)()
        n = C.c_int()
        check(self.lib.kge_ctx_get_timing_ex(self._h, out, 6, C.byref(n)))
        return dict(zip(("emit", "fwd_bwd", "sort_wait", "reduce_apply", "spans", "sort_after_emit"), [float(x) for x in out])), int(n.value)

    def sort_entries(self, entries, n_keys: int, algo: int = 0):
        """test hook: entries = int64 CUDA tensor of (key << 32 | slot); returns them sorted by key, equal keys in input order"""
        assert entries.is_cuda and entries.dtype == torch.int64 and entries.is_contiguous()
        out = torch.empty_like(entries)
        check(self.lib.kge_sort_entries(self._h, entries.data_ptr(), entries.numel(), int(n_keys), int(algo), out.data_ptr(), _stream()))
        self.launches += 3
        return out

    def workspace_bytes(self) -> int:
        return int(self.lib.kge_ctx_workspace_bytes(self._h))

    # ---------------------------------------------------------------- scoring
    def score(self, model: int, k: int, ent, rel, triples, non_linearity: int = 0):
        """ent: tensor [E,K] or KgeTable; returns fp32 CUDA tensor [n] (nl(score) when non_linearity != 0)."""
        tb = ent if isinstance(ent, KgeTable) else make_table(ent)
        _chk_f32(rel, "rel")
        _chk_i32(triples, "triples")
        n = triples.shape[0]
        out = torch.empty(n, dtype=torch.float32, device=self.tdev)
        check(self.lib.kge_predict(self._h, model, k, C.byref(tb), _ptr(rel), rel.shape[0], _ptr(triples), n, int(non_linearity),
                                   _ptr(out), _stream()))
        self.launches += 1 if n else 0
        return out

    # ---------------------------------------------------------------- training
    def train_args(self, *, model, loss, opt, k, eta, ent, rel, pos, loss_out, side=0, flags=0, margin=1.0,
                   lr=5e-4, beta1=0.9, beta2=0.999, eps=1e-7, momentum=0.9, seed=0, step=1, neg_index_base=0,
                   ent_m=None, ent_v=None, rel_m=None, rel_v=None, repl=None, keep_subj=None,
                   dbg_scores=None, dbg_grad_ent=None, dbg_grad_rel=None, stage=None, grad_tails=None,
                   grad_tail_stride=0, alpha=0.5, reg_p=0, reg_lambda_ent=0.0, reg_lambda_rel=0.0,
                   neg_entities=None, neg_entities_n=0, non_linearity=0, n_pos=None, k_model=0) -> KgeTrainArgs:
        """n_pos: batch size when `pos` is None (host-buffer step: the positives arrive with the call)."""
        a = KgeTrainArgs()
        a.model, a.loss, a.opt, a.side, a.flags = model, loss, opt, side, flags
        a.k, a.eta, a.margin, a.alpha = k, eta, margin, alpha
        a.reg_p, a.reg_lambda_ent, a.reg_lambda_rel = int(reg_p), float(reg_lambda_ent), float(reg_lambda_rel)
        a.non_linearity = int(non_linearity)
        a.k_model = int(k_model)
        a.lr, a.beta1, a.beta2, a.eps, a.momentum = lr, beta1, beta2, eps, momentum
        a.seed, a.step, a.neg_index_base = seed, step, neg_index_base
        a.ent = ent if isinstance(ent, KgeTable) else make_table(ent)
        K = a.ent.K
        for name, t in (("ent_m", ent_m), ("ent_v", ent_v)):
            if t is None:
                tb = make_table(None, rows=a.ent.rows, rows_per_shard=a.ent.rows_per_shard, K=K)
                tb.n_shards = a.ent.n_shards
            else:
                tb = t if isinstance(t, KgeTable) else make_table(t)
            setattr(a, name, tb)
        _chk_f32(rel, "rel")
        a.rel, a.rel_m, a.rel_v, a.R = rel.data_ptr(), (rel_m.data_ptr() if rel_m is not None else None), \
            (rel_v.data_ptr() if rel_v is not None else None), rel.shape[0]
        if pos is not None:
            _chk_i32(pos, "pos")
            a.pos, a.n_pos = pos.data_ptr(), pos.shape[0]
        if repl is not None:
            _chk_i32(repl, "repl")
            assert repl.numel() == eta * pos.shape[0]
            a.repl = repl.data_ptr()
        if keep_subj is not None:
            n_chk = pos.shape[0] if pos is not None else n_pos
            assert n_chk is not None, "keep_subj without pos needs n_pos"
            assert keep_subj.is_cuda and keep_subj.dtype == torch.uint8 and keep_subj.numel() == eta * n_chk
            a.keep_subj = keep_subj.data_ptr()
        a.loss_out = loss_out.data_ptr() if loss_out is not None else None
        a.dbg_scores = dbg_scores.data_ptr() if dbg_scores is not None else None
        a.dbg_grad_ent = dbg_grad_ent.data_ptr() if dbg_grad_ent is not None else None
        a.dbg_grad_rel = dbg_grad_rel.data_ptr() if dbg_grad_rel is not None else None
        if stage is not None:
            _chk_f32(stage, "stage")
            assert stage.numel() >= (2 + eta) * a.n_pos * K
            a.stage = stage.data_ptr()
        if grad_tails is not None:
            _chk_f32(grad_tails, "grad_tails")
            a.grad_tails, a.grad_tail_stride = grad_tails.data_ptr(), int(grad_tail_stride)
        if neg_entities is not None:
            _chk_i32(neg_entities, "neg_entities")
            a.neg_entities, a.neg_entities_n = neg_entities.data_ptr(), neg_entities.numel()
        else:
            a.neg_entities_n = int(neg_entities_n)
        # keep python references alive for the duration of the call
        a._keep = (ent, rel, pos, loss_out, ent_m, ent_v, rel_m, rel_v, repl, keep_subj, dbg_scores, dbg_grad_ent, dbg_grad_rel,
                   stage, grad_tails, neg_entities)
        return a

    def train_args_update(self, a: KgeTrainArgs, *, pos, step, lr, loss_out, flags) -> KgeTrainArgs:
        """Refresh the per-step fields of an argument block built by train_args (everything else of a fit() loop --
        tables, state, hyper-parameters -- stays what it was): the block is then byte-identical to a freshly built one
        (tests/test_abi_and_host.py::test_cached_argument_block_equals_a_fresh_one) at a fraction of the host cost."""
        if pos is not None:  # host-buffer steps pass pos=None: the batch arrives with the call, which sets n_pos
            _chk_i32(pos, "pos")
            a.pos, a.n_pos = pos.data_ptr(), pos.shape[0]
        a.step, a.lr, a.flags = step, lr, flags
        a.loss_out = loss_out.data_ptr()
        a._keep_step = (pos, loss_out)  # keep this step's tensors alive for the duration of the call
        return a

    # kernels per step: emit, fwd_bwd, loss-reduce, radix sort (CUB onesweep: histogram + exclusive
    # sum + ceil(bits/8) passes), reduce_apply, span_apply
    @staticmethod
    def launches_per_step(E_plus_R: int) -> int:
        bits = max(1, math.ceil(math.log2(max(2, E_plus_R))))
        return 3 + 2 + math.ceil(bits / 8) + 2

    def train_step(self, a: KgeTrainArgs):
        check(self.lib.kge_train_step(self._h, C.byref(a), _stream()))
        self.launches += self.launches_per_step(a.ent.rows + a.R)

    def train_step_host(self, a: KgeTrainArgs, pos_host, loss_host):
        """Host-buffer step: pos_host int32 [n,3] CPU tensor (pinned for an async copy), loss_host
        float32 [1] CPU tensor (pinned).  Copies in, runs the step, copies the loss out, syncs."""
        assert (not pos_host.is_cuda) and pos_host.dtype == torch.int32 and pos_host.is_contiguous()
        assert (not loss_host.is_cuda) and loss_host.dtype == torch.float32
        a.n_pos = pos_host.shape[0]
        check(self.lib.kge_train_step_host(self._h, C.byref(a), C.c_void_p(pos_host.data_ptr()),
                                           C.c_void_p(loss_host.data_ptr()), _stream()))
        self.launches += self.launches_per_step(a.ent.rows + a.R)

    def train_step_host_async(self, a: KgeTrainArgs, pos_host, loss_host) -> int:
        """Enqueue [H2D batch, step, D2H loss] without waiting; returns the ticket for train_host_wait.
        pos_host / loss_host must stay alive and untouched until the ticket has been waited for."""
        assert (not pos_host.is_cuda) and pos_host.dtype == torch.int32 and pos_host.is_contiguous()
        assert (not loss_host.is_cuda) and loss_host.dtype == torch.float32
        a.n_pos = pos_host.shape[0]
        t = C.c_int(0)
        check(self.lib.kge_train_step_host_async(self._h, C.byref(a), C.c_void_p(pos_host.data_ptr()),
                                                 C.c_void_p(loss_host.data_ptr()), _stream(), C.byref(t)))
        self.launches += self.launches_per_step(a.ent.rows + a.R)
        return int(t.value)

    def train_host_wait(self, ticket: int):
        check(self.lib.kge_train_host_wait(self._h, int(ticket)))

    def train_emit(self, a: KgeTrainArgs, keys_out):
        _chk_i32(keys_out, "keys_out")
        check(self.lib.kge_train_emit(self._h, C.byref(a), _ptr(keys_out), _stream()))
        self.launches += 1

    def train_grad_head_floats(self, eta: int, n_pos: int, K: int) -> int:
        return int(self.lib.kge_train_grad_head_floats(eta, n_pos, K))

    def train_push_rows(self, a: KgeTrainArgs, keys_all, stage: KgeTable, row_begin: int, row_end: int):
        """Owner-side push of this rank's rows into every rank's staging buffer (see kge_b200.h)."""
        _chk_i32(keys_all, "keys_all")
        check(self.lib.kge_train_push_rows(self._h, C.byref(a), _ptr(keys_all), keys_all.numel(), C.byref(stage),
                                           row_begin, row_end, _stream()))
        self.launches += 1

    def train_select(self, a: KgeTrainArgs, keys_all, row_begin: int, row_end: int):
        _chk_i32(keys_all, "keys_all")
        check(self.lib.kge_train_select(self._h, C.byref(a), _ptr(keys_all), keys_all.numel(), row_begin, row_end, _stream()))
        self.launches += 4

    def train_fwd_bwd(self, a: KgeTrainArgs, grad_rows):
        check(self.lib.kge_train_fwd_bwd(self._h, C.byref(a), _ptr(grad_rows), _stream()))
        self.launches += 2

    def grad_table(self, bufs, eta: int, n_pos: int, K: int) -> KgeTable:
        """kge_table describing the ranks' gradient buffers for train_apply (one tensor / pointer per rank)."""
        if isinstance(bufs, torch.Tensor):
            bufs = [bufs]
        S = (3 + eta) * n_pos
        return make_table(list(bufs), rows=S * len(bufs), rows_per_shard=S, K=K)

    def train_apply(self, a: KgeTrainArgs, keys_all, grads: KgeTable, row_begin: int, row_end: int):
        _chk_i32(keys_all, "keys_all")
        check(self.lib.kge_train_apply(self._h, C.byref(a), _ptr(keys_all), keys_all.numel(), C.byref(grads),
                                       row_begin, row_end, _stream()))
        self.launches += self.launches_per_step(a.ent.rows + a.R) - 3

    # dimension-sharded multi-GPU step (include/kge_b200.h: kge_train_partial / _backward / _reduce)
    def train_partial(self, a: KgeTrainArgs, sums, i_begin: int = 0, i_end: int | None = None):
        """Column-slice partial sums of positives [i_begin,i_end) and their negatives into `sums`
        ((1+eta)*(i_end-i_begin) floats); i_begin == 0 also draws the step's corruptions and starts the key sort."""
        _chk_f32(sums, "sums")
        i_end = a.n_pos if i_end is None else i_end
        assert sums.numel() >= (1 + a.eta) * (i_end - i_begin)
        check(self.lib.kge_train_partial(self._h, C.byref(a), i_begin, i_end, _ptr(sums), _stream()))
        self.launches += 1 + ((1 + self.launches_per_step(a.ent.rows + a.R) - 5) if i_begin == 0 else 0)

    def train_partial_sorted(self, a: KgeTrainArgs, sums, n_chunks: int = 1):
        """Phase 1 of the whole batch, entity rows streamed in sorted order; `sums` ((1+eta)*n_pos floats) holds the pieces of
        n_chunks consecutive positive ranges back to back (include/kge_b200.h)."""
        _chk_f32(sums, "sums")
        assert sums.numel() >= (1 + a.eta) * a.n_pos
        check(self.lib.kge_train_partial_sorted(self._h, C.byref(a), int(n_chunks), _ptr(sums), _stream()))
        self.launches += 2 + (1 + self.launches_per_step(a.ent.rows + a.R) - 5)

    def train_backward(self, a: KgeTrainArgs, sums, i_begin: int = 0, i_end: int | None = None):
        _chk_f32(sums, "sums")
        i_end = a.n_pos if i_end is None else i_end
        check(self.lib.kge_train_backward(self._h, C.byref(a), i_begin, i_end, _ptr(sums), _stream()))
        self.launches += 1

    def train_reduce(self, a: KgeTrainArgs):
        check(self.lib.kge_train_reduce(self._h, C.byref(a), _stream()))
        self.launches += 3

    def allreduce_p2p(self, sums: KgeTable, totals: KgeTable, flags: KgeTable, rank: int, off: int, length: int, seq: int):
        """Sum floats [off, off+length) of every rank's partial-sum buffer over peer memory into every rank's totals buffer
        (include/kge_b200.h: kge_allreduce_p2p); collective: every rank calls it with the same range and seq."""
        check(self.lib.kge_allreduce_p2p(self._h, C.byref(sums), C.byref(totals), C.byref(flags), int(rank), int(off), int(length),
                                         int(seq) & 0xFFFFFFFF, _stream()))
        self.launches += 1

    def normalize_rows(self, emb):
        _chk_f32(emb, "emb")
        check(self.lib.kge_normalize_rows(self._h, _ptr(emb), emb.shape[0], emb.shape[1], _stream()))
        self.launches += 1

    # ---------------------------------------------------------------- ranking
    def filter_build(self, triples, E: int, R: int):
        _chk_i32(triples, "filter triples")
        check(self.lib.kge_filter_build(self._h, _ptr(triples), triples.shape[0], E, R, _stream()))
        self._filter_keep = triples
        self.launches += 12 if triples.shape[0] else 0

    def filter_clear(self):
        check(self.lib.kge_filter_clear(self._h))

    def filter_size(self) -> int:
        return int(self.lib.kge_filter_size_sync(self._h))

    def rank_counts(self, model: int, k: int, ent, rel, test, *, side=0, filtered=False, use_tensor_cores=False,
                    ent_local=None, row_begin=0, row_end=None, counts=None, non_linearity=0):
        tb = ent if isinstance(ent, KgeTable) else make_table(ent)
        if ent_local is None:
            assert isinstance(ent, torch.Tensor)
            ent_local = ent
        if row_end is None:
            row_end = tb.rows
        _chk_f32(ent_local, "ent_local")
        _chk_f32(rel, "rel")
        _chk_i32(test, "test")
        T = test.shape[0]
        if counts is None:
            counts = torch.empty((T, 2, 4), dtype=torch.int32, device=self.tdev)
        check(self.lib.kge_rank_counts(self._h, model, k, C.byref(tb), _ptr(rel), rel.shape[0], _ptr(ent_local),
                                       row_begin, row_end, _ptr(test), T, side, int(bool(filtered)),
                                       int(bool(use_tensor_cores)), int(non_linearity), _ptr(counts), _stream()))
        self.launches += 3 if T else 0
        return counts

    def rank_counts_rows(self, model: int, k: int, E: int, rel, s_rows, o_rows, ent_local, test, *, row_begin, row_end, side=0,
                         filtered=False, use_tensor_cores=False, non_linearity=0, counts=None):
        """kge_rank_counts for a shard of a table this process cannot address: the test triples' subject / object rows
        come from the caller ([T,K] each)."""
        for t, nm in ((rel, "rel"), (s_rows, "s_rows"), (o_rows, "o_rows"), (ent_local, "ent_local")):
            _chk_f32(t, nm)
        _chk_i32(test, "test")
        T = test.shape[0]
        if counts is None:
            counts = torch.empty((T, 2, 4), dtype=torch.int32, device=self.tdev)
        check(self.lib.kge_rank_counts_rows(self._h, model, k, E, _ptr(rel), rel.shape[0], _ptr(s_rows), _ptr(o_rows), _ptr(ent_local),
                                            row_begin, row_end, _ptr(test), T, side, int(bool(filtered)), int(bool(use_tensor_cores)),
                                            int(non_linearity), _ptr(counts), _stream()))
        self.launches += 3 if T else 0
        return counts

    def rank_finalize(self, counts, *, side=0, strategy=0, filtered=False, self_is_candidate=None):
        """self_is_candidate: optional uint8 [T,2] (col 0 subject, col 1 object), 0 where the test triple's own
        entity was not among the swept candidates (entities_subset ranking)."""
        T = counts.shape[0]
        shape = (T, 2) if side == _lib.RANK_SIDE_IDS["s,o"] else (T,)
        ranks = torch.empty(shape, dtype=torch.int32, device=self.tdev)
        if self_is_candidate is not None:
            assert self_is_candidate.is_cuda and self_is_candidate.dtype == torch.uint8 and self_is_candidate.is_contiguous()
            assert self_is_candidate.numel() == 2 * T
        check(self.lib.kge_rank_finalize(self._h, _ptr(counts), T, side, strategy, int(bool(filtered)),
                                         _ptr(self_is_candidate) if self_is_candidate is not None else None, _ptr(ranks), _stream()))
        self.launches += 1 if T else 0
        return ranks

    def rank(self, model: int, k: int, ent, rel, test, *, side=0, strategy=0, filtered=False, use_tensor_cores=False,
             non_linearity=0):
        counts = self.rank_counts(model, k, ent, rel, test, side=side, filtered=filtered, use_tensor_cores=use_tensor_cores,
                                  non_linearity=non_linearity)
        return self.rank_finalize(counts, side=side, strategy=strategy, filtered=filtered)

    def rank_host(self, model: int, k: int, ent, rel, test_host, ranks_host, *, side=0, strategy=0, filtered=False,
                  use_tensor_cores=False, non_linearity=0):
        """Host-buffer ranking: test_host int32 [T,3] CPU (pinned), ranks_host int32 CPU out; syncs."""
        tb = make_table(ent)
        T = test_host.shape[0]
        assert (not test_host.is_cuda) and test_host.dtype == torch.int32 and test_host.is_contiguous()
        assert (not ranks_host.is_cuda) and ranks_host.dtype == torch.int32 and ranks_host.is_contiguous()
        assert ranks_host.numel() == T * (2 if side == _lib.RANK_SIDE_IDS["s,o"] else 1)
        check(self.lib.kge_rank_host(self._h, model, k, C.byref(tb), _ptr(rel), rel.shape[0],
                                     C.c_void_p(test_host.data_ptr()), T, side, strategy, int(bool(filtered)),
                                     int(bool(use_tensor_cores)), int(non_linearity), C.c_void_p(ranks_host.data_ptr()), _stream()))
        self.launches += 4 if T else 0
        return ranks_host

    # ---------------------------------------------------------------- IPC (multi-GPU peer shards)
    def ipc_export(self, t) -> bytes:
        buf = (C.c_char * 64)()
        check(self.lib.kge_ipc_export(_ptr(t), C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def ipc_export_ptr(self, ptr: int) -> bytes:
        buf = (C.c_char * 64)()
        check(self.lib.kge_ipc_export(C.c_void_p(ptr), C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def ipc_open(self, handle: bytes) -> int:
        buf = (C.c_char * 64).from_buffer_copy(handle)
        out = C.c_void_p()
        check(self.lib.kge_ipc_open(C.cast(buf, C.c_void_p), C.byref(out)))
        return int(out.value)

    def ipc_close(self, ptr: int):
        check(self.lib.kge_ipc_close(C.c_void_p(ptr)))


_ENGINES: dict = {}


def get_engine(device: int | None = None) -> Engine:
    """Process-wide engine per device."""
    if not torch.cuda.is_available():
        raise _lib.KgeError("emgraph_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    d = torch.cuda.current_device() if device is None else int(device)
    e = _ENGINES.get(d)
    if e is None:
        e = _ENGINES[d] = Engine(d)
    return e


def to_dev_i32(x, device):
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.int32)).to(device, non_blocking=False)
