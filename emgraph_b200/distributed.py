"""One-process-per-GPU driver of the row-sharded path.

The entity table (and its optimizer state) is split by contiguous row range over the ranks of one
NVSwitch domain.  ``torch.distributed`` (NCCL) carries the plumbing -- the all-gather of sort keys,
two tiny all-reduces that double as device-side barriers, the all-reduce of rank counts -- while the
data path runs over peer memory: every rank maps its peers' shards and gradient buffers with CUDA IPC
and the kernels read them directly (P2P loads over NVLink inside ``kge_fwd_bwd_kernel`` and
``kge_reduce_apply_kernel``).

Replaces the reference's host-paged "large graph" mode (models/EmbeddingModel.py:645-666,
:1070-1097, :1251-1281), which is single-process and SGD-only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check
from .engine import get_engine, internal_k, make_table, model_id


# ------------------------------------------------------------------------------------------------
# pure host logic (unit-tested on CPU with the gloo backend)
# ------------------------------------------------------------------------------------------------
def rows_per_shard(E: int, world: int) -> int:
    return (E + world - 1) // world


def shard_range(E: int, world: int, rank: int):
    """Contiguous row range [begin, end) owned by `rank`."""
    rps = rows_per_shard(E, world)
    return min(E, rank * rps), min(E, (rank + 1) * rps)


def owner_of(rows, E: int, world: int):
    return np.asarray(rows) // rows_per_shard(E, world)


def batch_slice(n_total: int, world: int, rank: int, step: int, n_per_rank: int):
    """Positives [lo, hi) of the training set that `rank` takes at `step` (sequential, unshuffled
    batches like datasets/numpy_adapter.py:105-111, dealt round-robin over ranks)."""
    n_batches = max(1, n_total // n_per_rank)
    b = (step * world + rank) % n_batches
    return b * n_per_rank, (b + 1) * n_per_rank


def neg_index_base(rank: int, eta: int, n_per_rank: int) -> int:
    """First global negative index of a rank: keeps the per-rank Philox streams disjoint."""
    return rank * eta * n_per_rank


def owned_mask(keys, E: int, world: int, rank: int):
    """Slots a rank reduces: entity keys of its row range + every relation key (relations are
    replicated and updated identically everywhere)."""
    keys = np.asarray(keys)
    b, e = shard_range(E, world, rank)
    return (keys >= E) | ((keys >= b) & (keys < e))


def push_block_order(S: int, world: int, own: int, group: int = 0):
    """Host mirror of kge_push_rows_kernel's schedule (csrc/kge_train.cu): the order in which owner `own`
    visits the 32-slot blocks of every rank's S slots, as (destination rank, block index) pairs.  `group`
    consecutive blocks go to one destination before the deal moves to the next; 0 = a whole rank's blocks, so
    each owner works through one destination region at a time, starting with the rank behind it (no two owners
    begin on the same destination).  Used by the host tests to pin that every block is visited exactly once."""
    bpr = (S + 31) // 32
    if group <= 0 or group > bpr:
        group = bpr
    gpr = (bpr + group - 1) // group
    t_start = (own + 1) % world
    out = []
    for vb0 in range(gpr * group * world):
        g, in_g = divmod(vb0, group)
        rr = (g + t_start) % world
        lb = (g // world) * group + in_g
        if lb < bpr:
            out.append((rr, lb))
    return out


def grad_tail_layout(eta: int, n: int, K: int):
    """(head floats, tail floats, padded tail stride) of a rank's gradient buffer: head = gs | go | gp rows, tail =
    Qo | Qs rows + eta*n coefficients + eta*n side flags (bytes); the stride keeps every rank's tail 16-byte aligned
    inside the all-gathered buffer (csrc/kge_train_fwd.cuh gbuf_floats)."""
    head = 3 * n * K
    tail = 2 * n * K + eta * n + (eta * n + 3) // 4
    return head, tail, (tail + 3) // 4 * 4


# ------------------------------------------------------------------------------------------------
# device memory that peers can map
# ------------------------------------------------------------------------------------------------
class _CudaArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerBuffer:
    """cudaMalloc'ed buffer (exportable with CUDA IPC) viewed as a torch tensor."""

    def __init__(self, eng, shape, dtype=torch.float32):
        self.eng = eng
        self.shape = tuple(int(s) for s in shape)
        nbytes = int(np.prod(self.shape)) * torch.empty((), dtype=dtype).element_size()
        p = C.c_void_p()
        check(eng.lib.kge_dev_alloc(max(nbytes, 256), C.byref(p)))
        self.ptr = int(p.value)
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
        self._arr = _CudaArray(self.ptr, self.shape, typestr)
        self.tensor = torch.as_tensor(self._arr, device=eng.tdev)
        self.peers = None  # device pointers of every rank's copy (own pointer at own rank)

    def export(self) -> bytes:
        return self.eng.ipc_export_ptr(self.ptr)

    def free(self):
        if self.ptr:
            self.tensor = None
            self.eng.lib.kge_dev_free(C.c_void_p(self.ptr))
            self.ptr = 0


def exchange_peers(eng, bufs, rank, world):
    """All-gather the IPC handles of `bufs` (list of PeerBuffer) and map every peer's copy."""
    handles = [b.export() for b in bufs]
    gathered = [None] * world
    dist.all_gather_object(gathered, handles)
    for i, b in enumerate(bufs):
        b.peers = [b.ptr if r == rank else eng.ipc_open(gathered[r][i]) for r in range(world)]


# ------------------------------------------------------------------------------------------------
# sharded training + ranking
# ------------------------------------------------------------------------------------------------
class ShardedKGE:
    """Row-sharded parameters of one model on `world` GPUs; `train_step` and `rank` are collective."""

    def __init__(self, model, k, eta, loss, optimizer, E, R, n_per_rank, *, lr=5e-4, margin=1.0, norm=1, seed=0,
                 init_ent=None, init_rel=None, device=None):
        assert dist.is_initialized(), "init torch.distributed (backend nccl) first"
        self.rank_id, self.world = dist.get_rank(), dist.get_world_size()
        assert self.world <= _lib.KGE_MAX_SHARDS
        self.eng = get_engine(device)
        eng = self.eng
        self.model, self.k, self.eta, self.loss, self.optimizer = model, k, eta, loss, optimizer
        self.E, self.R, self.n = int(E), int(R), int(n_per_rank)
        self.K = internal_k(model, k)
        self.mid = model_id(model, norm)
        self.rps = rows_per_shard(self.E, self.world)
        self.row_begin, self.row_end = shard_range(self.E, self.world, self.rank_id)
        K, rps = self.K, self.rps
        self.ent = PeerBuffer(eng, (rps, K))
        self.ent.tensor.zero_()
        n_loc = self.row_end - self.row_begin
        if init_ent is not None and n_loc > 0:
            self.ent.tensor[:n_loc].copy_(torch.as_tensor(np.ascontiguousarray(init_ent(self.row_begin, self.row_end), dtype=np.float32)))
        rel0 = init_rel() if init_rel is not None else np.zeros((R, K), np.float32)
        self.rel = torch.as_tensor(np.ascontiguousarray(rel0, dtype=np.float32)).to(eng.tdev)
        opt = _lib.OPT_IDS[optimizer]
        self.state = {}
        if opt == 0:
            self.state = dict(ent_m=torch.zeros((rps, K), device=eng.tdev), ent_v=torch.zeros((rps, K), device=eng.tdev),
                              rel_m=torch.zeros_like(self.rel), rel_v=torch.zeros_like(self.rel))
        elif opt == 1:
            self.state = dict(ent_m=torch.full((rps, K), 0.1, device=eng.tdev), rel_m=torch.full_like(self.rel, 0.1))
        elif opt == 2:
            self.state = dict(ent_m=torch.zeros((rps, K), device=eng.tdev), rel_m=torch.zeros_like(self.rel))
        # gradient buffer: [gs|go|gp] head (read by the owners over peer memory) + [Qo|Qs|coef|keep] tail
        # (all-gathered so that the per-negative reads of the reduction stay local)
        g_floats = eng.train_grad_floats(eta, self.n, K)
        self.g_head = eng.train_grad_head_floats(eta, self.n, K)
        self.tail_stride = (g_floats - self.g_head + 3) // 4 * 4
        assert (self.g_head, g_floats - self.g_head, self.tail_stride) == grad_tail_layout(eta, self.n, K)
        self.gbuf = PeerBuffer(eng, (self.g_head + self.tail_stride,))
        # staging copy of the entity row of every entity slot of this rank's batch, pushed by the owners
        self.ent_slots = (2 + eta) * self.n
        self.stage = PeerBuffer(eng, (self.ent_slots, K))
        exchange_peers(eng, [self.ent, self.gbuf, self.stage], self.rank_id, self.world)
        self.ent_table = make_table(self.ent.peers, rows=self.E, rows_per_shard=rps, K=K)
        self.S = (3 + eta) * self.n
        self.grads_table = make_table(self.gbuf.peers, rows=self.S * self.world, rows_per_shard=self.S, K=K)
        self.stage_table = make_table(self.stage.peers, rows=self.ent_slots * self.world, rows_per_shard=self.ent_slots, K=K)
        self.tails_all = torch.empty(self.tail_stride * self.world, dtype=torch.float32, device=eng.tdev)
        self.exchange = "push"  # "pull": fine-grained peer loads inside the kernels (the first design; A/B)
        self.timing = False
        self._marks = []
        self.keys_local = torch.empty(self.S, dtype=torch.int32, device=eng.tdev)
        self.keys_all = torch.empty(self.S * self.world, dtype=torch.int32, device=eng.tdev)
        self.loss_dev = torch.zeros(1, dtype=torch.float32, device=eng.tdev)
        self.loss_sum = torch.zeros(1, dtype=torch.float32, device=eng.tdev)
        self.sync_tok = torch.zeros(1, dtype=torch.float32, device=eng.tdev)
        self.step = 0
        self.kw = dict(model=self.mid, loss=_lib.LOSS_IDS[loss], opt=opt, k=k, eta=eta, margin=float(margin), lr=float(lr),
                       seed=int(seed), neg_index_base=neg_index_base(self.rank_id, eta, self.n))
        dist.barrier()

    def _state_tables(self):
        st = {}
        for name in ("ent_m", "ent_v"):
            if name in self.state:
                ptrs = [0] * self.world
                ptrs[self.rank_id] = self.state[name].data_ptr()
                st[name] = make_table(ptrs, rows=self.E, rows_per_shard=self.rps, K=self.K)
        for name in ("rel_m", "rel_v"):
            if name in self.state:
                st[name] = self.state[name]
        return st

    def make_args(self, pos_dev, repl=None, keep_subj=None, flags=0, step=None):
        push = self.exchange == "push"
        a = self.eng.train_args(ent=self.ent_table, rel=self.rel, pos=pos_dev, loss_out=self.loss_dev,
                                step=self.step if step is None else step, repl=repl, keep_subj=keep_subj, flags=flags,
                                **self.kw, **self._state_tables(),
                                stage=self.stage.tensor if push else None, grad_tails=self.tails_all if push else None,
                                grad_tail_stride=self.tail_stride)
        a._keep_more = (self.state, self.ent, self.gbuf, self.stage)
        return a

    def train_step(self, pos_dev, repl=None, keep_subj=None, flags=0):
        """pos_dev: this rank's int32 [n,3] positives (device).  Collective.  repl / keep_subj: optional
        supplied corruptions of this rank's positives (parity input), else in-kernel Philox."""
        assert pos_dev.shape[0] == self.n
        eng = self.eng
        self.step += 1
        push = self.exchange == "push"
        a = self.make_args(pos_dev, repl, keep_subj, flags)
        marks = []

        def mark(name):
            if self.timing:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))

        mark("start")
        eng.train_emit(a, self.keys_local)
        dist.all_gather_into_tensor(self.keys_all, self.keys_local)
        eng.train_select(a, self.keys_all, self.row_begin, self.row_end)
        mark("emit+keys")
        if push:
            # owners copy the rows of every rank's entity slots into that rank's staging buffer (peer stores)
            eng.train_push_rows(a, self.keys_all, self.stage_table, self.row_begin, self.row_end)
            mark("push")
            dist.all_reduce(self.sync_tok)  # every rank's pushes have landed once this returns on the stream
            mark("push_barrier")
        eng.train_fwd_bwd(a, self.gbuf.tensor)
        mark("fwd_bwd")
        if push:
            # the all-gather of the tails also orders every rank's forward (and its gradient buffer) before
            # the reduction below; the loss all-reduce at the end doubles as the closing barrier
            dist.all_gather_into_tensor(self.tails_all, self.gbuf.tensor[self.g_head:self.g_head + self.tail_stride])
            mark("tails")
            eng.train_apply(a, self.keys_all, self.grads_table, self.row_begin, self.row_end)
            mark("apply")
            # owners have finished reading the peers' gradient buffers / writing their rows once this returns
            self.loss_sum.copy_(self.loss_dev)
            dist.all_reduce(self.loss_sum)
            mark("loss+end_barrier")
        else:
            # every rank's forward reads and gradient buffer are complete once this returns on the stream
            self.loss_sum.copy_(self.loss_dev)
            dist.all_reduce(self.loss_sum)
            mark("loss")
            eng.train_apply(a, self.keys_all, self.grads_table, self.row_begin, self.row_end)
            mark("apply")
            dist.all_reduce(self.sync_tok)
            mark("end_barrier")
        if self.timing:
            self._marks.append(marks)
        return self.loss_sum

    def phase_times(self):
        """Average ms per phase over the steps run with self.timing = True (synchronises)."""
        torch.cuda.synchronize()
        acc = {}
        for marks in self._marks:
            for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
                acc[name] = acc.get(name, 0.0) + e0.elapsed_time(e1)
        n = max(1, len(self._marks))
        self._marks = []
        return {k: v / n for k, v in acc.items()}

    def rank_counts(self, test_dev, *, side=0, filtered=False, use_tensor_cores=False):
        """Per-shard sweep + all-reduce of the [T,2,4] counters.  Collective."""
        eng = self.eng
        n_loc = self.row_end - self.row_begin
        counts = eng.rank_counts(self.mid, self.k, self.ent_table, self.rel, test_dev, side=side, filtered=filtered,
                                 use_tensor_cores=use_tensor_cores, ent_local=self.ent.tensor[:max(n_loc, 1)],
                                 row_begin=self.row_begin, row_end=self.row_end)
        dist.all_reduce(counts)
        return counts

    def rank(self, test_dev, *, side=0, strategy=0, filtered=False, use_tensor_cores=False):
        counts = self.rank_counts(test_dev, side=side, filtered=filtered, use_tensor_cores=use_tensor_cores)
        return self.eng.rank_finalize(counts, side=side, strategy=strategy, filtered=filtered)

    def gather_entities(self):
        """Full [E,K] table on every rank (host), for checks."""
        parts = [torch.empty((self.rps, self.K), dtype=torch.float32, device=self.eng.tdev) for _ in range(self.world)]
        dist.all_gather(parts, self.ent.tensor.contiguous())
        return torch.cat(parts, 0)[: self.E].cpu().numpy()
