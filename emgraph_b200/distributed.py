"""One-process-per-GPU drivers of the multi-GPU path.

``ShardedKGE`` (the product path): the embedding tables and their optimizer state are split by COLUMN range over
the ranks of one NVSwitch domain and every rank processes the whole global batch on its slice; the only exchange
of a training step is the sum of one partial score per scored triple (260 bytes per positive at eta = 64).  For
ranking the column slices are transposed once into row-range shards (all-to-all) and every rank sweeps its rows;
the per-shard rank counts are all-reduced.

``RowShardedKGE`` (round-1 design, kept for A/B measurements): the entity table is split by contiguous row range;
``torch.distributed`` (NCCL) carries the all-gather of sort keys and the barriers, the rows travel over peer
memory (CUDA IPC mappings, owner-side push).

Replaces the reference's host-paged "large graph" mode (models/EmbeddingModel.py:645-666,
:1070-1097, :1251-1281), which is single-process and SGD-only.
"""
from __future__ import annotations

import ctypes as C
import os

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
class RowShardedKGE:
    """Row-sharded parameters of one model on `world` GPUs (round-1 design, kept for A/B: the exchange is NVLink-volume
    bound); `train_step` and `rank` are collective."""

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


# ------------------------------------------------------------------------------------------------
# dimension-sharded ("column-parallel") training, row-sharded ranking
# ------------------------------------------------------------------------------------------------
def dim_width(k: int, world: int) -> int:
    """Columns per half of every rank's slice: ceil(k / world) rounded up to a multiple of 4 (128-bit vectors);
    slices past the end of the model are zero columns, which every scoring function ignores and whose
    gradient is exactly zero."""
    return (-(-k // world) + 3) // 4 * 4


def dim_range(k: int, world: int, rank: int):
    """Logical columns [c0, c1) of the model's k that `rank` holds."""
    kc = dim_width(k, world)
    return min(k, rank * kc), min(k, (rank + 1) * kc)


def slice_columns(full, model: str, k: int, world: int, rank: int):
    """[rows, K] -> this rank's [rows, Kc] slice (ComplEx / HolE: the same column range of both halves, [re | im])."""
    full = np.asarray(full)
    kc = dim_width(k, world)
    c0, c1 = dim_range(k, world, rank)
    halves = 2 if model in ("ComplEx", "HolE") else 1
    out = np.zeros((full.shape[0], halves * kc), full.dtype)
    for h in range(halves):
        out[:, h * kc:h * kc + (c1 - c0)] = full[:, h * k + c0:h * k + c1]
    return out


def merge_columns(parts, model: str, k: int):
    """Inverse of slice_columns over all ranks: list of [rows, Kc] -> [rows, K]."""
    world = len(parts)
    kc = dim_width(k, world)
    halves = 2 if model in ("ComplEx", "HolE") else 1
    out = np.zeros((np.asarray(parts[0]).shape[0], halves * k), np.asarray(parts[0]).dtype)
    for r, p in enumerate(parts):
        c0, c1 = dim_range(k, world, r)
        for h in range(halves):
            out[:, h * k + c0:h * k + c1] = np.asarray(p)[:, h * kc:h * kc + (c1 - c0)]
    return out


def merge_index(model: str, k: int, world: int):
    """Column index into the rank-major concatenation [rank0 slice | rank1 slice | ...] (each Kc wide) that yields the
    model's [K] row: full[:, j] = cat[:, merge_index[j]]."""
    kc = dim_width(k, world)
    halves = 2 if model in ("ComplEx", "HolE") else 1
    Kc = halves * kc
    idx = np.empty(halves * k, np.int64)
    for r in range(world):
        c0, c1 = dim_range(k, world, r)
        for h in range(halves):
            idx[h * k + c0:h * k + c1] = r * Kc + h * kc + np.arange(c1 - c0)
    return idx


def chunk_bounds(n: int, chunks: int):
    """Positive ranges of the `chunks` pieces a step is cut into (the all-reduce of one piece overlaps the kernels
    of its neighbours); every piece non-empty."""
    chunks = max(1, min(int(chunks), n))
    base, extra = divmod(n, chunks)
    out, lo = [], 0
    for c in range(chunks):
        hi = lo + base + (1 if c < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


class ShardedKGE:
    """One model on `world` GPUs, tables split by column range; `train_step`, `rank`, `gather_*` are collective.

    n_per_rank positives per rank and step (weak scaling: the global batch is the concatenation of the ranks'
    batches in rank order; its corruption stream, loss and update are those of ONE GPU running that batch)."""

    def __init__(self, model, k, eta, loss, optimizer, E, R, n_per_rank, *, lr=5e-4, margin=1.0, norm=1, seed=0,
                 init_ent=None, init_rel=None, device=None, chunks=2, alpha=0.5, non_linearity="linear", side="s,o",
                 optimizer_params=None, pipeline=True, group=None, ent_slice=None, rel_slice=None, sorted_partial=None,
                 p2p_allreduce=None):
        self.group = group
        assert dist.is_initialized(), "init torch.distributed first"
        self.rank_id, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.eng = get_engine(device)
        eng = self.eng
        self.model, self.k, self.eta, self.loss, self.optimizer = model, int(k), int(eta), loss, optimizer
        self.E, self.R, self.n_local = int(E), int(R), int(n_per_rank)
        self.n = self.n_local * self.world
        self.K = internal_k(model, k)
        self.mid = model_id(model, norm)
        self.kc = dim_width(self.k, self.world)
        self.Kc = internal_k(model, self.kc)
        dev = eng.tdev

        def local(full_or_fn, rows):
            if full_or_fn is None:
                return torch.zeros((rows, self.Kc), dtype=torch.float32, device=dev)
            full = full_or_fn() if callable(full_or_fn) else full_or_fn
            return torch.from_numpy(np.ascontiguousarray(slice_columns(full, model, self.k, self.world, self.rank_id), dtype=np.float32)).to(dev)

        # ent_slice / rel_slice: this rank's [rows, Kc] slices already on the device (large tables are drawn per rank)
        self.ent = local(init_ent, self.E) if ent_slice is None else ent_slice.to(dev).contiguous()
        self.rel = local(init_rel, self.R) if rel_slice is None else rel_slice.to(dev).contiguous()
        assert tuple(self.ent.shape) == (self.E, self.Kc) and tuple(self.rel.shape) == (self.R, self.Kc)
        opt = _lib.OPT_IDS[optimizer]
        self.state = {}
        if opt == 0:
            self.state = dict(ent_m=torch.zeros_like(self.ent), ent_v=torch.zeros_like(self.ent),
                              rel_m=torch.zeros_like(self.rel), rel_v=torch.zeros_like(self.rel))
        elif opt == 1:
            self.state = dict(ent_m=torch.full_like(self.ent, 0.1), rel_m=torch.full_like(self.rel, 0.1))
        elif opt == 2:
            self.state = dict(ent_m=torch.zeros_like(self.ent), rel_m=torch.zeros_like(self.rel))
        self.chunks = int(chunks)
        # phase 1 over the SORTED slot list (every entity row streamed once, in address order) instead of one random gather
        # per scored triple (kge_train_partial_sorted): an alternative for narrow slices and batches dense in the table.
        # Measured on B200 (profiles/r02_summary.md): at cfg5 / 8 ranks the sorted kernel takes 250 us against 198 us for the
        # random gather -- the query gather then misses L2 half of the time -- so it is OFF unless asked for.
        self.sorted_partial = False if sorted_partial is None else bool(sorted_partial)
        n_sums = ((1 + self.eta) * self.n + 7) // 4 * 4
        # the sum over the ranks: NCCL's all-reduce by default; p2p_allreduce=True / KGE_P2P_ALLREDUCE=1 selects the library's own
        # exchange over peer memory (kge_allreduce_p2p: P2P loads and stores through CUDA-IPC mappings, flags in peer memory).
        # Measured on 8 x B200 (profiles/r02_summary.md): NCCL (in-switch reduction) moves the 10.7 MB pieces of cfg5 faster than
        # a plain pull/push kernel of a size that does not disturb the phase kernels, so it stays the default
        if p2p_allreduce is None:
            p2p_allreduce = os.environ.get("KGE_P2P_ALLREDUCE", "0")[:1] == "1"
        self.p2p = bool(p2p_allreduce) and self.world > 1 and dev.type == "cuda" and hasattr(eng, "allreduce_p2p")
        if self.p2p:
            self._sums_pb = PeerBuffer(eng, (n_sums,))
            self._tot_pb = PeerBuffer(eng, (n_sums,))
            self._flag_pb = PeerBuffer(eng, (64,), dtype=torch.int32)
            for b in (self._sums_pb, self._tot_pb, self._flag_pb):
                b.tensor.zero_()
            torch.cuda.synchronize(dev)
            exchange_peers(eng, [self._sums_pb, self._tot_pb, self._flag_pb], self.rank_id, self.world)
            self.sums_flat, self.totals_flat = self._sums_pb.tensor, self._tot_pb.tensor
            self._p2p_tables = tuple(make_table(b.peers, rows=n_sums * self.world, rows_per_shard=n_sums, K=1)
                                     for b in (self._sums_pb, self._tot_pb, self._flag_pb))
            self._p2p_seq = 0
            # the exchange of piece c runs on its own (high-priority) stream beside the phase kernel of piece c+1
            self._ar_stream = torch.cuda.Stream(device=dev, priority=-1)
            self._ar_events = []
        else:
            self.sums_flat = torch.zeros(n_sums, dtype=torch.float32, device=dev)  # chunk c at (1+eta)*lo_c
            self.totals_flat = self.sums_flat  # NCCL reduces in place
        self.pos_all = torch.empty((self.n, 3), dtype=torch.int32, device=dev)
        self.loss_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step = 0
        self.pipeline = bool(pipeline)
        op = dict(optimizer_params or {})
        self.kw = dict(model=self.mid, loss=_lib.LOSS_IDS[loss], opt=opt, k=self.kc, k_model=self.k, eta=self.eta, margin=float(margin),
                       lr=float(op.get("lr", lr)), seed=int(seed), alpha=float(alpha), side=_lib.TRAIN_SIDE_IDS[side],
                       non_linearity=_lib.NL_IDS[non_linearity], beta1=float(op.get("beta1", 0.9)), beta2=float(op.get("beta2", 0.999)),
                       eps=float(op.get("epsilon", 1e-7)), momentum=float(op.get("momentum", 0.9)))
        self.timing = False
        self._marks = []
        self._rows = None      # cached row-range shard [rps, K] of the current parameters (ranking)
        self._rows_step = -1
        self.rps = rows_per_shard(self.E, self.world)
        self.row_begin, self.row_end = shard_range(self.E, self.world, self.rank_id)
        self._merge_idx = torch.from_numpy(merge_index(model, self.k, self.world)).to(dev)
        dist.barrier(group)

    # ---------------------------------------------------------------- training
    def gather_batch(self, pos_local):
        """This rank's [n_local,3] positives -> the global batch (rank order) on every rank."""
        assert pos_local.shape[0] == self.n_local
        if self.world == 1:
            return pos_local
        dist.all_gather_into_tensor(self.pos_all, pos_local.contiguous(), group=self.group)
        return self.pos_all

    def make_args(self, pos_all, repl=None, keep_subj=None, flags=0, step=None, **dbg):
        a = self.eng.train_args(ent=self.ent, rel=self.rel, pos=pos_all, loss_out=self.loss_dev,
                                step=self.step if step is None else step, repl=repl, keep_subj=keep_subj, flags=flags,
                                **self.kw, **self.state, **dbg)
        return a

    def train_step(self, pos_local, repl=None, keep_subj=None, flags=0, pos_is_global=False, **dbg):
        """One optimisation step on the global batch.  pos_local: this rank's int32 [n_local,3] positives (device), or
        the whole [n,3] global batch when pos_is_global (every rank must pass the same).  repl / keep_subj: optional
        corruptions of the GLOBAL batch (parity input, identical on every rank), else the in-kernel Philox stream of
        the global batch.  Returns the device scalar holding the batch loss (identical on every rank)."""
        eng = self.eng
        pos_all = pos_local if pos_is_global else self.gather_batch(pos_local)
        n = pos_all.shape[0]
        assert 0 < n <= self.n, "global batch of %d positives; this model was sized for %d" % (n, self.n)
        # small batches: the all-reduce payload is a few hundred KB and latency-bound -- cutting it up only adds launches
        bounds = chunk_bounds(n, self.chunks if n * (1 + self.eta) * 4 >= (256 << 10) else 1)
        e1 = 1 + self.eta
        sums = [self.sums_flat[e1 * lo:e1 * hi] for lo, hi in bounds]
        totals = [self.totals_flat[e1 * lo:e1 * hi] for lo, hi in bounds]
        self.step += 1
        # the side-stream prologue may only read batches that were resident before the previous step was submitted
        pipe = self.pipeline and pos_is_global and repl is None and keep_subj is None
        a = self.make_args(pos_all, repl, keep_subj, flags | (_lib.F_PIPELINE if pipe else 0), **dbg)
        marks = []

        def mark(name):
            if self.timing:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))

        mark("start")
        def reduce_piece(c):
            """sum of piece c over the ranks: in-stream over peer memory, or an asynchronous NCCL all-reduce"""
            if self.world == 1:
                return None
            if self.p2p:
                lo, hi = bounds[c]
                off = e1 * lo // 4 * 4  # the kernel works on 16-byte vectors: the range is rounded outwards (the few
                end = (e1 * hi + 3) // 4 * 4  # neighbouring floats it also sums are rewritten by their own piece)
                self._p2p_seq += 1
                while len(self._ar_events) <= 2 * c + 1:
                    self._ar_events.append(torch.cuda.Event())
                ready, done = self._ar_events[2 * c], self._ar_events[2 * c + 1]
                main = torch.cuda.current_stream()
                ready.record(main)
                with torch.cuda.stream(self._ar_stream):
                    self._ar_stream.wait_event(ready)
                    eng.allreduce_p2p(*self._p2p_tables, self.rank_id, off, end - off, self._p2p_seq)
                    done.record(self._ar_stream)
                return done
            return dist.all_reduce(sums[c], group=self.group, async_op=True)

        works = []
        if self.sorted_partial:
            eng.train_partial_sorted(a, self.sums_flat, len(bounds))
            mark("partial_sorted")
            works = [reduce_piece(c) for c in range(len(bounds))]
        else:
            for c, (lo, hi) in enumerate(bounds):
                eng.train_partial(a, sums[c], lo, hi)
                mark("partial%d" % c)
                works.append(reduce_piece(c))
        for c, (lo, hi) in enumerate(bounds):
            if works[c] is not None:
                if self.p2p:
                    torch.cuda.current_stream().wait_event(works[c])
                else:
                    works[c].wait()
            mark("allreduce%d" % c)
            eng.train_backward(a, totals[c], lo, hi)
            mark("backward%d" % c)
        eng.train_reduce(a)
        mark("reduce_apply")
        if self.timing:
            self._marks.append(marks)
        return self.loss_dev

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

    def train_step_host(self, pos_host):
        """The step fed the way the reference feeds it (models/EmbeddingModel.py:1329-1337, :1421): this rank's batch comes
        from (pinned) host memory and the batch loss goes back to the host -- one step late, so that the GPU always has the
        next step queued: returns the loss of the step submitted one call earlier (None on the first call);
        host_flush() returns the last one.  Every step copies its batch in and its loss out."""
        assert (not pos_host.is_cuda) and pos_host.dtype == torch.int32 and pos_host.shape[0] == self.n_local
        dev = self.eng.tdev
        if not hasattr(self, "_hb"):
            self._hb = dict(stage=[torch.empty((self.n_local, 3), dtype=torch.int32, device=dev) for _ in range(2)],
                            ring=torch.zeros(4, dtype=torch.float32).pin_memory(),
                            ev=[torch.cuda.Event() for _ in range(4)], pending=None, tick=0)
        hb = self._hb
        i, st = hb["tick"] % 4, hb["stage"][hb["tick"] % 2]
        hb["tick"] += 1
        st.copy_(pos_host, non_blocking=True)
        self.train_step(st)
        hb["ring"][i:i + 1].copy_(self.loss_dev, non_blocking=True)
        hb["ev"][i].record()
        prev, hb["pending"] = hb["pending"], i
        if prev is None:
            return None
        hb["ev"][prev].synchronize()
        return float(hb["ring"][prev])

    def host_flush(self):
        hb = getattr(self, "_hb", None)
        if hb is None or hb["pending"] is None:
            return None
        prev, hb["pending"] = hb["pending"], None
        hb["ev"][prev].synchronize()
        return float(hb["ring"][prev])

    # ---------------------------------------------------------------- parameters
    def _gather_cols(self, t):
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t.contiguous(), group=self.group)
        return torch.cat(parts, 1).index_select(1, self._merge_idx)

    def gather_entities(self):
        """Full [E,K] table on every rank (host); for checks and checkpoints of small models."""
        return self._gather_cols(self.ent).cpu().numpy()

    def gather_relations(self, device=False):
        full = self._gather_cols(self.rel)
        return full if device else full.cpu().numpy()

    def gather_state(self, name):
        """Full-model optimizer-state table `name` (ent_m / ent_v / rel_m / rel_v) on every rank, as a CPU tensor."""
        return self._gather_cols(self.state[name]).cpu()

    def row_shard(self):
        """[rps, K] full-width rows [row_begin,row_end) of the CURRENT parameters (zero rows past the end): the
        column slices are transposed into row-range shards by one all-to-all; cached until the next step."""
        if self._rows is not None and self._rows_step == self.step:
            return self._rows
        W, rps, Kc = self.world, self.rps, self.Kc
        send = torch.zeros((W * rps, Kc), dtype=torch.float32, device=self.eng.tdev)
        send[: self.E].copy_(self.ent)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        # recv[r*rps + t] = rank r's columns of my row t  ->  [rps, W*Kc]  ->  the model's column order
        cat = recv.view(W, rps, Kc).permute(1, 0, 2).reshape(rps, W * Kc)
        self._rows = cat.index_select(1, self._merge_idx).contiguous()
        self._rows_step = self.step
        return self._rows

    # ---------------------------------------------------------------- ranking
    def rank_counts(self, test_dev, *, side=0, filtered=False, use_tensor_cores=False, non_linearity=0):
        """Per-shard sweep of this rank's rows + all-reduce of the [T,2,4] counters.  Collective."""
        eng = self.eng
        rows = self.row_shard()
        rel = self.gather_relations(device=True)
        T = test_dev.shape[0]
        # subject / object rows of the test triples: every rank has their columns, one all-gather makes them whole
        s_rows = self._gather_cols(self.ent.index_select(0, test_dev[:, 0].long()))
        o_rows = self._gather_cols(self.ent.index_select(0, test_dev[:, 2].long()))
        n_loc = self.row_end - self.row_begin
        counts = eng.rank_counts_rows(self.mid, self.k, self.E, rel, s_rows, o_rows, rows[:max(n_loc, 1)], test_dev,
                                      row_begin=self.row_begin, row_end=self.row_end, side=side, filtered=filtered,
                                      use_tensor_cores=use_tensor_cores, non_linearity=non_linearity)
        dist.all_reduce(counts, group=self.group)
        return counts

    def rank(self, test_dev, *, side=0, strategy=0, filtered=False, use_tensor_cores=False, non_linearity=0):
        counts = self.rank_counts(test_dev, side=side, filtered=filtered, use_tensor_cores=use_tensor_cores, non_linearity=non_linearity)
        return self.eng.rank_finalize(counts, side=side, strategy=strategy, filtered=filtered)


# ------------------------------------------------------------------------------------------------
# fit(engine_params={"n_gpus": N}) from a plain process: one worker per GPU
# ------------------------------------------------------------------------------------------------
def _free_port():
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn_worker(rank, world, port, cls_name, hyper, resume, Xi_path, E, R, out_path, backend):
    import os
    import pickle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if backend == "nccl":
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                                pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
    else:
        dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        from . import models
        hyper = dict(hyper)
        ep = dict(hyper.get("engine_params") or {})
        if backend == "nccl":
            ep["device"] = rank
        hyper["engine_params"] = ep
        m = getattr(models, cls_name)(**hyper)
        if resume is not None:
            m.trained_model_params = [resume["ent"], resume["rel"]]
            m._opt_state = {k: torch.from_numpy(v) for k, v in resume["state"].items()}
            m._opt_step = int(resume["step"])
            m._resume = True
            m.is_fitted = True
        m._fit_sharded_spmd(np.load(Xi_path), E, R)
        if rank == 0:
            with open(out_path, "wb") as fw:
                pickle.dump(dict(ent=m.trained_model_params[0], rel=m.trained_model_params[1], step=m._opt_step, loss_history=m.loss_history,
                                 state={k: v.numpy() for k, v in m._opt_state.items()}), fw, protocol=pickle.HIGHEST_PROTOCOL)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def spawn_fit(model, Xi, E, R, n_gpus, backend="nccl"):
    """Run model._fit_sharded_spmd on `n_gpus` freshly spawned workers; returns dict(ent, rel, state, step, loss_history)."""
    import os
    import pickle
    import tempfile
    import torch.multiprocessing as mp
    if backend == "nccl" and torch.cuda.device_count() < n_gpus:
        raise _lib.KgeError("engine_params['n_gpus']=%d but only %d CUDA devices are visible" % (n_gpus, torch.cuda.device_count()))
    resume = None
    if getattr(model, "_resume", False):
        st = getattr(model, "_opt_state", None) or {}
        resume = dict(ent=np.asarray(model.trained_model_params[0]), rel=np.asarray(model.trained_model_params[1]), step=int(getattr(model, "_opt_step", 0)),
                      state={k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in st.items()})
    with tempfile.TemporaryDirectory(prefix="kge_fit_") as d:
        xi_path, out_path = os.path.join(d, "xi.npy"), os.path.join(d, "out.pkl")
        np.save(xi_path, np.ascontiguousarray(Xi, dtype=np.int32))
        ctx = mp.get_context("spawn")
        port = _free_port()
        procs = [ctx.Process(target=_spawn_worker, args=(r, n_gpus, port, model.__class__.__name__, dict(model.all_params, engine_params=dict(model.engine_params)), resume, xi_path, E, R,
                                                         out_path, backend)) for r in range(n_gpus)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        bad = [p.exitcode for p in procs if p.exitcode != 0]
        if bad or not os.path.exists(out_path):
            raise _lib.KgeError("multi-GPU fit failed (worker exit codes %r)" % ([p.exitcode for p in procs],))
        with open(out_path, "rb") as fr:
            return pickle.load(fr)
