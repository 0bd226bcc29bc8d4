"""Host-side logic of the row-sharded multi-GPU path, exercised with world_size 2 on the gloo backend
(no GPU): shard ranges, batch dealing, Philox stream bases, and that the owner selection applied to an
all-gathered key list covers every entity slot exactly once (relation slots on every rank)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from emgraph_b200 import distributed as D


def test_shard_ranges_cover_rows_exactly():
    for E in (1, 7, 8, 14541, 4594485):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(E, np.int32)
            for r in range(world):
                b, e = D.shard_range(E, world, r)
                assert 0 <= b <= e <= E and e - b <= D.rows_per_shard(E, world)
                seen[b:e] += 1
                if e > b:
                    assert np.all(D.owner_of(np.arange(b, e), E, world) == r)
            assert np.all(seen == 1)


def test_batches_are_disjoint_across_ranks_within_a_step():
    n_total, n_per = 10000, 128
    for world in (2, 4, 8):
        for step in (0, 1, 5, 77):
            sl = [D.batch_slice(n_total, world, r, step, n_per) for r in range(world)]
            assert all(hi - lo == n_per and hi <= n_total for lo, hi in sl)
            assert len({lo for lo, _ in sl}) == world
    # negative index bases never overlap
    assert [D.neg_index_base(r, 20, 128) for r in range(3)] == [0, 2560, 5120]


def test_push_schedule_visits_every_block_once_and_spreads_destinations():
    for S, world in ((67 * 10308, 8), (23 * 4252, 2), (100, 4), (31, 3), (32 * 5, 5)):
        bpr = (S + 31) // 32
        for group in (0, 1, 3, 64):
            firsts = []
            for own in range(world):
                order = D.push_block_order(S, world, own, group)
                assert len(order) == bpr * world and len(set(order)) == bpr * world
                assert all(0 <= r < world and 0 <= b < bpr for r, b in order)
                firsts.append(order[0][0])
            assert sorted(firsts) == list(range(world))  # no two owners start on the same destination
    # default grouping: one destination region at a time
    order = D.push_block_order(320, 4, 1, 0)
    assert [r for r, _ in order] == [2] * 10 + [3] * 10 + [0] * 10 + [1] * 10


def test_grad_tail_layout_alignment():
    for eta, n, K in ((64, 10308, 256), (20, 4252, 400), (5, 3, 8), (7, 33, 10)):
        head, tail, stride = D.grad_tail_layout(eta, n, K)
        assert head == 3 * n * K and stride >= tail and stride % 4 == 0 and stride - tail < 4
        assert tail * 4 >= 2 * n * K * 4 + eta * n * 4 + eta * n  # rows + coefficients + one byte per negative


def _worker(rank, world, port, E, R, eta, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(10 + rank)
        pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
        repl = rng.integers(0, E, eta * n).astype(np.int32)
        # slot keys in the library's layout: subjects | objects | replacements | E + relation
        keys_local = torch.from_numpy(np.concatenate([pos[:, 0], pos[:, 2], repl, E + pos[:, 1]]).astype(np.int32))
        S = (3 + eta) * n
        assert keys_local.numel() == S
        keys_all = torch.empty(S * world, dtype=torch.int32)
        dist.all_gather_into_tensor(keys_all, keys_local)
        ka = keys_all.numpy()
        np.testing.assert_array_equal(ka[rank * S:(rank + 1) * S], keys_local.numpy())
        mine = D.owned_mask(ka, E, world, rank)
        cover = torch.from_numpy(mine.astype(np.int32))
        dist.all_reduce(cover)
        ent_slot = ka < E
        ok = bool(np.all(cover.numpy()[ent_slot] == 1) and np.all(cover.numpy()[~ent_slot] == world))
        # the counters of a sharded ranking sweep add up: per-shard candidate counts -> all-reduce
        b, e = D.shard_range(E, world, rank)
        cnt = torch.tensor([e - b], dtype=torch.int32)
        dist.all_reduce(cnt)
        q.put((rank, ok and int(cnt.item()) == E))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gloo_world2_key_gather_and_owner_selection():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1001, 7, 5, 64, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


# ------------------------------------------------------------------------------------------------
# owner-compute exchange (the round-2 plan): the decomposition must be exact
# ------------------------------------------------------------------------------------------------
import pytest  # noqa: E402

from oracle import kge_oracle as ko  # noqa: E402
from oracle import sharded_oracle as so  # noqa: E402


def _rank_batches(rng, E, R, n, eta, W):
    out = []
    for _ in range(W):
        pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
        out.append((pos, rng.integers(0, 2, n * eta).astype(np.uint8), rng.integers(0, E, n * eta).astype(np.int32)))
    return out


@pytest.mark.parametrize("model,loss,norm,nl", [("DistMult", "nll", 1, "linear"), ("ComplEx", "multiclass_nll", 1, "linear"),
                                                ("HolE", "self_adversarial", 1, "tanh"), ("TransE", "pairwise", 1, "linear"),
                                                ("TransE", "nll", 2, "linear"), ("ComplEx", "absolute_margin", 1, "sigmoid")])
def test_owner_compute_step_equals_the_oracle_step(model, loss, norm, nl):
    """Folded queries shipped to the owners, owner-side scores, per-owner partial sums of c*dF/dQ sent back: the
    summed loss and gradients of W ranks equal the single-process oracle step on the same batches."""
    rng = np.random.default_rng(7)
    E, R, k, eta, n, W = 97, 5, 6, 5, 23, 4
    K = ko.internal_k(model, k)
    ent = (rng.normal(size=(E, K)) * 0.5).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.5).astype(np.float32)
    batches = _rank_batches(rng, E, R, n, eta, W)
    got = so.owner_compute_step(model, k, loss, eta, ent, rel, batches, W, margin=2.0, norm=norm, nl=nl)
    exp_loss, exp_ge, exp_gr = 0.0, np.zeros((E, K)), np.zeros((R, K))
    for pos, keep, repl in batches:
        o = ko.train_step(model, k, loss, eta, ent, rel, pos, keep, repl, margin=2.0, norm=norm, dtype=np.float64, nl=nl)
        exp_loss += float(o["loss"])
        exp_ge += o["grad_ent"]
        exp_gr += o["grad_rel"]
    np.testing.assert_allclose(got["loss"], exp_loss, rtol=1e-10)
    np.testing.assert_allclose(got["grad_ent"], exp_ge, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(got["grad_rel"], exp_gr, rtol=1e-9, atol=1e-11)


def test_owner_compute_volume_against_the_row_push():
    """Per-rank NVLink volume of the two exchanges at a cfg5-like shape (eta = 64, 8 ranks, 1 KiB rows): shipping
    queries and partial sums moves under half of what shipping rows moves (DESIGN.md section 7 sizes round 2 with
    this: ~33 rows per positive against ~74)."""
    rng = np.random.default_rng(3)
    E, R, k, eta, n, W = 20000, 9, 256, 64, 120, 8
    ent = np.zeros((E, k), np.float32)
    rel = np.zeros((R, k), np.float32)
    batches = _rank_batches(rng, E, R, n, eta, W)
    oc = so.owner_compute_step("DistMult", k, "nll", eta, ent, rel, batches, W)
    push = so.push_step_bytes(eta, n, k, W, E, batches)
    row = 4 * k
    assert oc["bytes"]["queries_in"] == 2 * n * (W - 1) * row
    assert oc["bytes"]["partials_in"] <= 2 * n * (W - 1) * row
    assert abs(push["rows_in"] - (2 + eta) * n * row * (W - 1) / W) < 0.05 * push["rows_in"]
    assert 30 * n * row < oc["bytes_total"] < 36 * n * row
    assert 72 * n * row < push["total"] < 76 * n * row
