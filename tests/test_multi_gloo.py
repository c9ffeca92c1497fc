"""N > 1 host-side logic on CPU: world_size-2 gloo processes shard the views, all-gather their coverage rows the way
prv_allgather_bitsets does (rank-major), and must select the same greedy sequence as a single process."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _greedy_with_ids(orc, rows, ids, first_id, max_iter):
    """oracle greedy over a table whose row r has global id ids[r]: reorder rows by id (ties break on the id)."""
    order = np.argsort(ids, kind="stable")
    seq, gain, cov, _ = orc.greedy(rows[order], int(np.where(ids[order] == first_id)[0][0]), max_iter)
    return ids[order][seq].tolist(), gain.tolist()


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import load_pkg
    import oracle as orc
    prv = load_pkg.load()
    from nerf_prv_b200 import sharding, synth
    V = 7  # odd on purpose: rank 1 gets a padded (empty) view
    w = synth.build_workload(prv, "C1", n_views=V, size=(64, 48))
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    it = orc.make_intrinsics(w["intr"].width, w["intr"].height, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, w["intr"].model,
                             list(w["intr"].coeffs))
    words = orc.bitset_words(m.n)
    ids = sharding.pad_view_ids(sharding.shard_view_ids(V, rank, world), V, rank, world)
    rows = np.zeros((len(ids), words), dtype=np.uint64)
    for r, v in enumerate(ids):
        if v < V:
            ok, ranks, _ = m.cast_view_dense(it, w["pose_world"][v], w["init_pos"][v], want_depth=False, num_threads=1)
            rows[r] = orc.bitset_from_ranks(ranks, words)
    # rank-major all-gather, exactly the layout ncclAllGather produces for prv_allgather_bitsets_async
    t = torch.from_numpy(rows.view(np.int64))
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    table = torch.cat(out).numpy().view(np.uint64)
    all_ids = sharding.gathered_view_ids(V, world)
    assert np.array_equal(all_ids[rank * len(ids):(rank + 1) * len(ids)], ids)
    seq, gain = _greedy_with_ids(orc, table, all_ids, 0, 64)
    np.save(os.path.join(tmp, "seq_%d.npy" % rank), np.array(seq + gain))
    if rank == 0:
        # single-process reference
        full = np.zeros((V, words), dtype=np.uint64)
        for v in range(V):
            ok, ranks, _ = m.cast_view_dense(it, w["pose_world"][v], w["init_pos"][v], want_depth=False, num_threads=1)
            full[v] = orc.bitset_from_ranks(ranks, words)
        s1, g1, _, _ = orc.greedy(full, 0, 64)
        np.save(os.path.join(tmp, "single.npy"), np.array(s1.tolist() + g1.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_view_sharding_matches_single(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    single = np.load(tmp_path / "single.npy")
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("seq_%d.npy" % r)), single)


def test_sharding_helpers():
    sys.path.insert(0, ROOT)
    import load_pkg
    load_pkg.load()
    from nerf_prv_b200 import sharding
    assert sharding.shard_view_ids(10, 1, 4).tolist() == [1, 5, 9]
    assert sharding.views_per_rank(10, 4) == 3
    ids = sharding.gathered_view_ids(10, 4)
    assert len(ids) == 12 and sorted(ids.tolist())[:10] == list(range(10)) and len(set(ids.tolist())) == 12
    assert sharding.gathered_view_ids(1024, 8).tolist()[:3] == [0, 8, 16]
    assert sharding.shard_objects(64, 3, 8) == [3, 11, 19, 27, 35, 43, 51, 59]
