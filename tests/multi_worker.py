"""Worker of tests/test_gpu_multi.py::test_small_sharded_paths (one process per GPU, launched by torch.distributed.run).

Small view-sharded workload (13 views of C1 at 160x120: an uneven split, so padding rows exist) against the oracle, for every
way the coverage rows can travel: peer-memory stores fused into the count kernel (PRV_CAST_PUBLISH), peer-memory stores from
the scoring stream (no publish flag), and ncclAllGather -- each with several steps enqueued back to back (the scoring of step k
overlapping the cast of step k+1 on the ctx's second stream)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    import torch.distributed as dist
    import load_pkg
    import oracle as orc
    prv = load_pkg.load()
    from nerf_prv_b200 import sharding, synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    V = 13
    w = synth.build_workload(prv, "C1", n_views=V, size=(160, 120))
    m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
    it = orc.make_intrinsics(w["intr"].width, w["intr"].height, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, w["intr"].model, list(w["intr"].coeffs))
    words = orc.bitset_words(m.n)
    rows = np.stack([orc.bitset_from_ranks(m.cast_view_dense(it, w["pose_world"][v], w["init_pos"][v])[1], words) for v in range(V)])
    o_seq, o_gain, o_cov, _ = orc.greedy(rows, 0, 64)

    ids = sharding.pad_view_ids(sharding.shard_view_ids(V, rank, world), V, rank, world)
    real = ids < V
    pose = np.ascontiguousarray(np.where(real[:, None, None], w["pose_world"][np.minimum(ids, V - 1)], np.eye(4)[None]))
    init = np.ascontiguousarray(np.where(real[:, None], w["init_pos"][np.minimum(ids, V - 1)], 1.0e6))
    ctx = prv.Context(local)
    ctx.set_map(w["keys"], w["map_rgb"], w["resolution"])
    ctx.set_camera(w["intr"], 1.0)
    uid = prv.comm_unique_id() if rank == 0 else bytes(128)
    t = torch.tensor(list(uid), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    ctx.comm_init(bytes(t.cpu().tolist()), rank, world)

    def check(tag):
        g_rows, g_ids = ctx.get_gathered()
        assert sorted(g_ids.tolist()) == sorted(sharding.gathered_view_ids(V, world).tolist()), tag
        table = np.zeros((V, ctx.words), dtype=np.uint64)
        table[g_ids[g_ids < V]] = g_rows[g_ids < V]
        assert np.array_equal(table, rows), tag + ": gathered rows differ from the oracle"
        assert not g_rows[g_ids >= V].any(), tag + ": padding rows must be empty"
        seq, gain, cov = ctx.get_greedy()
        assert seq.tolist() == o_seq.tolist() and gain.tolist() == o_gain.tolist() and np.array_equal(cov, o_cov), tag + ": greedy differs"

    for transport, publish in (("nccl", False), ("p2p", True), ("p2p", False), ("p2p", True)):
        if transport == "p2p":
            handles = [None] * world
            dist.all_gather_object(handles, ctx.p2p_export(1 << 20))
            ctx.p2p_import(handles, rank, world)
        ctx.set_views(pose, init, view_ids=ids)
        dist.barrier()
        for rep in range(2):
            for _ in range(5):  # five steps back to back: the second stream overlaps scoring and casting
                ctx.cast_async(prv.MODE_DENSE, want_pixels=True, publish=publish)
                ctx.allgather_bitsets_async()
                ctx.greedy_async(0, 64)
            check("%s publish=%s rep %d" % (transport, publish, rep))
            counts = ctx.get_coverage_counts()
            assert counts.tolist() == [int(np.unpackbits(rows[i].view(np.uint8)).sum()) if i < V else 0 for i in ids.tolist()]
        ctx.sync()
        dist.barrier()
        if transport == "p2p":
            ctx.comm_destroy_p2p()
            dist.barrier()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_WORKER_OK world=%d" % world)


if __name__ == "__main__":
    main()
