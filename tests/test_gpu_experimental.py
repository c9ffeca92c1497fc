"""Opt-in paths that are exact on the CPU check (tests/test_kernel_on_host.py) but have not been measured on a B200 yet.
Skipped unless PRV_TEST_EXPERIMENTAL=1, so the default GPU suite only holds what the default path runs:

    PRV_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q
"""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("PRV_TEST_EXPERIMENTAL") != "1", reason="set PRV_TEST_EXPERIMENTAL=1")]


@pytest.mark.parametrize("name,n_views,size", [("C1", 4, (160, 120)), ("C1", 6, (640, 480)), ("C2", 8, (640, 480))])
def test_fine_cull_is_exact_and_marches_fewer_rays(prv, orc, synth, name, n_views, size):
    """prv_set_fine_cull: second level of the brick cull (cells of 4 / 2 / 1 voxels).  Same rows, ranks, depths and greedy
    sequence as the default pipeline (and as the oracle on the small case); fewer rays reach the march kernel."""
    w = synth.build_workload(prv, name, n_views=n_views, size=size)
    c = prv.Context(0)
    try:
        def run(cell, entry=False):
            c.set_fine_cull(cell, entry)
            c.set_map(w["keys"], w["map_rgb"], w["resolution"])
            c.set_camera(w["intr"], 1.0)
            bits, counts, hit, depth = c.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_DENSE, want_hit_rank=True, want_depth=True)
            st = c.get_cast_stats()
            seq, gains = c.greedy(0, 64)
            return bits, counts, hit, depth, st, seq, gains
        base = run(0)
        if size[0] <= 160:
            m = orc.Map.from_keys(w["keys"], w["map_rgb"], w["resolution"])
            it = orc.make_intrinsics(w["intr"].width, w["intr"].height, w["intr"].fx, w["intr"].fy, w["intr"].ppx, w["intr"].ppy, w["intr"].model,
                                     list(w["intr"].coeffs))
            for v in range(n_views):
                ok, r, d = m.cast_view_dense(it, w["pose_world"][v], w["init_pos"][v])
                assert np.array_equal(base[2][v], r) and np.array_equal(base[3][v], d)
        prev = base[4]["marched"]
        for cell in (4, 2, 1):
            got = run(cell)
            for a, b in zip(got[:4], base[:4]):
                assert np.array_equal(a, b), "fine cull %d changed a result" % cell
            assert got[5].tolist() == base[5].tolist() and got[6].tolist() == base[6].tolist()
            assert got[4]["hits"] == base[4]["hits"] and got[4]["rays"] == base[4]["rays"]
            assert got[4]["hits"] <= got[4]["marched"] <= prev
            prev = got[4]["marched"]
            # the exact march started at the first set fine cell: same results, same DDA steps, fewer probes
            ent = run(cell, True)
            for a, b in zip(ent[:4], base[:4]):
                assert np.array_equal(a, b), "fine cull %d with entry at the cell changed a result" % cell
            assert ent[5].tolist() == base[5].tolist() and ent[6].tolist() == base[6].tolist()
            assert ent[4]["marched"] == got[4]["marched"] and ent[4]["steps"] == got[4]["steps"] and ent[4]["hits"] == base[4]["hits"]
            assert ent[4]["probes_in"] < got[4]["probes_in"]
        assert prev < base[4]["marched"]
        # voxel-driven mode goes through the same coarse kernel
        c.set_fine_cull(1, True)
        c.set_map(w["keys"], w["map_rgb"], w["resolution"])
        c.set_camera(w["intr"], 1.0)
        bv1, cv1, hv1, _ = c.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_VOXEL, want_hit_rank=True)
        c.set_fine_cull(0)
        c.set_map(w["keys"], w["map_rgb"], w["resolution"])
        c.set_camera(w["intr"], 1.0)
        bv0, cv0, hv0, _ = c.cast_views(w["pose_world"], w["init_pos"], mode=prv.MODE_VOXEL, want_hit_rank=True)
        assert np.array_equal(bv1, bv0) and np.array_equal(hv1, hv0)
    finally:
        c.close()
