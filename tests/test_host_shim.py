"""Host-side logic of the product (pure host entry points of include/prv.h and the C++ mirror classes) against the
oracle, and the C-ABI surface itself.  No GPU needed."""
import os
import subprocess

import numpy as np
import pytest


def test_abi_exports_every_declared_symbol(prv):
    L = prv.lib()
    assert L.prv_abi_version() == 2
    out = subprocess.run(["nm", "-D", prv.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    declared = prv.declared_symbols()
    assert len(declared) >= 45
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    # the ctypes binding covers the whole header
    for s in declared:
        assert hasattr(L, s)


def test_no_cpu_fallback(prv):
    """Without a CUDA device the library must fail loudly, never compute on the CPU."""
    import ctypes as C
    try:
        import torch
        if torch.cuda.is_available():
            import pytest
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    h = C.c_void_p()
    rc = prv.lib().prv_create(C.byref(h), 0)
    assert rc == prv.ERR_NO_DEVICE and not h
    assert b"no CPU fallback" in prv.lib().prv_last_error(None)
    try:
        prv.Context(0)
        assert False
    except prv.PrvError as e:
        assert e.code == prv.ERR_NO_DEVICE


def test_product_does_not_reference_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for d, _, files in os.walk(os.path.join(root, "nerf-prv_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(d, f), errors="ignore").read()
                assert "prv_oracle" not in src and "import oracle" not in src and "orc_" not in src, os.path.join(d, f)


def test_mat4_inverse_and_poses_match_oracle(prv, orc):
    rng = np.random.default_rng(11)
    for _ in range(50):
        m = rng.normal(size=(4, 4))
        assert np.array_equal(prv.host_mat4_inverse(m), orc.mat4_inverse(m))
    c = np.array([3e-10, -1e-10, 2e-10])
    for _ in range(40):
        ip = rng.normal(size=3)
        ip[2] = abs(ip[2])
        ip = ip / np.linalg.norm(ip) * 0.3 + c
        p1, p2 = prv.host_view_pose(ip, c), orc.view_pose(ip, c)
        assert np.array_equal(p1, p2)
        assert np.array_equal(prv.host_view_pose_world(p1), orc.view_pose_world(p2))


def test_view_space_normalisation_and_map_match_oracle(prv, orc, synth):
    raw = synth.raw_surface("torus", 5, 4000)
    lat, rgb = synth.lattice_cloud(raw)
    a, sa = prv.host_normalize_cloud(lat, 0.10)
    b, sb = orc.normalize_cloud(lat, 0.10)
    assert np.array_equal(a, b) and sa == sb
    sphere = synth.hemisphere_set(32)
    c1, s1, ip1 = prv.host_view_space(a, sphere, 0.3)
    c2, s2, ip2 = orc.view_space(a, sphere, 0.3)
    assert np.array_equal(c1, c2) and s1 == s2 and np.array_equal(ip1, ip2)
    assert abs(s1 - 0.10) < 1e-6 and len(ip1) == 32
    k1, r1 = prv.host_build_map(a, rgb, 0.002)
    m = orc.Map.from_points(a, rgb, 0.002)
    assert np.array_equal(k1, m.keys) and np.array_equal(r1, m.rgb)
    # leaf order = ascending Morton code
    def morton(k):
        out = 0
        for bit in range(15, -1, -1):
            out = (out << 3) | (((int(k[2]) >> bit) & 1) << 2) | (((int(k[1]) >> bit) & 1) << 1) | ((int(k[0]) >> bit) & 1)
        return out
    codes = [morton(k) for k in k1]
    assert codes == sorted(codes) and len(set(codes)) == len(codes)


def test_hemisphere_sets(synth):
    for n in (3, 5, 32, 100):
        s = synth.hemisphere_set(n)
        assert s.shape == (n, 3) and np.all(s[:, 2] >= 0)
        assert np.allclose(np.linalg.norm(s, axis=1), 1.0, atol=1e-5)
        assert np.any(np.all(np.abs(s - [0, 0, 1]) < 1e-6, axis=1))  # every reference set contains the pole
    f = synth.hemisphere_set(1024)
    assert f.shape == (1024, 3) and f[0].tolist() == [0, 0, 1] and np.all(f[:, 2] > 0)
    assert np.allclose(np.linalg.norm(f, axis=1), 1.0, atol=1e-12)


def test_synthetic_workloads_are_deterministic(prv, synth):
    a = synth.build_workload(prv, "C1", n_views=4, size=(64, 48))
    b = synth.build_workload(prv, "C1", n_views=4, size=(64, 48))
    assert np.array_equal(a["keys"], b["keys"]) and np.array_equal(a["pose_world"], b["pose_world"])
    assert 15000 < len(a["keys"]) < 25000 and abs(a["predicted_size"] - 0.10) < 1e-6
    assert not np.any(np.all(a["cloud_rgb"] == 255, axis=1))


def test_leaf_order_check(prv, synth):
    """prv_host_check_leaf_order (what prv_set_map validates with; pdep fast path on BMI2 CPUs): agrees with a plain
    Morton-code comparison on a real key table, on duplicates, on swaps at every scale and on the empty table."""
    w = synth.build_workload(prv, "C1", n_views=1, size=(32, 24))
    keys = np.ascontiguousarray(w["keys"], dtype=np.uint16)
    n = len(keys)
    assert prv.host_check_leaf_order(keys) == n
    assert prv.host_check_leaf_order(keys[:0]) == 0 and prv.host_check_leaf_order(keys[:1]) == 1

    def code(k):
        c = 0
        for b in range(16):
            c |= ((int(k[0]) >> b) & 1) << (3 * b) | ((int(k[1]) >> b) & 1) << (3 * b + 1) | ((int(k[2]) >> b) & 1) << (3 * b + 2)
        return c

    def reference(kk):
        for i in range(1, len(kk)):
            if code(kk[i]) <= code(kk[i - 1]):
                return i
        return len(kk)

    rng = np.random.default_rng(5)
    for trial in range(60):
        kk = keys[: int(rng.integers(2, 400))].copy()
        i = int(rng.integers(0, len(kk) - 1))
        mode = trial % 3
        if mode == 0:
            kk[i + 1] = kk[i]                      # duplicate
        elif mode == 1:
            kk[[i, i + 1]] = kk[[i + 1, i]]        # neighbours swapped
        else:
            kk[i] = rng.integers(0, 65536, size=3)  # random key anywhere in the 16-bit cube
        assert prv.host_check_leaf_order(kk) == reference(kk), (trial, mode, i)
    # full-range keys: the top bits of every axis take part in the order
    big = rng.integers(0, 65536, size=(300, 3)).astype(np.uint16)
    order = np.argsort([code(k) for k in big], kind="stable")
    assert prv.host_check_leaf_order(big[order]) in (300, reference(big[order]))
    assert prv.host_check_leaf_order(big[order]) == reference(big[order])
    assert prv.host_check_leaf_order(big) == reference(big)


@pytest.mark.parametrize("model", [0, 1, 2, 3, 4, 5])
def test_host_camera_maths_equal_the_oracle_for_every_distortion_model(prv, orc, model):
    """prv_host_project_point_to_pixel / prv_host_deproject_pixel_to_point (what the library tabulates for models 3 / 5) against
    the oracle's rs2_* restatement, which tests/test_oracle_kat.py pins bit for bit to the reference's own compiled code."""
    rng = np.random.default_rng(300 + model)
    for coeffs in ((0.12042199820280075, -0.21373499929904938, -0.0021210000850260258, 0.0053860000334680080, 0.0), (0.31, -0.47, 0.013, -0.009, 0.12),
                   (0.9, 0, 0, 0, 0)):
        it = prv.make_intrinsics(1280, 720, 915.60669, 913.32666, 647.14532, 372.51532, model, coeffs)
        oit = orc.make_intrinsics(1280, 720, 915.60669, 913.32666, 647.14532, 372.51532, model, coeffs)
        for _ in range(300):
            pt = np.array([rng.normal() * 0.2, rng.normal() * 0.2, 0.05 + rng.random()], dtype=np.float32)
            assert prv.host_project_point_to_pixel(it, pt).tobytes() == orc.project_point_to_pixel(oit, pt).tobytes()
            px = np.floor(np.array([rng.random() * 1281, rng.random() * 721], dtype=np.float32))
            if model == 1:
                with pytest.raises(prv.PrvError):
                    prv.host_deproject_pixel_to_point(it, px)
            else:
                assert prv.host_deproject_pixel_to_point(it, px, 1.0).tobytes() == orc.deproject_pixel_to_point(oit, px, 1.0).tobytes()
