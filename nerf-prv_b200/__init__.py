"""nerf-prv_b200 -- Python host binding (ctypes) over libprv_b200.so's C ABI (include/prv.h).

The directory name carries a hyphen, so load it with `load_pkg.load()` (repo root) or importlib;
inside Python the module is called `nerf_prv_b200`.

This module is plumbing for tests and bench.py: every compute call goes through the C ABI into the
sm_100a kernels.  There is no CPU fallback: if the shared library is missing or no B200 is present,
`Context()` raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PRV_B200_LIB") or os.path.join(_HERE, "libprv_b200.so")  # (PRV_B200_LIB: A/B builds of tools/ab_build.sh)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "prv.h")

NONE = 0xFFFFFFFF
MODE_VOXEL, MODE_DENSE = 0, 1
VARIANT_PLAIN, VARIANT_FAST, VARIANT_AXIS = 0, 1, 2

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_OOM, ERR_UNSUPPORTED, ERR_NCCL, ERR_IO = 0, -1, -2, -3, -4, -5, -6, -7


class PrvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("prv error %d: %s" % (code, msg))
        self.code = code


class Intrinsics(C.Structure):
    """prv_intrinsics == rs2_intrinsics (reference Share_Data.hpp:79-89)."""
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("ppx", C.c_float), ("ppy", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("model", C.c_int), ("coeffs", C.c_float * 5)]


class CastStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("probes_in", C.c_uint64), ("hits", C.c_uint64), ("steps", C.c_uint64), ("marched", C.c_uint64)]

    def as_dict(self):
        return {"rays": self.rays, "probes_in": self.probes_in, "hits": self.hits, "steps": self.steps, "marched": self.marched}


class Timing(C.Structure):
    _fields_ = [("cast_ms", C.c_float), ("cast_launches", C.c_uint32), ("cull_ms", C.c_float), ("cull_launches", C.c_uint32),
                ("march_ms", C.c_float), ("march_launches", C.c_uint32), ("project_ms", C.c_float), ("project_launches", C.c_uint32),
                ("count_ms", C.c_float), ("count_launches", C.c_uint32), ("greedy_ms", C.c_float), ("greedy_launches", C.c_uint32),
                ("splat_ms", C.c_float), ("splat_launches", C.c_uint32), ("resolve_ms", C.c_float), ("resolve_launches", C.c_uint32),
                ("other_ms", C.c_float), ("other_launches", C.c_uint32), ("gather_ms", C.c_float), ("gather_launches", C.c_uint32),
                ("flush_ms", C.c_float), ("dropped", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# pcl::PointXYZRGB memory image (32 bytes)
POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("w", "<f4"), ("b", "u1"), ("g", "u1"), ("r", "u1"), ("a", "u1"),
                        ("pad", "<f4", (3,))])
assert POINT_DTYPE.itemsize == 32

_lib = None


def build(force=False, verbose=False):
    """Compile libprv_b200.so in-tree (nvcc, sm_100a)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_prv_build", os.path.join(_HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force, verbose=verbose)


def lib():
    """Load the C-ABI library.  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PrvError(ERR_NO_DEVICE, "libprv_b200.so not built (run `python nerf-prv_b200/build.py`); no CPU fallback exists")
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    d, f, u8, u16, u32, u64, vp, i = C.c_double, C.c_float, C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64, C.c_void_p, C.c_int
    sig = {
        "prv_abi_version": (i, []),
        "prv_create": (i, [P(vp), i]),
        "prv_destroy": (None, [vp]),
        "prv_last_error": (C.c_char_p, [vp]),
        "prv_device_info": (i, [vp, P(i), P(i), P(i), P(u64)]),
        "prv_sync": (i, [vp]),
        "prv_set_variant": (i, [vp, i]),
        "prv_set_brick_cull": (i, [vp, i, i]),
        "prv_set_staging": (i, [vp, i, i]),
        "prv_host_mat4_inverse": (i, [P(d), P(d)]),
        "prv_host_view_pose": (i, [P(d), P(d), P(d), P(d)]),
        "prv_host_view_pose_world": (i, [P(d), P(d), P(d)]),
        "prv_host_view_space": (i, [P(f), u64, P(d), i, d, d, P(d), P(d), P(d), P(i)]),
        "prv_host_normalize_cloud": (i, [P(f), u64, d, P(d)]),
        "prv_host_build_map": (i, [P(f), P(u8), u64, d, P(u16), P(u8), P(u32)]),
        "prv_host_project_point_to_pixel": (i, [P(Intrinsics), P(f), P(f)]),
        "prv_host_deproject_pixel_to_point": (i, [P(Intrinsics), P(f), f, P(f)]),
        "prv_host_check_leaf_order": (i, [P(u16), u32, P(u32)]),
        "prv_set_map": (i, [vp, P(u16), P(u8), u32, d]),
        "prv_set_map_from_cloud": (i, [vp, P(f), P(u8), u64, d]),
        "prv_get_map": (i, [vp, P(u16), P(u8)]),
        "prv_set_camera": (i, [vp, P(Intrinsics), d]),
        "prv_set_views": (i, [vp, P(d), P(d), u32]),
        "prv_set_view_ids": (i, [vp, P(u32), u32]),
        "prv_full_voxels": (u32, [vp]),
        "prv_bitset_words": (u32, [vp]),
        "prv_num_views": (u32, [vp]),
        "prv_cast_async": (i, [vp, i, i]),
        "prv_greedy_async": (i, [vp, u32, u32]),
        "prv_get_bitsets": (i, [vp, P(u64)]),
        "prv_get_coverage_counts": (i, [vp, P(u32)]),
        "prv_get_hit_rank": (i, [vp, u32, u32, P(u32)]),
        "prv_get_depth": (i, [vp, u32, u32, P(f)]),
        "prv_get_greedy": (i, [vp, P(u32), P(u32), u32, P(u32), P(u64)]),
        "prv_greedy_path": (i, [vp]),
        "prv_get_cast_stats": (i, [vp, P(CastStats)]),
        "prv_cast_views": (i, [vp, P(d), P(d), u32, i, P(u64), P(u32), P(u32), P(f)]),
        "prv_precept": (i, [vp, P(d), P(d), vp, P(i)]),
        "prv_greedy": (i, [vp, u32, u32, P(u32), P(u32), u32, P(u32)]),
        "prv_set_cloud": (i, [vp, P(f), P(u8), u64]),
        "prv_render_views": (i, [vp, P(d), u32, i, P(u8), P(f)]),
        "prv_render_async": (i, [vp, u32, i]),
        "prv_object_pixel_rate": (i, [vp, P(d), u32, i, P(d), P(u32)]),
        "prv_splat_focal": (f, [P(Intrinsics)]),
        "prv_score_ensemble": (i, [vp, P(u8), u32, u32, i, i, i, P(u8), P(d), P(C.c_int32)]),
        "prv_timing_reset": (i, [vp]),
        "prv_get_timing": (i, [vp, P(Timing)]),
        "prv_event_record": (i, [vp, i]),
        "prv_event_elapsed_ms": (i, [vp, i, i, P(f)]),
        "prv_flush_l2": (i, [vp]),
        "prv_reset_counters": (i, [vp]),
        "prv_get_counters": (i, [vp, P(u64), P(u64), P(u64)]),
        "prv_map_bytes": (i, [vp, P(u64)]),
        "prv_comm_unique_id": (i, [vp]),
        "prv_comm_init": (i, [vp, vp, i, i]),
        "prv_allgather_bitsets_async": (i, [vp]),
        "prv_get_gathered": (i, [vp, P(u64), P(u32), P(u32)]),
        "prv_comm_p2p_export": (i, [vp, vp, u64]),
        "prv_comm_p2p_import": (i, [vp, vp, i, i]),
        "prv_comm_p2p_close": (i, [vp]),
        "prv_comm_destroy": (i, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = None  # filled lazily by exported_symbols()


def declared_symbols():
    """Function names declared in include/prv.h."""
    import re
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(prv_[a-z0-9_]+)\s*\(", src)))


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def make_intrinsics(width, height, fx, fy, ppx, ppy, model=2, coeffs=(0, 0, 0, 0, 0)):
    it = Intrinsics()
    it.width, it.height = int(width), int(height)
    it.fx, it.fy, it.ppx, it.ppy = fx, fy, ppx, ppy
    it.model = int(model)
    for k in range(5):
        it.coeffs[k] = coeffs[k]
    return it


# ---------------------------------------------------------------- pure-host logic (no device needed)
def _d16(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(16))


def host_mat4_inverse(m):
    out = np.zeros(16)
    lib().prv_host_mat4_inverse(_p(_d16(m), C.c_double), _p(out, C.c_double))
    return out.reshape(4, 4)


def host_view_pose(init_pos, object_center, now_pose=None):
    """View::get_next_camera_pos(now_camera_pose_world, object_center_world, 0) -> View::pose."""
    now = _d16(np.eye(4) if now_pose is None else now_pose)
    ip = np.ascontiguousarray(init_pos, dtype=np.float64)
    oc = np.ascontiguousarray(object_center, dtype=np.float64)
    out = np.zeros(16)
    rc = lib().prv_host_view_pose(_p(now, C.c_double), _p(ip, C.c_double), _p(oc, C.c_double), _p(out, C.c_double))
    if rc:
        raise PrvError(rc, "prv_host_view_pose")
    return out.reshape(4, 4)


def host_view_pose_world(pose, now_pose=None):
    now = _d16(np.eye(4) if now_pose is None else now_pose)
    out = np.zeros(16)
    lib().prv_host_view_pose_world(_p(now, C.c_double), _p(_d16(pose), C.c_double), _p(out, C.c_double))
    return out.reshape(4, 4)


def host_view_space(points, sphere, view_space_radius, pt_norm=None):
    """View_Space::get_view_space -> (object_center_world, predicted_size, init_pos[nv,3])."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    sph = np.ascontiguousarray(sphere, dtype=np.float64)
    if pt_norm is None:  # Share_Data: pt_norm = norm of row 0
        pt_norm = float(np.sqrt(sph[0, 0] * sph[0, 0] + (sph[0, 1] * sph[0, 1] + sph[0, 2] * sph[0, 2])))
    center = np.zeros(3)
    size = C.c_double(0)
    init = np.zeros((sph.shape[0], 3))
    nv = C.c_int(0)
    rc = lib().prv_host_view_space(_p(pts, C.c_float), pts.shape[0], _p(sph, C.c_double), sph.shape[0], pt_norm, view_space_radius,
                                   _p(center, C.c_double), C.byref(size), _p(init, C.c_double), C.byref(nv))
    if rc:
        raise PrvError(rc, "prv_host_view_space")
    return center, size.value, init[:nv.value].copy()


def host_normalize_cloud(points, target_size):
    pts = np.array(points, dtype=np.float32, order="C", copy=True)
    before = C.c_double(0)
    rc = lib().prv_host_normalize_cloud(_p(pts, C.c_float), pts.shape[0], target_size, C.byref(before))
    if rc:
        raise PrvError(rc, "prv_host_normalize_cloud")
    return pts, before.value


def host_build_map(points, rgb, resolution):
    """ground_truth_model insertion (first point's colour wins) -> (keys[N,3] u16 leaf order, rgb[N,3])."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    col = np.ascontiguousarray(rgb, dtype=np.uint8)
    keys = np.zeros((pts.shape[0], 3), dtype=np.uint16)
    out_rgb = np.zeros((pts.shape[0], 3), dtype=np.uint8)
    n = C.c_uint32(0)
    rc = lib().prv_host_build_map(_p(pts, C.c_float), _p(col, C.c_uint8), pts.shape[0], resolution, _p(keys, C.c_uint16), _p(out_rgb, C.c_uint8),
                                  C.byref(n))
    if rc:
        raise PrvError(rc, "prv_host_build_map")
    return keys[:n.value].copy(), out_rgb[:n.value].copy()


def host_project_point_to_pixel(intr, point):
    p = np.ascontiguousarray(point, dtype=np.float32)
    out = np.zeros(2, dtype=np.float32)
    rc = lib().prv_host_project_point_to_pixel(C.byref(intr), _p(p, C.c_float), _p(out, C.c_float))
    if rc:
        raise PrvError(rc, "prv_host_project_point_to_pixel")
    return out


def host_deproject_pixel_to_point(intr, pixel, depth=1.0):
    p = np.ascontiguousarray(pixel, dtype=np.float32)
    out = np.zeros(3, dtype=np.float32)
    rc = lib().prv_host_deproject_pixel_to_point(C.byref(intr), _p(p, C.c_float), depth, _p(out, C.c_float))
    if rc:
        raise PrvError(rc, "prv_host_deproject_pixel_to_point")
    return out


def host_check_leaf_order(keys):
    """Index of the first key that is not above its predecessor in leaf (Morton) order, or len(keys) when in order."""
    k = np.ascontiguousarray(keys, dtype=np.uint16).reshape(-1, 3)
    bad = C.c_uint32(0)
    rc = lib().prv_host_check_leaf_order(_p(k, C.c_uint16), k.shape[0], C.byref(bad))
    if rc:
        raise PrvError(rc, "prv_host_check_leaf_order")
    return int(bad.value)


def view_poses(init_pos, object_center):
    """view_pose_world (main.cpp:108-109) for every view position, through the host shim."""
    out = np.zeros((len(init_pos), 4, 4))
    for v, ip in enumerate(init_pos):
        out[v] = host_view_pose_world(host_view_pose(ip, object_center))
    return out


def splat_focal(intr):
    return float(lib().prv_splat_focal(C.byref(intr)))


# ---------------------------------------------------------------- device context
class Context:
    """One prv_ctx (one GPU)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().prv_create(C.byref(self._h), device)
        if rc:
            raise PrvError(rc, lib().prv_last_error(None).decode())
        self.intr = None
        self._greedy_max_iter = 0

    def close(self):
        if self._h:
            lib().prv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise PrvError(rc, lib().prv_last_error(self._h).decode())

    def device_info(self):
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_uint64()
        self._chk(lib().prv_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "mem_bytes": mem.value}

    def sync(self):
        self._chk(lib().prv_sync(self._h))

    def set_variant(self, v):
        self._chk(lib().prv_set_variant(self._h, v))

    def set_brick_cull(self, cell=0, enter_at_brick=True):
        """Brick edge of the conservative cull (4, 8 or 16 voxels; 0 = by map size) and whether the exact march starts at the first set
        brick.  Applies to the next set_map; results are identical for every setting."""
        self._chk(lib().prv_set_brick_cull(self._h, cell, 1 if enter_at_brick else 0))

    def set_staging(self, bitmap_in_shared_memory=False, l2_persisting_window=False):
        self._chk(lib().prv_set_staging(self._h, 1 if bitmap_in_shared_memory else 0, 1 if l2_persisting_window else 0))

    def set_map(self, keys, rgb, resolution):
        k = np.ascontiguousarray(keys, dtype=np.uint16)
        col = None if rgb is None else np.ascontiguousarray(rgb, dtype=np.uint8)
        self._chk(lib().prv_set_map(self._h, _p(k, C.c_uint16), _p(col, C.c_uint8), k.shape[0], resolution))

    def set_map_from_cloud(self, xyz, rgb, resolution):
        """GPU ingest of the normalised cloud (same rule as host_build_map)."""
        p = np.ascontiguousarray(xyz, dtype=np.float32)
        c = None if rgb is None else np.ascontiguousarray(rgb, dtype=np.uint8)
        self._chk(lib().prv_set_map_from_cloud(self._h, _p(p, C.c_float), _p(c, C.c_uint8), p.shape[0], resolution))

    def get_map(self):
        n = self.full_voxels
        keys = np.zeros((n, 3), dtype=np.uint16)
        rgb = np.zeros((n, 3), dtype=np.uint8)
        self._chk(lib().prv_get_map(self._h, _p(keys, C.c_uint16), _p(rgb, C.c_uint8)))
        return keys, rgb

    def set_camera(self, intr, max_range=1.0):
        self.intr = intr
        self._chk(lib().prv_set_camera(self._h, C.byref(intr), max_range))

    def set_views(self, pose_world, init_pos, view_ids=None):
        pw = np.ascontiguousarray(np.asarray(pose_world, dtype=np.float64).reshape(-1, 16))
        ip = np.ascontiguousarray(np.asarray(init_pos, dtype=np.float64).reshape(-1, 3))
        if view_ids is not None:
            ids = np.ascontiguousarray(view_ids, dtype=np.uint32)
            self._chk(lib().prv_set_view_ids(self._h, _p(ids, C.c_uint32), ids.size))
        else:
            self._chk(lib().prv_set_view_ids(self._h, None, 0))
        self._chk(lib().prv_set_views(self._h, _p(pw, C.c_double), _p(ip, C.c_double), pw.shape[0]))

    @property
    def full_voxels(self):
        return lib().prv_full_voxels(self._h)

    @property
    def words(self):
        return lib().prv_bitset_words(self._h)

    @property
    def num_views(self):
        return lib().prv_num_views(self._h)

    def cast_async(self, mode=MODE_DENSE, want_pixels=False, publish=False):
        """publish: PRV_CAST_PUBLISH (multi-GPU peer-memory exchange fused into the count kernel)."""
        self._chk(lib().prv_cast_async(self._h, mode, (1 if want_pixels else 0) | (2 if publish else 0)))

    def greedy_async(self, first_view, max_iter):
        self._chk(lib().prv_greedy_async(self._h, first_view, max_iter))
        self._greedy_max_iter = int(max_iter)

    @property
    def greedy_path(self):
        return lib().prv_greedy_path(self._h)

    def get_bitsets(self):
        out = np.zeros((self.num_views, self.words), dtype=np.uint64)
        self._chk(lib().prv_get_bitsets(self._h, _p(out, C.c_uint64)))
        return out

    def get_coverage_counts(self):
        out = np.zeros(self.num_views, dtype=np.uint32)
        self._chk(lib().prv_get_coverage_counts(self._h, _p(out, C.c_uint32)))
        return out

    def get_hit_rank(self, mode, view_begin=0, view_count=None):
        if view_count is None:
            view_count = self.num_views - view_begin
        if mode == MODE_DENSE:
            out = np.zeros((view_count, self.intr.height, self.intr.width), dtype=np.uint32)
        else:
            out = np.zeros((view_count, self.full_voxels), dtype=np.uint32)
        self._chk(lib().prv_get_hit_rank(self._h, view_begin, view_count, _p(out, C.c_uint32)))
        return out

    def get_depth(self, view_begin=0, view_count=None):
        if view_count is None:
            view_count = self.num_views - view_begin
        out = np.zeros((view_count, self.intr.height, self.intr.width), dtype=np.float32)
        self._chk(lib().prv_get_depth(self._h, view_begin, view_count, _p(out, C.c_float)))
        return out

    def get_greedy(self, max_iter=None, want_covered=True):
        """Result of the last greedy_async.  The arrays are sized from the max_iter that call ran with (remembered here), not
        from the caller's argument, which is only kept for source compatibility."""
        cap = self._greedy_max_iter + 1
        seq = np.zeros(cap, dtype=np.uint32)
        gains = np.zeros(cap, dtype=np.uint32)
        n = C.c_uint32(0)
        cov = np.zeros(self.words, dtype=np.uint64) if want_covered else None
        self._chk(lib().prv_get_greedy(self._h, _p(seq, C.c_uint32), _p(gains, C.c_uint32), cap, C.byref(n), _p(cov, C.c_uint64)))
        return seq[:n.value].copy(), gains[:n.value].copy(), cov

    def get_cast_stats(self):
        st = CastStats()
        self._chk(lib().prv_get_cast_stats(self._h, C.byref(st)))
        return st.as_dict()

    # host-buffer one-call API
    def cast_views(self, pose_world, init_pos, mode=MODE_DENSE, want_bitsets=True, want_counts=True, want_hit_rank=False, want_depth=False,
                   out_bitsets=None, out_counts=None):
        """prv_cast_views.  out_bitsets / out_counts: caller-owned result buffers ([V][words] uint64, [V] uint32), e.g. pinned
        host memory, which the library then fills with a direct device->host copy."""
        pw = np.ascontiguousarray(np.asarray(pose_world, dtype=np.float64).reshape(-1, 16))
        ip = np.ascontiguousarray(np.asarray(init_pos, dtype=np.float64).reshape(-1, 3))
        V = pw.shape[0]
        words = self.words
        bits = None
        if want_bitsets:
            bits = out_bitsets if out_bitsets is not None else np.empty((V, words), dtype=np.uint64)
            if bits.shape != (V, words) or bits.dtype != np.uint64 or not bits.flags.c_contiguous:
                raise ValueError("out_bitsets must be a C-contiguous uint64 array of shape (%d, %d)" % (V, words))
        counts = None
        if want_counts:
            counts = out_counts if out_counts is not None else np.empty(V, dtype=np.uint32)
            if counts.shape != (V,) or counts.dtype != np.uint32 or not counts.flags.c_contiguous:
                raise ValueError("out_counts must be a C-contiguous uint32 array of shape (%d,)" % V)
        hit = None
        if want_hit_rank:
            hit = np.zeros((V, self.intr.height, self.intr.width) if mode == MODE_DENSE else (V, self.full_voxels), dtype=np.uint32)
        depth = np.zeros((V, self.intr.height, self.intr.width), dtype=np.float32) if (want_depth and mode == MODE_DENSE) else None
        self._chk(lib().prv_set_view_ids(self._h, None, 0))
        self._chk(lib().prv_cast_views(self._h, _p(pw, C.c_double), _p(ip, C.c_double), V, mode, _p(bits, C.c_uint64), _p(counts, C.c_uint32),
                                       _p(hit, C.c_uint32), _p(depth, C.c_float)))
        return bits, counts, hit, depth

    def precept(self, pose_world, init_pos):
        """Perception_3D::precept for one view -> (cloud->points image, view_in_map)."""
        pw = _d16(pose_world)
        ip = np.ascontiguousarray(init_pos, dtype=np.float64)
        out = np.zeros(self.full_voxels, dtype=POINT_DTYPE)
        ok = C.c_int(0)
        self._chk(lib().prv_precept(self._h, _p(pw, C.c_double), _p(ip, C.c_double), out.ctypes.data_as(C.c_void_p), C.byref(ok)))
        return out, bool(ok.value)

    def greedy(self, first_view, max_iter):
        seq = np.zeros(max_iter + 1, dtype=np.uint32)
        gains = np.zeros(max_iter + 1, dtype=np.uint32)
        n = C.c_uint32(0)
        self._chk(lib().prv_greedy(self._h, first_view, max_iter, _p(seq, C.c_uint32), _p(gains, C.c_uint32), max_iter + 1, C.byref(n)))
        self._greedy_max_iter = int(max_iter)
        return seq[:n.value].copy(), gains[:n.value].copy()

    def set_cloud(self, xyz, rgb):
        p = np.ascontiguousarray(xyz, dtype=np.float32)
        c = np.ascontiguousarray(rgb, dtype=np.uint8)
        self._chk(lib().prv_set_cloud(self._h, _p(p, C.c_float), _p(c, C.c_uint8), p.shape[0]))

    def render_views(self, pose_world, point_size=5, want_depth=True):
        pw = np.ascontiguousarray(np.asarray(pose_world, dtype=np.float64).reshape(-1, 16))
        V = pw.shape[0]
        H, W = self.intr.height, self.intr.width
        rgba = np.zeros((V, H, W, 4), dtype=np.uint8)
        depth = np.zeros((V, H, W), dtype=np.float32) if want_depth else None
        self._chk(lib().prv_render_views(self._h, _p(pw, C.c_double), V, point_size, _p(rgba, C.c_uint8), _p(depth, C.c_float)))
        return rgba, depth

    def object_pixel_rate(self, pose_world, point_size=5):
        """prv_object_pixel_rate -> (mean non-white pixel fraction, per-view counts)."""
        pw = np.ascontiguousarray(np.asarray(pose_world, dtype=np.float64).reshape(-1, 16))
        rate = C.c_double(0)
        counts = np.zeros(pw.shape[0], dtype=np.uint32)
        self._chk(lib().prv_object_pixel_rate(self._h, _p(pw, C.c_double), pw.shape[0], point_size, C.byref(rate), _p(counts, C.c_uint32)))
        return rate.value, counts

    def render_async(self, V, point_size=5):
        self._chk(lib().prv_render_async(self._h, V, point_size))

    def score_ensemble(self, images, method, chosen=None):
        """nbv_loop cases 2/3: images [V][E][H][W][4] uint8 -> (best view id, scores[V])."""
        im = np.ascontiguousarray(images, dtype=np.uint8)
        V, E, H, W, _ = im.shape
        scores = np.zeros(V)
        best = C.c_int32(-1)
        ch = None if chosen is None else np.ascontiguousarray(chosen, dtype=np.uint8)
        self._chk(lib().prv_score_ensemble(self._h, _p(im, C.c_uint8), V, E, W, H, method, _p(ch, C.c_uint8), _p(scores, C.c_double), C.byref(best)))
        return best.value, scores

    def timing_reset(self):
        self._chk(lib().prv_timing_reset(self._h))

    def get_timing(self):
        t = Timing()
        self._chk(lib().prv_get_timing(self._h, C.byref(t)))
        return t.as_dict()

    def event_record(self, slot):
        self._chk(lib().prv_event_record(self._h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float(0)
        self._chk(lib().prv_event_elapsed_ms(self._h, a, b, C.byref(ms)))
        return ms.value

    def reset_counters(self):
        self._chk(lib().prv_reset_counters(self._h))

    def get_counters(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._chk(lib().prv_get_counters(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"kernel_launches": a.value, "h2d_bytes": b.value, "d2h_bytes": c.value}

    def map_bytes(self):
        a = C.c_uint64()
        self._chk(lib().prv_map_bytes(self._h, C.byref(a)))
        return a.value

    def flush_l2(self):
        self._chk(lib().prv_flush_l2(self._h))

    def comm_init(self, unique_id_bytes, rank, nranks):
        buf = C.create_string_buffer(bytes(unique_id_bytes), 128)
        self._chk(lib().prv_comm_init(self._h, C.cast(buf, C.c_void_p), rank, nranks))

    def p2p_export(self, table_bytes_max=0):
        h = (C.c_char * 64)()
        self._chk(lib().prv_comm_p2p_export(self._h, h, table_bytes_max))
        return bytes(h)

    def p2p_import(self, handles, rank, nranks):
        buf = b"".join(handles)
        assert len(buf) == 64 * nranks
        self._chk(lib().prv_comm_p2p_import(self._h, buf, rank, nranks))

    def comm_destroy_p2p(self):
        """Unmap the peer arenas (the NCCL communicator, if any, is kept)."""
        self._chk(lib().prv_comm_p2p_close(self._h))

    def get_gathered(self):
        """(rows [nranks*V][words], view ids [nranks*V]) of the all-gathered coverage table."""
        n = C.c_uint32(0)
        self._chk(lib().prv_get_gathered(self._h, None, None, C.byref(n)))
        rows = np.zeros((n.value, self.words), dtype=np.uint64)
        ids = np.zeros(n.value, dtype=np.uint32)
        self._chk(lib().prv_get_gathered(self._h, _p(rows, C.c_uint64), _p(ids, C.c_uint32), C.byref(n)))
        return rows, ids

    def allgather_bitsets_async(self):
        self._chk(lib().prv_allgather_bitsets_async(self._h))


def comm_unique_id():
    buf = C.create_string_buffer(128)
    rc = lib().prv_comm_unique_id(C.cast(buf, C.c_void_p))
    if rc:
        raise PrvError(rc, lib().prv_last_error(None).decode())
    return buf.raw
