"""ctypes loader for the CPU oracle (oracle/prv_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- may be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never from the product package.  Camera maths, View pose logic and the per-voxel
precept logic are pinned against the reference's own code compiled into oracle/_ref/ (ref_* below,
tests/test_oracle_kat.py); PARITY UNPINNED for OctoMap's castRay / key maths and Eigen's rounding (libraries absent; see
prv_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libprv_oracle.so")
NONE = 0xFFFFFFFF


class Intrinsics(C.Structure):
    """rs2_intrinsics layout, reference Share_Data.hpp:79-89."""
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("ppx", C.c_float), ("ppy", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("model", C.c_int), ("coeffs", C.c_float * 5)]


class CastStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("steps", C.c_uint64), ("probes_in", C.c_uint64), ("hits", C.c_uint64)]

    def as_dict(self):
        return {"rays": self.rays, "steps": self.steps, "probes_in": self.probes_in, "hits": self.hits}


POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1"), ("pad", "u1")])


def build(force=False):
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    src = os.path.join(_HERE, "prv_oracle.cpp")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "prv_oracle.h")))):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    d, f, u8, u16, u32, u64 = C.c_double, C.c_float, C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64
    P = C.POINTER
    L.orc_mat4_inverse.argtypes = [P(d), P(d)]
    L.orc_mat4_mul.argtypes = [P(d), P(d), P(d)]
    L.orc_project_point_to_pixel.argtypes = [P(f), P(Intrinsics), P(f)]
    L.orc_deproject_pixel_to_point.argtypes = [P(f), P(Intrinsics), P(f), f]
    L.orc_project_pixel_to_ray_end.argtypes = [C.c_int, C.c_int, P(Intrinsics), P(d), f, P(f)]
    L.orc_view_pose.argtypes = [P(d), P(d), P(d), P(d)]
    L.orc_view_pose_world.argtypes = [P(d), P(d), P(d)]
    L.orc_view_space.argtypes = [P(f), u64, P(d), C.c_int, d, d, P(d), P(d), P(d)]
    L.orc_view_space.restype = C.c_int
    L.orc_normalize_cloud.argtypes = [P(f), u64, d, P(d)]
    L.orc_coord_to_key.argtypes = [d, d, P(u16)]
    L.orc_coord_to_key.restype = C.c_int
    L.orc_key_to_coord.argtypes = [u16, d]
    L.orc_key_to_coord.restype = d
    L.orc_map_build.argtypes = [P(f), P(u8), u64, d]
    L.orc_map_build.restype = C.c_void_p
    L.orc_map_from_keys.argtypes = [P(u16), P(u8), u32, d]
    L.orc_map_from_keys.restype = C.c_void_p
    L.orc_map_free.argtypes = [C.c_void_p]
    L.orc_map_size.argtypes = [C.c_void_p]
    L.orc_map_size.restype = u32
    L.orc_map_keys.argtypes = [C.c_void_p, P(u16)]
    L.orc_map_rgb.argtypes = [C.c_void_p, P(u8)]
    L.orc_map_aabb.argtypes = [C.c_void_p, P(C.c_int), P(C.c_int)]
    L.orc_map_set_slow_lookup.argtypes = [C.c_void_p, C.c_int]
    L.orc_cast_ray.argtypes = [C.c_void_p, P(f), P(f), C.c_int, d, P(f), P(u32), P(CastStats)]
    L.orc_cast_ray.restype = C.c_int
    L.orc_precept.argtypes = [C.c_void_p, P(Intrinsics), P(d), P(d), d, C.c_void_p, P(u32), P(CastStats)]
    L.orc_precept.restype = C.c_int
    L.orc_precept_threads.argtypes = [C.c_void_p, P(Intrinsics), P(d), P(d), d, C.c_int, C.c_void_p, P(u32)]
    L.orc_precept_threads.restype = C.c_int
    L.orc_cast_view_dense.argtypes = [C.c_void_p, P(Intrinsics), P(d), P(d), d, P(u32), P(f), P(CastStats), C.c_int]
    L.orc_cast_view_dense.restype = C.c_int
    L.orc_bitset_words.argtypes = [u32]
    L.orc_bitset_words.restype = u32
    L.orc_bitset_from_ranks.argtypes = [P(u32), u64, P(u64), u32]
    L.orc_popcount_row.argtypes = [P(u64), u32]
    L.orc_popcount_row.restype = u32
    L.orc_greedy.argtypes = [P(u64), u32, u32, u32, u32, P(u32), P(u32), P(u64), P(u64)]
    L.orc_greedy.restype = u32
    L.orc_splat.argtypes = [P(f), P(u8), u64, P(Intrinsics), P(d), C.c_int, P(u8), P(f), P(u32)]
    L.orc_splat_focal.argtypes = [P(Intrinsics)]
    L.orc_splat_focal.restype = f
    L.orc_score_ensemble.argtypes = [P(u8), u32, u32, C.c_int, C.c_int, C.c_int, P(u8), P(d)]
    L.orc_score_ensemble.restype = C.c_int
    _lib = L
    return L


_REF_RS2_PATH = os.path.join(_HERE, "_ref", "librs2_ref.so")
_REF_HDR = "/root/reference/PRV_simulation/Share_Data.hpp"
_ref_rs2 = None


def build_ref(force=False):
    """oracle/_ref/*.so: the compilable parts of the reference (rs2_* camera functions, class View, precept_thread_process)
    built from where they lie (oracle/Makefile target `ref`).  Needs /root/reference; returns the path of librs2_ref.so, or
    None when neither the reference nor previously built libraries are there."""
    have_all = all(os.path.exists(os.path.join(_HERE, "_ref", f)) for f in ("librs2_ref.so", "libview_ref.so", "libprecept_ref.so"))
    if os.path.exists(_REF_HDR) and (force or not have_all):
        subprocess.run(["make", "-C", _HERE, "-B", "ref"], check=True, stdout=subprocess.DEVNULL)
    return _REF_RS2_PATH if os.path.exists(_REF_RS2_PATH) else None


def ref_rs2():
    """ctypes handle of oracle/_ref/librs2_ref.so (the real reference camera code), or None if it is not available."""
    global _ref_rs2
    if _ref_rs2 is None:
        path = build_ref()
        if path is None:
            return None
        L = C.CDLL(path)
        f = C.c_float
        L.ref_rs2_project_point_to_pixel.argtypes = [C.POINTER(f), C.POINTER(Intrinsics), C.POINTER(f)]
        L.ref_rs2_deproject_pixel_to_point.argtypes = [C.POINTER(f), C.POINTER(Intrinsics), C.POINTER(f), f]
        L.ref_rs2_sizeof_intrinsics.restype = C.c_int
        assert L.ref_rs2_sizeof_intrinsics() == C.sizeof(Intrinsics)
        _ref_rs2 = L
    return _ref_rs2


_REF_VIEW_PATH = os.path.join(_HERE, "_ref", "libview_ref.so")
_ref_view = None


def ref_view():
    """ctypes handle of oracle/_ref/libview_ref.so: the reference's own `class View` (View_Space.hpp:40-199) compiled
    against oracle/ref_eigen_shim.hpp, or None if it is not available."""
    global _ref_view
    if _ref_view is None:
        build_ref()
        if not os.path.exists(_REF_VIEW_PATH):
            return None
        L = C.CDLL(_REF_VIEW_PATH)
        d = C.c_double
        L.ref_view_get_next_camera_pos.argtypes = [C.POINTER(d), C.POINTER(d), C.POINTER(d), C.c_int, C.POINTER(d)]
        L.ref_view_space.argtypes = [C.POINTER(C.c_float), C.c_ulonglong, C.POINTER(d), C.c_int, d, d, C.POINTER(d), C.POINTER(d), C.POINTER(d)]
        L.ref_view_space.restype = C.c_int
        _ref_view = L
    return _ref_view


def ref_view_pose(init_pos, object_center, now_pose=None, type_of_pose=0):
    """View::get_next_camera_pos of the reference itself -> view.pose (4x4)."""
    now = _d16(np.eye(4) if now_pose is None else now_pose)
    ip = np.ascontiguousarray(init_pos, dtype=np.float64)
    oc = np.ascontiguousarray(object_center, dtype=np.float64)
    out = np.zeros(16)
    ref_view().ref_view_get_next_camera_pos(_ptr(now, C.c_double), _ptr(ip, C.c_double), _ptr(oc, C.c_double), int(type_of_pose), _ptr(out, C.c_double))
    return out.reshape(4, 4)


_REF_PRECEPT_PATH = os.path.join(_HERE, "_ref", "libprecept_ref.so")
_ref_precept = None


def ref_precept_lib():
    """ctypes handle of oracle/_ref/libprecept_ref.so: the reference's own Perception_3D::precept_thread_process
    (main.cpp:238-284) + project_pixel_to_ray_end (Share_Data.hpp:719-726) compiled against the Eigen / OctoMap / PCL
    shims (castRay & co. answered by this oracle), or None if it is not available."""
    global _ref_precept
    if _ref_precept is None:
        build_ref()
        if not os.path.exists(_REF_PRECEPT_PATH):
            return None
        lib()  # libprv_oracle.so must be loadable (the shim calls into it)
        L = C.CDLL(_REF_PRECEPT_PATH)
        L.ref_precept.argtypes = [C.c_void_p, C.c_double, C.POINTER(Intrinsics), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]
        L.ref_precept.restype = C.c_int
        _ref_precept = L
    return _ref_precept


def ref_view_space(points, sphere, view_space_radius, pt_norm):
    """View_Space::get_view_space of the reference itself -> (center[3], predicted_size, init_pos[nv,3])."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    sph = np.ascontiguousarray(sphere, dtype=np.float64)
    center = np.zeros(3)
    size = C.c_double(0)
    init = np.zeros((sph.shape[0], 3))
    nv = ref_view().ref_view_space(_ptr(pts, C.c_float), pts.shape[0], _ptr(sph, C.c_double), sph.shape[0], pt_norm, view_space_radius,
                                   _ptr(center, C.c_double), C.byref(size), _ptr(init, C.c_double))
    return center, size.value, init[:nv].copy()


def ref_project_point_to_pixel(intr, point):
    p = np.ascontiguousarray(point, dtype=np.float32)
    out = np.zeros(2, dtype=np.float32)
    ref_rs2().ref_rs2_project_point_to_pixel(_ptr(out, C.c_float), C.byref(intr), _ptr(p, C.c_float))
    return out


def ref_deproject_pixel_to_point(intr, pixel, depth):
    p = np.ascontiguousarray(pixel, dtype=np.float32)
    out = np.zeros(3, dtype=np.float32)
    ref_rs2().ref_rs2_deproject_pixel_to_point(_ptr(out, C.c_float), C.byref(intr), _ptr(p, C.c_float), C.c_float(depth))
    return out


def make_intrinsics(width, height, fx, fy, ppx, ppy, model=2, coeffs=(0, 0, 0, 0, 0)):
    it = Intrinsics()
    it.width, it.height = int(width), int(height)
    it.fx, it.fy, it.ppx, it.ppy = fx, fy, ppx, ppy
    it.model = int(model)
    for i in range(5):
        it.coeffs[i] = coeffs[i]
    return it


def _d16(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(16))


def mat4_inverse(m):
    out = np.zeros(16)
    lib().orc_mat4_inverse(_ptr(_d16(m), C.c_double), _ptr(out, C.c_double))
    return out.reshape(4, 4)


def project_point_to_pixel(intr, point):
    p = np.ascontiguousarray(point, dtype=np.float32)
    out = np.zeros(2, dtype=np.float32)
    lib().orc_project_point_to_pixel(_ptr(out, C.c_float), C.byref(intr), _ptr(p, C.c_float))
    return out


def deproject_pixel_to_point(intr, pixel, depth):
    p = np.ascontiguousarray(pixel, dtype=np.float32)
    out = np.zeros(3, dtype=np.float32)
    lib().orc_deproject_pixel_to_point(_ptr(out, C.c_float), C.byref(intr), _ptr(p, C.c_float), C.c_float(depth))
    return out


def project_pixel_to_ray_end(x, y, intr, pose_world, max_range=1.0):
    out = np.zeros(3, dtype=np.float32)
    pw = _d16(pose_world)
    lib().orc_project_pixel_to_ray_end(int(x), int(y), C.byref(intr), _ptr(pw, C.c_double), C.c_float(max_range), _ptr(out, C.c_float))
    return out


def view_pose(init_pos, object_center, now_pose=None):
    """View::get_next_camera_pos case 0 -> pose (4x4)."""
    now = _d16(np.eye(4) if now_pose is None else now_pose)
    ip = np.ascontiguousarray(init_pos, dtype=np.float64)
    oc = np.ascontiguousarray(object_center, dtype=np.float64)
    out = np.zeros(16)
    lib().orc_view_pose(_ptr(now, C.c_double), _ptr(ip, C.c_double), _ptr(oc, C.c_double), _ptr(out, C.c_double))
    return out.reshape(4, 4)


def view_pose_world(pose, now_pose=None):
    now = _d16(np.eye(4) if now_pose is None else now_pose)
    out = np.zeros(16)
    lib().orc_view_pose_world(_ptr(now, C.c_double), _ptr(_d16(pose), C.c_double), _ptr(out, C.c_double))
    return out.reshape(4, 4)


def view_space(points, sphere, view_space_radius, pt_norm=None):
    """View_Space::get_view_space -> (center[3], predicted_size, init_pos[nv,3])."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    sph = np.ascontiguousarray(sphere, dtype=np.float64)
    if pt_norm is None:
        pt_norm = float(np.sqrt(sph[0, 0] * sph[0, 0] + (sph[0, 1] * sph[0, 1] + sph[0, 2] * sph[0, 2])))
    center = np.zeros(3)
    size = C.c_double(0)
    init = np.zeros((sph.shape[0], 3))
    nv = lib().orc_view_space(_ptr(pts, C.c_float), pts.shape[0], _ptr(sph, C.c_double), sph.shape[0], pt_norm,
                              view_space_radius, _ptr(center, C.c_double), C.byref(size), _ptr(init, C.c_double))
    return center, size.value, init[:nv].copy()


def normalize_cloud(points, target_size):
    pts = np.array(points, dtype=np.float32, order="C", copy=True)
    before = C.c_double(0)
    lib().orc_normalize_cloud(_ptr(pts, C.c_float), pts.shape[0], target_size, C.byref(before))
    return pts, before.value


class Map:
    """ground_truth_model stand-in: occupied leaf keys in begin_leafs() (Morton) order + first-point colours."""

    def __init__(self, handle):
        self._h = handle
        n = lib().orc_map_size(handle)
        self.n = n
        self.keys = np.zeros((n, 3), dtype=np.uint16)
        self.rgb = np.zeros((n, 3), dtype=np.uint8)
        if n:
            lib().orc_map_keys(handle, _ptr(self.keys, C.c_uint16))
            lib().orc_map_rgb(handle, _ptr(self.rgb, C.c_uint8))
        lo = (C.c_int * 3)()
        hi = (C.c_int * 3)()
        lib().orc_map_aabb(handle, lo, hi)
        self.lo, self.hi = list(lo), list(hi)

    @classmethod
    def from_points(cls, points, rgb, resolution):
        pts = np.ascontiguousarray(points, dtype=np.float32)
        col = np.ascontiguousarray(rgb, dtype=np.uint8)
        m = cls(lib().orc_map_build(_ptr(pts, C.c_float), _ptr(col, C.c_uint8), pts.shape[0], resolution))
        m.resolution = resolution
        return m

    @classmethod
    def from_keys(cls, keys, rgb, resolution):
        k = np.ascontiguousarray(keys, dtype=np.uint16)
        col = None if rgb is None else np.ascontiguousarray(rgb, dtype=np.uint8)
        m = cls(lib().orc_map_from_keys(_ptr(k, C.c_uint16), None if col is None else _ptr(col, C.c_uint8), k.shape[0], resolution))
        m.resolution = resolution
        return m

    def set_slow_lookup(self, on):
        lib().orc_map_set_slow_lookup(self._h, 1 if on else 0)

    def __del__(self):
        try:
            if self._h:
                lib().orc_map_free(self._h)
                self._h = None
        except Exception:
            pass

    def cast_ray(self, origin, direction, ignore_unknown=True, max_range=1.0, stats=None):
        o = np.ascontiguousarray(origin, dtype=np.float32)
        dd = np.ascontiguousarray(direction, dtype=np.float32)
        end = np.zeros(3, dtype=np.float32)
        rank = C.c_uint32(NONE)
        found = lib().orc_cast_ray(self._h, _ptr(o, C.c_float), _ptr(dd, C.c_float), 1 if ignore_unknown else 0, max_range,
                                   _ptr(end, C.c_float), C.byref(rank), None if stats is None else C.byref(stats))
        return bool(found), end, rank.value

    def precept(self, intr, pose_world, init_pos, max_range=1.0, stats=None):
        out = np.zeros(self.n, dtype=POINT_DTYPE)
        ranks = np.zeros(self.n, dtype=np.uint32)
        pw = _d16(pose_world)
        ip = np.ascontiguousarray(init_pos, dtype=np.float64)
        ok = lib().orc_precept(self._h, C.byref(intr), _ptr(pw, C.c_double), _ptr(ip, C.c_double), max_range,
                               out.ctypes.data_as(C.c_void_p), _ptr(ranks, C.c_uint32), None if stats is None else C.byref(stats))
        return bool(ok), out, ranks

    def ref_precept(self, intr, pose_world, init_pos):
        """The REFERENCE's per-voxel method run for every voxel of this map (oracle/_ref/libprecept_ref.so)."""
        out = np.zeros(self.n, dtype=POINT_DTYPE)
        pw = _d16(pose_world)
        ip = np.ascontiguousarray(init_pos, dtype=np.float64)
        ok = ref_precept_lib().ref_precept(self._h, self.resolution, C.byref(intr), _ptr(pw, C.c_double), _ptr(ip, C.c_double),
                                           out.ctypes.data_as(C.c_void_p))
        return bool(ok), out

    def precept_threads(self, intr, pose_world, init_pos, max_range=1.0, num_of_thread=20):
        """precept with the reference's execution structure: one std::thread per voxel in batches of num_of_thread."""
        out = np.zeros(self.n, dtype=POINT_DTYPE)
        ranks = np.zeros(self.n, dtype=np.uint32)
        pw = _d16(pose_world)
        ip = np.ascontiguousarray(init_pos, dtype=np.float64)
        ok = lib().orc_precept_threads(self._h, C.byref(intr), _ptr(pw, C.c_double), _ptr(ip, C.c_double), max_range, num_of_thread,
                                       out.ctypes.data_as(C.c_void_p), _ptr(ranks, C.c_uint32))
        return bool(ok), out, ranks

    def cast_view_dense(self, intr, pose_world, init_pos, max_range=1.0, want_depth=True, stats=None, num_threads=0):
        ranks = np.zeros((intr.height, intr.width), dtype=np.uint32)
        depth = np.zeros((intr.height, intr.width), dtype=np.float32) if want_depth else None
        pw = _d16(pose_world)
        ip = np.ascontiguousarray(init_pos, dtype=np.float64)
        ok = lib().orc_cast_view_dense(self._h, C.byref(intr), _ptr(pw, C.c_double), _ptr(ip, C.c_double), max_range,
                                       _ptr(ranks, C.c_uint32), None if depth is None else _ptr(depth, C.c_float),
                                       None if stats is None else C.byref(stats), num_threads)
        return bool(ok), ranks, depth


def bitset_words(n_occ):
    return int(lib().orc_bitset_words(n_occ))


def bitset_from_ranks(ranks, words):
    r = np.ascontiguousarray(ranks, dtype=np.uint32).reshape(-1)
    row = np.zeros(words, dtype=np.uint64)
    lib().orc_bitset_from_ranks(_ptr(r, C.c_uint32), r.size, _ptr(row, C.c_uint64), words)
    return row


def greedy(vis, first_view, max_iter):
    vis = np.ascontiguousarray(vis, dtype=np.uint64)
    V, words = vis.shape
    seq = np.zeros(max_iter + 1, dtype=np.uint32)
    gains = np.zeros(max_iter + 1, dtype=np.uint32)
    covered = np.zeros(words, dtype=np.uint64)
    scored = C.c_uint64(0)
    n = lib().orc_greedy(_ptr(vis, C.c_uint64), V, words, first_view, max_iter, _ptr(seq, C.c_uint32), _ptr(gains, C.c_uint32),
                         _ptr(covered, C.c_uint64), C.byref(scored))
    return seq[:n].copy(), gains[:n].copy(), covered, scored.value


def splat(points, rgb, intr, pose_world, point_size=5):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    col = np.ascontiguousarray(rgb, dtype=np.uint8)
    H, W = intr.height, intr.width
    rgba = np.zeros((H, W, 4), dtype=np.uint8)
    depth = np.zeros((H, W), dtype=np.float32)
    index = np.zeros((H, W), dtype=np.uint32)
    pw = _d16(pose_world)
    lib().orc_splat(_ptr(pts, C.c_float), _ptr(col, C.c_uint8), pts.shape[0], C.byref(intr), _ptr(pw, C.c_double), point_size,
                    _ptr(rgba, C.c_uint8), _ptr(depth, C.c_float), _ptr(index, C.c_uint32))
    return rgba, depth, index


def splat_focal(intr):
    return float(lib().orc_splat_focal(C.byref(intr)))


def score_ensemble(images, method, chosen=None):
    """images: [V][E][H][W][4] uint8 -> (best view id, scores[V])."""
    im = np.ascontiguousarray(images, dtype=np.uint8)
    V, E, H, W, _ = im.shape
    scores = np.zeros(V)
    ch = None if chosen is None else np.ascontiguousarray(chosen, dtype=np.uint8)
    best = lib().orc_score_ensemble(_ptr(im, C.c_uint8), V, E, W, H, method, None if ch is None else _ptr(ch, C.c_uint8), _ptr(scores, C.c_double))
    return best, scores


class HostShim:
    """The names nerf_prv_b200/synth.py::build_workload needs from a host provider (normalisation, map insertion, view
    space, poses), answered by the oracle's own restatements instead of the product's host shim.  The reference arm of
    bench.py and the golden generators build their workloads through this, so neither needs libprv_b200.so."""
    make_intrinsics = staticmethod(make_intrinsics)
    host_normalize_cloud = staticmethod(normalize_cloud)
    host_view_space = staticmethod(view_space)

    @staticmethod
    def host_build_map(points, rgb, resolution):
        m = Map.from_points(points, rgb, resolution)
        return m.keys.copy(), m.rgb.copy()

    @staticmethod
    def view_poses(init_pos, object_center):
        out = np.zeros((len(init_pos), 4, 4))
        for v, ip in enumerate(init_pos):
            out[v] = view_pose_world(view_pose(ip, object_center))
        return out
