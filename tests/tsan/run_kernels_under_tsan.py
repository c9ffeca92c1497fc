"""Child process of tests/test_kernel_on_host.py::test_kernels_are_race_free_under_threadsanitizer: runs the cast, voxel-mode
and splat kernels on the SIMT emulator (tests/cpp/pipeline_on_host.cpp built with -fsanitize=thread, libtsan preloaded).
argv[1] = the instrumented library."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import load_pkg  # noqa: E402

prv = load_pkg.load()
from nerf_prv_b200 import synth  # noqa: E402
import test_kernel_on_host as T  # noqa: E402

poh = C.CDLL(sys.argv[1])
w = synth.build_workload(prv, "C1", n_views=3, size=(96, 72))
ref = None
for brick, entry in T.CONFIGS:
    out = T.run_kernels(poh, w, range(3), 1, brick, entry, grid=3)
    if ref is None:
        ref = out
    assert np.array_equal(out["hit"], ref["hit"]) and np.array_equal(out["bits"], ref["bits"])
w["init_pos"][1] = np.array([1.0e6, 0.0, 0.0])
T.run_kernels(poh, w, range(3), 1, 8, 0, max_range=0.3)          # literal march, a view out of the map
T.run_kernels(poh, w, range(2), 0, 4, 1)                          # voxel mode (masked region queue, gather)
xyz = np.ascontiguousarray(w["cloud"][::8], dtype=np.float32)
rgb = np.ascontiguousarray(w["cloud_rgb"][::8], dtype=np.uint8)
pw = np.ascontiguousarray(w["pose_world"][:2], dtype=np.float64)
it = w["intr"]
rgba = np.zeros((2, it.height, it.width, 4), dtype=np.uint8)
depth = np.zeros((2, it.height, it.width), dtype=np.float32)
assert poh.poh_render_views(T._p(xyz, C.c_float), T._p(rgb, C.c_uint8), C.c_uint64(len(xyz)), C.byref(it), T._p(pw, C.c_double), C.c_uint32(2), 5,
                            T._p(rgba, C.c_uint8), T._p(depth, C.c_float)) == 0
print("KERNELS-RAN-UNDER-TSAN hits", int((ref["hit"] != 0xFFFFFFFF).sum()))
