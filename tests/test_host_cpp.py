"""The C++ host mirror (Share_Data / View_Space / NBV_Net_Labeler / Perception_3D, PNG + JSON writers, the
prv_simulation driver): configuration and output-file contract of the reference."""
import json
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# The 49 keys of the reference's DefaultConfiguration.yaml (values for this test; same syntax quirks: '%YAML:1.0'
# header, 'key : value' with a space before the colon, quoted paths, scientific floats, '0.' float).
YAML = """%YAML:1.0
pre_path: "{pre}/"
model_path: "{models}/"
shape_net: "E:/HRL/ShapeNetCore.v2/"
orginalviews_path: "../../view_space/Tammes_sphere/"
viewspace_path: "{hemi}/"
instant_ngp_path: "D:/Software/instant-ngp/scripts/"
pvb_path: "D:/Networks_pytorch/ConvNeXt-V2/"
is_shape_net: 1
id_of_batch: -1
name_of_pcd: "LM5"
num_of_thread: 20
method_of_IG : 0
octomap_resolution: 0.00625
ground_truth_resolution: 0.002
coverage_view_num_max: {vmax}
coverage_view_num_add: 2
points_size_cloud: 5
n_steps: 2500
evaluate: 0
ensemble_num: 5
object_pixel_rate: 0.035
num_of_neighbors_with_self: 1
num_of_choose: 64
num_of_random_test: 10
num_of_max_iteration: 64
num_of_most_cover: 1
cost_on: 0
cost_rate: 1.0
visit_weight_type: 1
trunc_threshold: 10
approaching_threshold: 0.03
show: 0
num_of_views : {nviews}
num_of_novel_test_views : 100
ray_casting_aabb_scale : 1
view_space_radius : 0.3
color_width: {w}
color_height: {h}
color_fx: {fx}
color_fy: {fy}
color_ppx: {ppx}
color_ppy: {ppy}
color_model: 2
color_k1: 1.2042199820280075e-01
color_k2: -2.1373499929904938e-01
color_k3: 5.3860000334680080e-03
color_p1: -2.1210000850260258e-03
color_p2: 0.
depth_scale: 1.0000000474974513e-03
"""


@pytest.fixture(scope="module")
def host_check(tmp_path_factory):
    out = tmp_path_factory.mktemp("cpp") / "host_check"
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-O1", "-std=c++17", "-ffp-contract=off", "-o", str(out), os.path.join(ROOT, "tests", "cpp", "host_check.cpp"),
                    "-L" + os.path.join(ROOT, "nerf-prv_b200"), "-lprv_b200", "-lz", "-Wl,-rpath," + os.path.join(ROOT, "nerf-prv_b200")], check=True)
    return str(out)


def write_env(tmp, synth, w=1280, h=720, nviews=32, vmax=5, sets=(3, 5, 32, 100)):
    hemi = tmp / "Hemisphere"
    hemi.mkdir(exist_ok=True)
    fixture = json.load(open(os.path.join(ROOT, "tests", "golden", "hemisphere_sets.json")))["sets"]
    for n in sets:
        with open(hemi / ("%d.txt" % n), "w") as f:
            for row in fixture[str(n)]:
                f.write(" ".join(row) + "\n")
    (tmp / "models" / "ShapeNet").mkdir(parents=True, exist_ok=True)
    (tmp / "out").mkdir(exist_ok=True)
    f32 = np.float32
    cfg = tmp / "cfg.yaml"
    cfg.write_text(YAML.format(pre=tmp / "out", models=tmp / "models", hemi=hemi, w=w, h=h, nviews=nviews, vmax=vmax,
                               fx=repr(float(f32(915.60668945312500) * f32(w) / f32(1280))), fy=repr(float(f32(913.32666015625) * f32(w) / f32(1280))),
                               ppx=repr(float(f32(647.14532470703125) * f32(w) / f32(1280))), ppy=repr(float(f32(372.51531982421875) * f32(h) / f32(720)))))
    return cfg


def results(out):
    d = {}
    for line in out.splitlines():
        if line.startswith("RESULT "):
            k, _, v = line[7:].partition("=")
            d[k] = v
    return d


def test_share_data_parses_reference_yaml(host_check, tmp_path, synth):
    cfg = write_env(tmp_path, synth)
    r = results(subprocess.run([host_check, "yaml", str(cfg), "chair_01", "3"], capture_output=True, text=True, check=True).stdout)
    assert r["name_of_pcd"] == "chair_01" and r["num_of_thread"] == "20" and r["num_of_max_iteration"] == "64"
    assert r["gt_path"] == "%s/out/Coverage_images/ShapeNet_3/chair_01" % tmp_path   # Share_Data.hpp:482-498 (+ id_of_batch)
    assert r["save_path"] == "%s/out/Compare/ShapeNet_3/chair_01" % tmp_path
    assert float(r["ground_truth_resolution"]) == 0.002 and float(r["view_space_radius"]) == 0.3
    assert (r["width"], r["height"], r["model"]) == ("1280", "720", "2")
    f32 = np.float32
    assert f32(float(r["fx"])) == f32(915.606689453125) * f32(1280) / f32(1280) and f32(float(r["ppy"])) == f32(372.51531982421875) * f32(720) / f32(720)
    # name/index quirk: color_k3 -> coeffs[2], color_p1 -> coeffs[3], color_p2 -> coeffs[4]
    assert np.float32(float(r["c2"])) == np.float32(5.3860000334680080e-03) and np.float32(float(r["c3"])) == np.float32(-2.1210000850260258e-03)
    assert float(r["c4"]) == 0.0 and np.float32(float(r["depth_scale"])) == np.float32(1.0000000474974513e-03)
    assert r["pt_sphere"] == "32" and abs(float(r["pt_norm"]) - 1.0) < 1e-5 and r["coverage_view_num_add"] == "2"
    if os.path.exists("/root/reference/PRV_simulation/DefaultConfiguration.yaml"):
        r = results(subprocess.run([host_check, "yaml", "/root/reference/PRV_simulation/DefaultConfiguration.yaml"], capture_output=True, text=True,
                                   check=True).stdout)
        assert r["name_of_pcd"] == "LM5" and r["num_of_views"] == "540" and r["pre_path"] == "D:/Data/NeRF_coverage/"
        assert r["gt_path"] == "D:/Data/NeRF_coverage/Coverage_images/ShapeNet/LM5" and r["coverage_view_num_max"] == "50"


def test_view_space_and_poses_match_oracle(host_check, tmp_path, prv, orc, synth):
    cfg = write_env(tmp_path, synth)
    w = synth.build_workload(prv, "C1", n_views=32, size=(64, 48), n_points=3000)
    np.savetxt(tmp_path / "cloud.txt", w["cloud"], fmt="%.9g")
    cloud = np.loadtxt(tmp_path / "cloud.txt", dtype=np.float32)
    out = subprocess.run([host_check, "views", str(cfg), str(tmp_path / "cloud.txt")], capture_output=True, text=True, check=True).stdout
    lines = [l.split() for l in out.splitlines() if l.startswith("RESULT ")]
    center = np.array([float(x) for x in next(l for l in lines if l[1] == "center")[2:]])
    size = float(next(l for l in lines if l[1] == "size")[2])
    c2, s2, ip2 = orc.view_space(cloud, synth.hemisphere_set(32), 0.3)
    assert np.array_equal(center, c2) and size == s2
    views = [l for l in lines if l[1] == "view"]
    assert len(views) == 32
    for l in views:
        i = int(l[2])
        ip = np.array([float(x) for x in l[3:6]])
        pw = np.array([float(x) for x in l[6:22]]).reshape(4, 4)
        assert np.array_equal(ip, ip2[i])
        assert np.array_equal(pw, orc.view_pose_world(orc.view_pose(ip2[i], c2)))
    toward = {int(l[2]): np.array([float(x) for x in l[3:]]).reshape(3, 3) for l in lines if l[1] == "toward"}
    assert toward[0].tolist() == np.eye(3).tolist()
    assert toward[4].tolist() == [[1, 0, 0], [0, 0, 1], [0, 1, 0]] and toward[5].tolist() == [[1, 0, 0], [0, 0, 1], [0, -1, 0]]
    assert toward[2].tolist() == [[0, 0, 1], [0, 1, 0], [1, 0, 0]] and toward[3].tolist() == [[0, 0, 1], [0, 1, 0], [-1, 0, 0]]
    assert toward[1].tolist() == [[1, 0, 0], [0, 1, 0], [0, 0, -1]]


def read_png(path):
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w = 8, b"", None
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(typ + body)
        if typ == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert depth == 8 and ctype in (2, 6)
            ch = 4 if ctype == 6 else 3
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, w * ch + 1)
    assert np.all(raw[:, 0] == 0)
    return raw[:, 1:].reshape(h, w, ch)


def test_png_and_json_writers(host_check, tmp_path):
    subprocess.run([host_check, "png", str(tmp_path / "t.png"), "37", "21"], check=True)
    img = read_png(tmp_path / "t.png")
    assert img.shape == (21, 37, 4)
    yy, xx = np.mgrid[0:21, 0:37]
    assert np.array_equal(img[..., 0], ((xx * 7 + yy) & 255).astype(np.uint8)) and np.array_equal(img[..., 3], np.where((xx + yy) & 1, 255, 0))
    try:
        import cv2
        assert np.array_equal(cv2.imread(str(tmp_path / "t.png"), cv2.IMREAD_UNCHANGED)[..., [2, 1, 0, 3]], img)
    except ImportError:
        pass
    subprocess.run([host_check, "json", str(tmp_path / "t.json")], check=True)
    d = json.load(open(tmp_path / "t.json"))
    assert d["w"] == 1280 and d["aabb_scale"] == 1 and d["fl_x"] == 915.606689453125 and d["offset"] == [0.5, 0.5000001, 0.25]
    assert d["k1"] == 0.12042199820280075 and len(d["frames"]) == 2 and d["frames"][1]["file_path"] == "3/rgbaClip_1.png"
    assert d["frames"][0]["transform_matrix"][1] == [0.2, 1.0, 0.22, 0.23]
    keys = [l.split('"')[1] for l in open(tmp_path / "t.json") if l.startswith('   "')]
    assert keys == sorted(keys)  # Json::Value member order


def write_ply(path, xyz, rgb, binary):
    n = len(xyz)
    hdr = "ply\nformat %s 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n" % (
        "binary_little_endian" if binary else "ascii", n)
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if binary:
            rec = np.zeros(n, dtype=[("p", "<f4", 3), ("c", "u1", 3)])
            rec["p"], rec["c"] = xyz, rgb
            f.write(rec.tobytes())
        else:
            for p, c in zip(xyz, rgb):
                f.write(("%.9g %.9g %.9g %d %d %d\n" % (p[0], p[1], p[2], c[0], c[1], c[2])).encode())


@pytest.mark.gpu
@pytest.mark.parametrize("binary", [True, False])
def test_prv_simulation_mode3_end_to_end(tmp_path, prv, orc, synth, binary):
    """The drop-in driver: stdin protocol, size.txt reuse, N = 3,5 then 100 view sets, rgbaClip PNGs + transforms JSON,
    idempotent resume -- and every image / matrix / coverage number equals the Python-side C-ABI results."""
    W, H = 160, 120
    cfg = write_env(tmp_path, synth, w=W, h=H, nviews=32, vmax=5)
    raw = synth.raw_surface("torus", 77, 6000)
    lat, rgb = synth.lattice_cloud(raw)
    write_ply(tmp_path / "models" / "ShapeNet" / "obj_a.ply", lat, rgb, binary)
    gt = tmp_path / "out" / "Coverage_images" / "ShapeNet" / "obj_a"
    gt.mkdir(parents=True)
    (gt / "size.txt").write_text("0.1")
    drv = os.path.join(ROOT, "nerf-prv_b200", "prv_simulation")
    r = subprocess.run([drv, str(cfg)], input="3\nobj_a\n-1\n", capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for n in (3, 5, 100):
        assert (gt / ("%d.json" % n)).exists() and (gt / str(n) / ("rgbaClip_%d.png" % (n - 1))).exists()
    # python-side recomputation through the same C ABI
    cloud, _ = prv.host_normalize_cloud(lat, 0.1)
    keys, map_rgb = prv.host_build_map(cloud, rgb, 0.002)
    intr = synth.intrinsics_for(prv.make_intrinsics, W, H)
    n = 5
    center, psize, init_pos = prv.host_view_space(cloud, synth.hemisphere_set(n), 0.3)
    pose_world = prv.view_poses(init_pos, center)
    ctx = prv.Context(0)
    ctx.set_camera(intr, 1.0)
    ctx.set_cloud(cloud, rgb)
    rgba, _ = ctx.render_views(pose_world, 5)
    d = json.load(open(gt / "5.json"))
    assert d["w"] == W and d["h"] == H and abs(d["scale"] - 0.5 / psize) < 1e-12 and d["aabb_scale"] == 1
    assert d["offset"] == [0.5 + center[2], 0.5 + center[0], 0.5 + center[1]]
    P = np.array([[0, 0, 1, 0], [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1.0]])
    P1 = np.diag([1.0, -1, -1, 1])
    for i in range(n):
        img = read_png(gt / "5" / ("rgbaClip_%d.png" % i))
        assert np.array_equal(img, rgba[i]), i
        assert d["frames"][i]["file_path"] == "5/rgbaClip_%d.png" % i
        assert np.allclose(np.array(d["frames"][i]["transform_matrix"]), P @ pose_world[i] @ P1, rtol=0, atol=1e-15)
    assert (rgba[..., 3] == 255).sum() > 100
    # coverage side file
    ctx.set_map(keys, map_rgb, 0.002)
    bits, counts, _, _ = ctx.cast_views(pose_world, init_pos, mode=prv.MODE_DENSE)
    seq, gains = ctx.greedy(0, 64)
    cov = dict(l.split(" ", 1) for l in open(gt / "5_coverage.txt").read().strip().splitlines())
    assert int(cov["full_voxels"]) == len(keys)
    assert [int(x) for x in cov["coverage_count"].split()] == counts.tolist()
    assert [int(x) for x in cov["greedy_seq"].split()] == seq.tolist() and [int(x) for x in cov["greedy_gain"].split()] == gains.tolist()
    ctx.close()
    # idempotent resume: nothing is rewritten when <N>.json exists
    stamp = {p: os.path.getmtime(p) for p in [gt / "3.json", gt / "5" / "rgbaClip_0.png", gt / "100.json"]}
    os.remove(gt / "5.json")
    r = subprocess.run([drv, str(cfg)], input="3\nobj_a\n-1\n", capture_output=True, text=True)
    assert r.returncode == 0
    assert (gt / "5.json").exists() and os.path.getmtime(gt / "3.json") == stamp[gt / "3.json"] and os.path.getmtime(gt / "100.json") == stamp[gt / "100.json"]
    # other modes are refused, unknown objects fail cleanly
    assert subprocess.run([drv, str(cfg)], input="21\nobj_a\n-1\n", capture_output=True, text=True).returncode == 2
    assert subprocess.run([drv, str(cfg)], input="3\nmissing_obj\n-1\n", capture_output=True, text=True).returncode == 3


def test_yaml_values_equal_what_opencv_filestorage_reads(host_check, tmp_path, synth):
    """The reference reads DefaultConfiguration.yaml with cv::FileStorage (Share_Data.hpp:339-401).  OpenCV's own parser is
    importable here (cv2, used as a checker only): every value the host mirror's parser produces must equal what
    cv::FileStorage returns for the same key, narrowed to the type of the reference's field (float intrinsics, double
    resolutions, int counters) -- on the test file with the reference's syntax quirks and on the shipped file when present."""
    cv2 = pytest.importorskip("cv2")
    files = [str(write_env(tmp_path, synth))]
    if os.path.exists("/root/reference/PRV_simulation/DefaultConfiguration.yaml"):
        files.append("/root/reference/PRV_simulation/DefaultConfiguration.yaml")
    # host_check name -> (yaml key, type of the reference's field)
    fields = {"pre_path": ("pre_path", str), "model_path": ("model_path", str), "viewspace_path": ("viewspace_path", str),
              "num_of_thread": ("num_of_thread", int), "ground_truth_resolution": ("ground_truth_resolution", float),
              "octomap_resolution": ("octomap_resolution", float), "coverage_view_num_max": ("coverage_view_num_max", int),
              "coverage_view_num_add": ("coverage_view_num_add", int), "points_size_cloud": ("points_size_cloud", int),
              "num_of_max_iteration": ("num_of_max_iteration", int), "view_space_radius": ("view_space_radius", float),
              "width": ("color_width", int), "height": ("color_height", int), "fx": ("color_fx", np.float32), "fy": ("color_fy", np.float32),
              "ppx": ("color_ppx", np.float32), "ppy": ("color_ppy", np.float32), "model": ("color_model", int),
              "c0": ("color_k1", np.float32), "c1": ("color_k2", np.float32), "c2": ("color_k3", np.float32), "c3": ("color_p1", np.float32),
              "c4": ("color_p2", np.float32), "depth_scale": ("depth_scale", float), "object_pixel_rate": ("object_pixel_rate", float),
              "is_shape_net": ("is_shape_net", int), "show": ("show", int), "ensemble_num": ("ensemble_num", int),
              "ray_casting_aabb_scale": ("ray_casting_aabb_scale", int), "n_steps": ("n_steps", int)}
    for path in files:
        r = results(subprocess.run([host_check, "yaml", path], capture_output=True, text=True, check=True).stdout)
        fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
        assert fs.isOpened()
        for name, (key, typ) in fields.items():
            node = fs.getNode(key)
            assert not node.isNone(), key
            if typ is str:
                assert r[name] == node.string(), (path, key)
            elif typ is int:
                assert int(r[name]) == int(node.real()), (path, key)
            elif typ is np.float32:
                assert np.float32(float(r[name])) == np.float32(node.real()), (path, key)
            else:
                assert float(r[name]) == float(node.real()), (path, key)
        fs.release()


@pytest.mark.gpu
def test_size_augmentation_probe_without_size_txt(tmp_path, prv, orc, synth):
    """SURVEY 8(f) #4 / main.cpp:851-964: without size.txt the labeler draws sizes in [last, 0.115) with the C library's
    rand() until the object fills more than object_pixel_rate of 5 test renders (<= 6 draws), and writes the accepted size
    (or -1).  The driver runs here with a pinned seed; the draws are recomputed with the same libc, the object rate of every
    draw with the ORACLE's splat (the device side is prv_object_pixel_rate: render + non-white count without a read-back),
    and the chosen size, the printed rates and the C-ABI counts must agree."""
    import ctypes
    W, H = 160, 120
    cfg = write_env(tmp_path, synth, w=W, h=H, nviews=32, vmax=3)
    raw = synth.raw_surface("torus", 78, 5000)
    lat, rgb = synth.lattice_cloud(raw)
    write_ply(tmp_path / "models" / "ShapeNet" / "obj_b.ply", lat, rgb, True)
    gt = tmp_path / "out" / "Coverage_images" / "ShapeNet" / "obj_b"
    drv = os.path.join(ROOT, "nerf-prv_b200", "prv_simulation")
    seed = 4242
    r = subprocess.run([drv, str(cfg), "--no-coverage"], input="3\nobj_b\n-1\n", capture_output=True, text=True, env=dict(os.environ, PRV_SIM_SEED=str(seed)))
    assert r.returncode == 0, r.stdout + r.stderr
    assert (gt / "size.txt").exists(), r.stdout
    chosen = float((gt / "size.txt").read_text())
    printed_sizes = [float(l.split()[-1]) for l in r.stdout.splitlines() if l.startswith("random size is")]
    printed_rates = [float(l.split()[-1]) for l in r.stdout.splitlines() if l.startswith("now object rate is")]
    assert 1 <= len(printed_sizes) == len(printed_rates) <= 6

    # the same draws with the same libc (View_Space.hpp:32-38 get_random_coordinate; RAND_MAX = 2^31 - 1 in glibc)
    libc = ctypes.CDLL(None)
    libc.srand(seed)
    RAND_MAX = 2147483647

    def draw(lo, hi):
        x = libc.rand() * (RAND_MAX + 1) + libc.rand()
        field = RAND_MAX * RAND_MAX + 2 * RAND_MAX
        return float(x) / float(field) * (hi - lo) + lo

    rate_threshold = float([l for l in open(cfg) if "object_pixel_rate" in l][0].split(":")[1])
    intr = synth.intrinsics_for(prv.make_intrinsics, W, H)
    oit = orc.make_intrinsics(intr.width, intr.height, intr.fx, intr.fy, intr.ppx, intr.ppy, intr.model, list(intr.coeffs))
    sphere5 = synth.hemisphere_set(5)
    ctx = prv.Context(0)
    size, expected, tests = 0.075, None, 0
    while True:
        size = draw(size, 0.115)
        assert abs(size - printed_sizes[tests]) <= 1e-6 * size  # (stdout prints 6 significant digits)
        cloud, _ = orc.normalize_cloud(lat, size)
        center = cloud.astype(np.float64).mean(axis=0)
        center = np.array([np.sum(cloud[:, a].astype(np.float64)) / len(cloud) for a in range(3)])
        init = sphere5 / np.linalg.norm(sphere5, axis=1, keepdims=True) * 0.3 + center
        pw = prv.view_poses(init, center)
        rate = 0.0
        counts = []
        for v in range(5):
            rgba, _, _ = orc.splat(cloud, rgb, oit, pw[v], 5)
            nonwhite = int(np.any(rgba[..., :3] != 255, axis=2).sum())
            counts.append(nonwhite)
            rate += nonwhite / float(W * H)
        rate /= 5
        ctx.set_camera(intr, 1.0)
        ctx.set_cloud(cloud, rgb)
        g_rate, g_counts = ctx.object_pixel_rate(pw, 5)
        assert g_counts.tolist() == counts and g_rate == rate, (tests, g_counts.tolist(), counts)
        assert abs(rate - printed_rates[tests]) <= 1e-5 * max(rate, 1e-9)
        tests += 1
        if not (rate <= rate_threshold and tests <= 5):
            expected = size if tests <= 5 else -1.0
            break
    ctx.close()
    assert tests == len(printed_sizes)
    assert abs(chosen - expected) <= 1e-6 * abs(expected), (chosen, expected)
    if expected > 0:
        assert (gt / "3.json").exists()  # the driver went on to generate the view sets with that size
