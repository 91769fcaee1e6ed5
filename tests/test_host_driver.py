"""The C++ host driver (host/pose): reference CLI, dataset readers, PLY outputs, on top of the C-ABI."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from online_3d_reconstruction_b200 import abi, synth, tmat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POSE = os.path.join(ROOT, "host", "pose")


@pytest.fixture(scope="module")
def pose_bin():
    if not os.path.exists(POSE):
        import __graft_entry__ as g
        g.build()
    return POSE


def write_png(path, img):
    """Minimal PNG writer (8-bit gray or RGB, filter type 1 on odd rows to exercise the reader's unfiltering)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else 3
    rows = img.reshape(h, w * ch).astype(np.int16)
    raw = bytearray()
    for y in range(h):
        if y % 2:   # Sub filter
            d = rows[y].copy()
            d[ch:] -= rows[y][:-ch]
            raw += b"\x01" + (d & 255).astype(np.uint8).tobytes()
        else:
            raw += b"\x00" + rows[y].astype(np.uint8).tobytes()

    def chunk(t, d):
        c = struct.pack(">I", len(d)) + t + d
        return c + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    ihdr = struct.pack(">IIBBBBB", w, h, 8, 0 if ch == 1 else 2, 0, 0, 0)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(bytes(raw))) + chunk(b"IEND", b""))


def read_ply(path):
    raw = open(path, "rb").read()
    h = raw.index(b"end_header\n") + len(b"end_header\n")
    n = int([l for l in raw[:h].decode().splitlines() if l.startswith("element vertex")][0].split()[-1])
    v = np.frombuffer(raw[h:h + 15 * n], dtype=np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1")]))
    pts = np.zeros(n, dtype=abi.POINT)
    pts["x"], pts["y"], pts["z"] = v["x"], v["y"], v["z"]
    pts["rgb"] = (v["r"].astype(np.uint32) << 16) | (v["g"].astype(np.uint32) << 8) | v["b"]
    return pts, raw[:h], raw[h + 15 * n:]


def test_cli_usage_and_reference_error_messages(pose_bin, tmp_path):
    # (the driver creates output/<timestamp>/ under its working directory, like the reference: keep that out of the repo)
    r = subprocess.run([pose_bin], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and "--seq_len" in r.stdout and "--dont_downsample" in r.stdout
    r = subprocess.run([pose_bin, "1", "2", "--seq_len", "0"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "Exception: invalid seq_len value!" in r.stdout       # pose_functions.cpp:199-200
    r = subprocess.run([pose_bin, "1", "2", "--voxel_size", "0.05"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "seq_len" in r.stdout                                   # pose.h:97
    r = subprocess.run([pose_bin, "--visualize", "x.ply"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 2 and "host-side tool" in r.stdout
    # every flag north_star lists is parsed and echoed the way the reference echoes it
    r = subprocess.run([pose_bin, "5", "6", "--seq_len", "50", "--voxel_size", "0.05", "--jump_pixels", "15",
                        "--range_width", "30", "--dist_nearby", "2", "--min_points_per_voxel", "1", "--blur_kernel", "1",
                        "--dont_downsample", "--data_root", "/nonexistent"], capture_output=True, text=True, cwd=tmp_path)
    for s in ("seq_len 50", "voxel_size 0.05", "jump_pixels 15", "range_width 30", "dist_nearby 2",
              "min_points_per_voxel 1", "blur_kernel 1", "dont_downsample"):
        assert s in r.stdout, s
    assert r.returncode == 1   # no such data set


def _write_dataset(root, n, rows, cols, seed=77, labels=None):
    os.makedirs(os.path.join(root, "data_files")); os.makedirs(os.path.join(root, "images"))
    os.makedirs(os.path.join(root, "disparities"))
    if labels is not None:
        os.makedirs(os.path.join(root, "segmentlabels"))
    rng = np.random.default_rng(seed)
    Q = abi.Q_CAM13
    with open(os.path.join(root, "data_files", "cam13calib.yml"), "w") as f:
        f.write("%YAML:1.0\nM1: !!opencv-matrix\n   rows: 1\n   cols: 1\n   dt: d\n   data: [ 1. ]\n"
                "Q: !!opencv-matrix\n   rows: 4\n   cols: 4\n   dt: d\n   data: [ " +
                ",\n       ".join(", ".join(repr(float(v)) for v in Q[i:i + 3]) for i in range(0, 16, 3)) + " ]\n")
    seq = synth.sequence(seed, n, rows, cols)
    # flat-ish disparities so the reference's variance gate (getVariance > 5 -> rejected, pose.cpp:187-196) accepts
    # them; frame 2 gets a big step and must be rejected
    xx = np.arange(cols)[None, :]
    for i in range(n):
        d = np.clip(np.rint(105 + rng.normal(0, 1.2, (rows, cols)) + 1.5 * np.sin(xx / 40.0 + i)), 0, 127).astype(np.uint8)
        if i == 2:
            d[:, cols // 2:] += 20
        seq[i] = (d, seq[i][1], seq[i][2])
    poses, times = [], []
    for i, (d, img, _) in enumerate(seq):
        num = 100 + i
        write_png(os.path.join(root, "disparities", f"{num}.png"), d)
        if labels is not None:
            write_png(os.path.join(root, "segmentlabels", f"{num}.png"), labels)
        write_png(os.path.join(root, "images", f"{num}.png"), img[:, :, ::-1])   # file is RGB, imread gives BGR
        t_ns = 1532043429586341888 + i * 227000000
        times.append((num, t_ns))
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        q = np.array([0.01, -0.02, 0.05, -0.9985]) + 0.002 * q; q /= np.linalg.norm(q)
        poses.append((t_ns, 0.45 * i, 0.1 * i, 22.0, *q))
    with open(os.path.join(root, "data_files", "images.txt"), "w") as f:
        for num, t in times:
            f.write(f"{num}.000000,{t / 1e9:.6f},{t}.000000\n")
    # pose.txt is much longer than the image list (2142 vs 171 rows in the reference's data); the reference's
    # binarySearchUsingTime returns index 0 / last as soon as it probes them, so the frames sit in the interior
    with open(os.path.join(root, "data_files", "pose.txt"), "w") as f:
        t0 = poses[0][0]
        rows_out = [(t0 - (25 - k) * 227000000, 99.0, 99.0, 99.0, 0.0, 0.0, 0.0, 1.0) for k in range(25)] + poses + \
                   [(poses[-1][0] + (k + 1) * 227000000, 99.0, 99.0, 99.0, 0.0, 0.0, 0.0, 1.0) for k in range(25)]
        for k, (t, *p) in enumerate(rows_out):
            f.write(f"{900 + k},{t / 1e9:.6f},{t}," + ",".join(f"{v:.9f}" for v in p) + "\n")
    return seq, poses


@pytest.mark.gpu
def test_driver_end_to_end_matches_python_path(pose_bin, tmp_path):
    """./pose 100 105 --seq_len 4 ... on a synthetic data set: cloud.ply must equal the Python mirror's result
    bit for bit, log.txt must carry the reference's phase names."""
    from online_3d_reconstruction_b200.pose import Pose
    rows, cols, n = 128, 256, 6
    root = str(tmp_path / "data")
    seq, poses = _write_dataset(root, n, rows, cols)
    out = str(tmp_path / "out")
    r = subprocess.run([pose_bin, "100", "105", "--seq_len", "4", "--voxel_size", "0.05", "--jump_pixels", "1",
                        "--min_points_per_voxel", "1", "--only_MAVLink", "--data_root", root, "--output", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    run_dir = os.path.join(out, sorted(os.listdir(out))[0])
    got, header, trailer = read_ply(os.path.join(run_dir, "cloud.ply"))
    log = open(os.path.join(run_dir, "log.txt")).read()
    for s in ("Cycle 0", "Cycle 1", "Point Cloud Creation time:", "Cycle time:", "disp_img_var", "Finished Pose Estimation"):
        assert s in log, s
    # the same frames through the Python mirror; poses as the driver parses them from pose.txt (9 decimals)
    p = abi.make_params(rows=rows, cols=cols, jump_pixels=1, voxel_size=0.05, min_points_per_voxel=1, sor_mean_k=50)
    keep = []
    frames = []
    for i, ((d, img, _), (t, *pp)) in enumerate(zip(seq, poses)):
        if i == 2:
            continue   # rejected by the variance gate
        pv = [float(f"{v:.9f}") for v in pp]
        frames.append(abi.make_frame(d, img, tmat.generate_tmat(*pv), keep=keep))
    assert "102  disp_img_var" in r.stdout and "> 5.\tRejected!" in r.stdout
    with Pose(p) as P:
        P.createCycleClouds(frames[:4])   # seq_len 4 accepted frames per cycle
        P.createCycleClouds(frames[4:])
        exp = P.downsamplePtCloud()
    assert len(exp) > 100
    assert np.array_equal(got, exp)
    uav, _, _ = read_ply(os.path.join(run_dir, "cloud_uavpos.ply"))
    assert len(uav) == 2 * (n - 1) and set(np.unique(uav["rgb"])) == {0x00FF00, 0xFF0000}


@pytest.mark.gpu
def test_driver_use_segment_labels_runs_plane_fit_on_gpu(pose_bin, tmp_path):
    """--use_segment_labels: the driver fits the per-label planes (createPlaneFittedDisparityImages, pose_functions.cpp:
    900-985) and gates frames through the GPU pre-pass, then scans the plane-fitted disparity.  cloud.ply must equal the
    Python mirror fed with the same labels and the GPU-fitted coefficients."""
    from online_3d_reconstruction_b200.pose import Pose
    rows, cols, n = 128, 256, 4
    labels = np.ones((rows, cols), np.uint8)
    labels[:, 130:200] = 2
    labels[:, 200:] = 3
    labels[60:63, 100:104] = 0   # a few unlabelled pixels: 0.0 -> masked out by disp > 64 (more than ~1 % of them would
                                 # trip the reference's plane_fitted_disp_img_var > 3 gate, whose mean divides by ALL pixels)
    root = str(tmp_path / "data")
    seq, poses = _write_dataset(root, n, rows, cols, seed=78, labels=labels)
    out = str(tmp_path / "out")
    r = subprocess.run([pose_bin, "100", "103", "--seq_len", "4", "--voxel_size", "0.05", "--jump_pixels", "1",
                        "--min_points_per_voxel", "1", "--only_MAVLink", "--use_segment_labels", "--data_root", root,
                        "--output", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    run_dir = os.path.join(out, sorted(os.listdir(out))[0])
    got, _, _ = read_ply(os.path.join(run_dir, "cloud.ply"))
    assert "plane_fitted_disp_img_var" in open(os.path.join(run_dir, "log.txt")).read()
    p = abi.make_params(rows=rows, cols=cols, jump_pixels=1, voxel_size=0.05, min_points_per_voxel=1, use_segment_labels=True,
                        sor_mean_k=50)   # the driver applies the reference's per-frame SOR (jump_pixels > 0)
    keep, frames = [], []
    with Pose(p) as P:
        for i, ((d, img, _), (t, *pp)) in enumerate(zip(seq, poses)):
            if i == 2:
                continue   # rejected by the variance gate
            coef, var = P.createPlaneFittedDisparityImages(labels, d)
            assert var <= 3
            pv = [float(f"{v:.9f}") for v in pp]
            frames.append(abi.make_frame(None, img, tmat.generate_tmat(*pv), labels=labels, plane_coef=coef, keep=keep))
        P.createCycleClouds(frames, abi.DISP_F64)
        exp = P.downsamplePtCloud()
    assert len(exp) > 100
    assert np.array_equal(got, exp)


@pytest.mark.gpu
def test_downsample_tool_reproduces_reference_cloud_ply_bytes(pose_bin, tmp_path, golden_dir):
    """pose --downsample <ply> (pose.cpp:71-87) on the reference's own build/cloud.ply: the combined VoxelGrid is
    idempotent on it, so the written file must be byte-identical to the reference's (header, vertices, camera)."""
    g = np.load(os.path.join(golden_dir, "cloud_ply.npz"))
    v = np.zeros(len(g["xyz"]), dtype=np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1")]))
    v["x"], v["y"], v["z"] = g["xyz"].T
    v["r"], v["g"], v["b"] = g["rgb"].T
    ref_bytes = g["header"].tobytes() + v.tobytes() + g["trailer"].tobytes()
    src = tmp_path / "cloud.ply"
    src.write_bytes(ref_bytes)
    r = subprocess.run([pose_bin, "--downsample", str(src), "--voxel_size", "0.05", "--output", str(tmp_path / "o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = (tmp_path / "downsampled_cloud.ply").read_bytes()
    assert out == ref_bytes
