"""GPU tests for the corners the r01 review and the tile engine's window bound point at: state kept across calls,
abandoned prefetches, the u8 reprojection LUT against the generic path when Q[15] != 0, degenerate clouds, frames in arbitrary
orientations (the window bound of csrc/tile.cuh goes through the inverse of each frame's matrix), non-rigid matrices."""
import math

import numpy as np
import pytest

import oracle_binding as ob
from online_3d_reconstruction_b200 import abi, synth, tmat
from online_3d_reconstruction_b200.pose import Pose
from test_gpu_fused import FUSED, TILE, _multiset, _run_cycles
from test_gpu_parity import SMALL4, _close, _eq, _frames

pytestmark = pytest.mark.gpu


def test_frame_mask_between_a_cycle_and_its_consumers_leaves_the_batch_state_alone():
    """o3r_frame_mask runs the scan in mask-only mode; the last batch's per-frame clouds, its offsets and the pending exchange
    must be exactly what they were (the mask call used to clear them)."""
    keep = []
    geom = SMALL4
    p = abi.make_params(jump_pixels=1, voxel_size=0.05, merge_mode=abi.MERGE_ACCUMULATE_TILED, **geom)
    frames = _frames(600, 4, geom["rows"], geom["cols"], keep=keep)
    with Pose(p) as P, Pose(p) as Q:
        P.createCycleClouds(frames)
        before = P.lastCyclePoints()
        m = P.validityMask(frames[1])
        assert 0 < int(m.sum()) < m.size
        _eq(P.lastCyclePoints(), before)
        plain = P.downsamplePtCloud()
        # the same with the library's exchange between the cycle and the merge (world of one rank)
        Q.commInit(1, 0, Pose.commUniqueId(), 1 << 16)
        Q.createCycleClouds(frames)
        Q.validityMask(frames[2])
        Q.exchangeCycle()
        got = Q.downsamplePtCloud()
        Q.commDestroy()
    _close(got, plain)
    assert len(got) == len(plain) > 100


def test_cancelled_prefetch_never_touches_its_buffers_again():
    """A caller that abandons a prefetch cancels it and may then overwrite or free the buffers: the next cycle (other frames)
    is computed from its own inputs only (a deferred prefetch kept raw host pointers)."""
    keep = []
    geom = SMALL4
    rows, cols = geom["rows"], geom["cols"]
    p = abi.make_params(jump_pixels=1, voxel_size=0.05, **geom)
    seq = synth.sequence(610, 4, rows, cols)
    abandoned_d = [s[0].copy() for s in seq[:2]]
    abandoned = [abi.make_frame(abandoned_d[i], seq[i][1], seq[i][2], keep=keep) for i in range(2)]
    real = [abi.make_frame(seq[i][0], seq[i][1], seq[i][2], keep=keep) for i in (2, 3)]
    cloud, n, counts = ob.run_cycle(p, real, abi.DISP_U8, 2)
    with Pose(p) as P:
        P.prefetchCycle(abandoned)
        P.cancelPrefetch()
        for d in abandoned_d:
            d[:] = 0                       # the buffers are the caller's again
        got_counts = P.createCycleClouds(real)
        assert np.array_equal(got_counts, counts)
        _eq(P.lastCyclePoints(), cloud[:n])


def test_u8_lut_path_equals_generic_path_when_q15_is_not_zero():
    """Q[15] != 0 (a principal-point difference between the two cameras): the 256-entry reciprocal table of the u8 path and the
    per-pixel double arithmetic of the f32 path must give the same bits (the table was built with the host
    compiler's FP contraction)."""
    keep = []
    geom = SMALL4
    q = list(abi.Q_CAM13)
    q[15] = 0.37
    p = abi.make_params(jump_pixels=1, voxel_size=0.05, Q=tuple(q), **geom)
    d, img, T = synth.sequence(620, 1, geom["rows"], geom["cols"])[0]
    fr_u8 = abi.make_frame(d, img, T, keep=keep)
    fr_f32 = abi.make_frame(d.astype(np.float32), img, T, keep=keep)
    exp = ob.create_and_transform_pt_cloud(p, fr_u8, abi.DISP_U8)
    with Pose(p) as P:
        a = P.createAndTransformPtCloud(fr_u8, abi.DISP_U8)
        b = P.createAndTransformPtCloud(fr_f32, abi.DISP_F32)
    assert len(exp) > 1000
    _eq(a, exp)
    _eq(b, exp)


def test_voxel_grid_of_a_thin_cloud():
    """A cloud that is one line of points: two of the three grid extents are a single cell, the third is long (ADVICE r01:
    key_bits of degenerate grids)."""
    n = 60000
    rng = np.random.default_rng(630)
    pts = np.zeros(n, dtype=abi.POINT)
    pts["x"] = np.sort(rng.uniform(-40.0, 40.0, n)).astype(np.float32)
    pts["y"] = np.float32(3.25)
    pts["z"] = np.float32(-1.5)
    pts["rgb"] = rng.integers(0, 1 << 24, n, dtype=np.uint32)
    rng.shuffle(pts)
    for leaf in (0.01, 0.0005):
        exp_pts, exp_keys, exp_counts, exp_pass = ob.voxel_grid(pts, (leaf, leaf, leaf))
        p = abi.make_params(jump_pixels=1, voxel_size=0.05, **SMALL4)
        with Pose(p) as P:
            got, keys, counts, passthrough = P.voxelGrid(pts, (leaf, leaf, leaf))
        assert passthrough == exp_pass
        _eq(got, exp_pts)
        if not passthrough:
            assert np.array_equal(keys, exp_keys) and np.array_equal(counts, exp_counts) and int(counts.sum()) == n


def _quat(yaw, pitch, roll):
    cy, sy = math.cos(yaw / 2), math.sin(yaw / 2)
    cp, sp = math.cos(pitch / 2), math.sin(pitch / 2)
    cr, sr = math.cos(roll / 2), math.sin(roll / 2)
    return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)


def test_fused_tile_engine_frames_in_arbitrary_orientations():
    """The pixel window that holds a leaf's points is bounded through the inverse of each frame's matrix (tile.cuh): frames
    whose camera axes are nowhere near the world axes (yaw 37 / 123 / 211 / 300 degrees, up to 40 degrees of pitch and roll) make
    a leaf's footprint up to sqrt(3) wider than an axis-aligned one.  Per-frame voxels bit-exact, cells exact."""
    keep = []
    geom = SMALL4
    rows, cols = geom["rows"], geom["cols"]
    p = abi.make_params(jump_pixels=1, voxel_size=0.05, merge_mode=FUSED, **geom)
    seq = synth.sequence(640, 4, rows, cols)
    poses = [(37.0, 25.0, -10.0), (123.0, -40.0, 15.0), (211.0, 5.0, 40.0), (300.0, -30.0, -35.0)]
    frames = []
    for i, (yaw, pitch, roll) in enumerate(poses):
        T = tmat.generate_tmat(3.0 * i, -2.0 * i, 22.0, *_quat(math.radians(yaw), math.radians(pitch), math.radians(roll)))
        frames.append(abi.make_frame(seq[i][0], seq[i][1], T, keep=keep))
    got, exp = _run_cycles(p, [frames[:2], frames[2:]], expect_engine=TILE)
    assert len(exp) > 300
    _close(got, exp)


def test_fused_tile_engine_takes_a_scaled_and_sheared_matrix():
    """The matrix of a frame need not be rigid for the path (pcl::transformPointCloud takes any affine matrix): the window
    bound uses the true inverse, so a frame scaled by 0.6 / 1.7 along two axes and sheared still groups its leaves exactly."""
    keep = []
    geom = SMALL4
    p = abi.make_params(jump_pixels=1, voxel_size=0.05, merge_mode=FUSED, **geom)
    seq = synth.sequence(650, 2, geom["rows"], geom["cols"])
    frames = []
    for i in range(2):
        T = np.array(seq[i][2], dtype=np.float32).reshape(4, 4).copy()
        A = np.array([[0.6, 0.2, 0.0, 0.0], [0.0, 1.7, 0.1, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]], dtype=np.float32)
        frames.append(abi.make_frame(seq[i][0], seq[i][1], tmat.mat4_mul(A, T), keep=keep))
    with Pose(p) as P:
        P.setKeepFrameVoxels(True)
        cloud, n, counts = ob.run_cycle(p, frames, abi.DISP_U8, 2)
        got_counts = P.createCycleClouds(frames)
        assert P.lastCycleEngine() in (1, 2)   # (a wide enough window, or the bucket engine: both must be exact)
        assert np.array_equal(got_counts, counts)
        assert np.array_equal(_multiset(P.lastCyclePoints()), _multiset(cloud[:n]))
        got = P.downsamplePtCloud()
    _close(got, ob.downsample_pt_cloud(p, cloud[:n], True))


@pytest.mark.parametrize("disp_type,blur", [(abi.DISP_U8, 1), (abi.DISP_U16, 1), (abi.DISP_U8, 5)])
def test_adjacent_host_planes_move_as_grouped_copies(disp_type, blur):
    """Frames held back to back in one host arena are staged with ONE 3-D copy per plane type and group of <= 16 frames
    (host_frames.cuh::stage_copy; extent = the ROIs of the group's planes) — the staged ROI, and with it every result, must
    equal the per-plane copies of separately allocated frames: 21 frames = groups of 16 + 5 when prefetched, 10 + 10 + 1 when
    chunked; frame 7 lives in its own buffer and splits its group."""
    keep = []
    rows, cols = SMALL4["rows"], SMALL4["cols"]
    p = abi.make_params(jump_pixels=1, voxel_size=0.05, max_batch_frames=32, blur_kernel=blur, **SMALL4)
    seq = synth.sequence(520, 21, rows, cols, disp_type=disp_type)
    d_all = np.stack([d for d, _, _ in seq])
    c_all = np.stack([c for _, c, _ in seq])
    lone_d, lone_c = d_all[7].copy(), c_all[7].copy()
    arena = [abi.make_frame(lone_d if i == 7 else d_all[i], lone_c if i == 7 else c_all[i], T, keep=keep)
             for i, (_, _, T) in enumerate(seq)]
    assert arena[1].disp == arena[0].disp + rows * arena[0].disp_step and arena[8].disp != arena[7].disp + rows * arena[7].disp_step
    apart = [abi.make_frame(d.copy(), c.copy(), T, keep=keep) for d, c, T in seq]
    cloud, n, counts = ob.run_cycle(p, apart, disp_type, 4)
    exp = ob.downsample_pt_cloud(p, cloud[:n], True)
    for prefetch in (False, True):
        for frames in (arena, apart):
            with Pose(p) as P:
                arr = P.prefetchCycle(frames, disp_type) if prefetch else frames
                got_counts = P.createCycleClouds(arr, disp_type)
                assert np.array_equal(got_counts, counts)
                _eq(P.lastCyclePoints(), cloud[:n])
                _eq(P.downsamplePtCloud(), exp)
