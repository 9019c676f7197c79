"""CPU tests (-m "not gpu"): the oracle against the reference's golden vectors / real-reference fixtures,
property tests for the un-pinned third-party semantics, host logic, and the C-ABI export surface."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


# ----------------------------------------------------------------------------- golden vectors from the reference
def test_fk_matches_reference_known_answer(oracle):
    """interactive_demo/mpinets_ros/nodes/interaction_node.py:54-75"""
    g = np.load(os.path.join(HERE, "golden", "fk_reference.npz"))
    frames, eef = oracle.fk(g["q"][None].astype(np.float32))
    assert np.abs(eef[0, :, 3] - g["xyz"]).max() < 2e-7
    R = eef[0, :, :3].astype(np.float64)
    w = np.sqrt(max(0.0, 1 + np.trace(R))) / 2
    # quaternion is near a half-turn (w ~ 0.02): recover xyz from the symmetric part
    x, y, z, ww = g["xyzw"]
    Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * ww), 2 * (x * z + y * ww)],
                   [2 * (x * y + z * ww), 1 - 2 * (x * x + z * z), 2 * (y * z - x * ww)],
                   [2 * (x * z - y * ww), 2 * (y * z + x * ww), 1 - 2 * (x * x + y * y)]])
    assert np.abs(R - Rq).max() < 1e-6
    assert abs(abs(w) - abs(ww)) < 2e-5   # w ~ 0.02 from the trace is ill-conditioned in fp32


def test_fk_f64_helper_agrees(oracle):
    from mpinets_b200 import franka
    rng = np.random.RandomState(0)
    q = rng.uniform(franka.REAL_JOINT_LIMITS[:, 0], franka.REAL_JOINT_LIMITS[:, 1], size=(64, 7)).astype(np.float32)
    frames, eef = oracle.fk(q)
    for i in range(64):
        F64, g = franka.fk_reference_f64(q[i].astype(np.float64))
        assert np.abs(frames[i] - F64[:, :3]).max() < 5e-7
        assert np.abs(eef[i] - g[:3]).max() < 5e-7


@pytest.mark.parametrize("tag", ["yaw", "free"])
def test_sdf_matches_real_reference(oracle, tag):
    """tests/golden/sdf_reference.npz was produced by mpinets/geometry.py itself (make_golden.py)."""
    g = np.load(os.path.join(HERE, "golden", "sdf_reference.npz"))
    s = {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + "_")}
    for which, key in ((1, "sdf_cuboids"), (2, "sdf_cylinders")):
        mine = oracle.sdf_points(s, s["points"], quirk=True, which=which)
        ref = s[key]
        assert (np.isfinite(mine) == np.isfinite(ref)).all()       # all-masked scenes -> +inf (geometry.py:251-254)
        fin = np.isfinite(ref)
        assert np.abs(mine[fin] - ref[fin]).max() < 2e-6
    B, T, NS, _ = s["seq"].shape
    mine = oracle.sdf_points(s, s["seq"].reshape(B, T * NS, 3), quirk=True, which=0).reshape(B, T, NS)
    fin = np.isfinite(s["sdf_sequence"])
    assert np.abs(mine[fin] - s["sdf_sequence"][fin]).max() < 2e-6
    assert ((mine.reshape(B, -1) <= 0.06).any(-1) == s["has_collision_r006"]).all()  # model.py:309-311


def test_quirk_is_latent_for_yaw_only_scenes(oracle):
    g = np.load(os.path.join(HERE, "golden", "sdf_reference.npz"))
    s = {k[4:]: g[k] for k in g.files if k.startswith("yaw_")}
    a = oracle.sdf_points(s, s["points"], quirk=True)
    b = oracle.sdf_points(s, s["points"], quirk=False)
    assert np.array_equal(a, b)


# ----------------------------------------------------------------------------- spec arithmetic
def test_sincos_accuracy(oracle):
    x = np.linspace(-8, 8, 200001).astype(np.float32)
    s, c = oracle.sincos(x)
    assert np.abs(s - np.sin(x.astype(np.float64))).max() < 2e-7
    assert np.abs(c - np.cos(x.astype(np.float64))).max() < 2e-7


def test_normalize_roundtrip_and_formula(oracle, tables):
    rng = np.random.RandomState(1)
    lim = tables.joint_limits
    q = rng.uniform(lim[:, 0], lim[:, 1], size=(100, 7)).astype(np.float32)
    qn = oracle.normalize(q, lim)
    assert qn.min() >= -1 - 1e-6 and qn.max() <= 1 + 1e-6
    ref = (q - lim[:, 0]) / (lim[:, 1] - lim[:, 0]) * 2 + -1       # utils.py:91-93 in fp32
    assert np.array_equal(qn, ref.astype(np.float32))
    back = oracle.unnormalize(qn, lim)
    ref_back = (qn - -1) * (lim[:, 1] - lim[:, 0]) / 2 + lim[:, 0]  # utils.py:207-209
    assert np.array_equal(back, ref_back.astype(np.float32))
    assert np.abs(back - q).max() < 1e-6


def test_philox_known_answer(oracle):
    # Random123 known-answer test for philox4x32-10: counter 0, key 0
    out = oracle.philox((0, 0, 0, 0), (0, 0))
    assert [hex(v) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    out = oracle.philox((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF))
    assert [hex(v) for v in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


@pytest.mark.parametrize("n", [1, 2, 7, 478, 4096, 4097, 10613, 44096])
def test_feistel_is_a_permutation(oracle, n):
    key = np.array([1, 2, 3, 4], np.uint32) * 0x9E3779B1
    p = oracle.feistel(n, key, n)
    assert sorted(p.tolist()) == list(range(n))


# ----------------------------------------------------------------------------- cloud construction properties
def test_obstacle_cloud_properties(oracle):
    from mpinets_b200 import scenes
    p = scenes.config_problems(4, 12)
    pts, prim = oracle.sample_obstacles(p, 4096, seed=7, return_prims=True)
    assert (pts[..., 3] == 1).all()
    B, M1 = p["cuboid_dims"].shape[:2]
    for b in range(B):
        valid_c = np.abs(p["cuboid_dims"][b]).min(-1) > 1e-8
        valid_y = (p["cylinder_radii"][b, :, 0] > 1e-8) & (p["cylinder_heights"][b, :, 0] > 1e-8)
        areas = np.concatenate([2 * (p["cuboid_dims"][b, :, 0] * p["cuboid_dims"][b, :, 1] + p["cuboid_dims"][b, :, 0] * p["cuboid_dims"][b, :, 2]
                                     + p["cuboid_dims"][b, :, 1] * p["cuboid_dims"][b, :, 2]) * valid_c,
                                (2 * np.pi * p["cylinder_radii"][b, :, 0] * (p["cylinder_heights"][b, :, 0] + p["cylinder_radii"][b, :, 0])) * valid_y])
        counts = np.bincount(prim[b], minlength=len(areas))
        assert counts[areas == 0].sum() == 0                      # zero-volume rows are never sampled
        pool = np.floor(areas / areas.sum() * 4096).astype(int) + 500 * (areas > 0)   # geometry.py:599
        expect = pool / pool.sum() * 4096                          # hypergeometric mean of the subsample (:608)
        assert np.abs(counts - expect).max() < 6 * np.sqrt(expect.max()) + 10
        # every point lies on the surface of the primitive it was drawn from
        for m in np.unique(prim[b]):
            sel = prim[b] == m
            one = {k: np.zeros_like(v[b:b + 1]) for k, v in p.items() if k.startswith(("cuboid", "cylinder"))}
            one["cuboid_quats"][..., 0] = 1; one["cylinder_quats"][..., 0] = 1
            if m < M1:
                for k in ("cuboid_centers", "cuboid_dims", "cuboid_quats"):
                    one[k][0, 0] = p[k][b, m]
            else:
                for k in ("cylinder_centers", "cylinder_radii", "cylinder_heights", "cylinder_quats"):
                    one[k][0, 0] = p[k][b, m - M1]
            d = oracle.sdf_points(one, pts[b:b + 1, sel, :3], quirk=False)
            assert np.abs(d).max() < 1e-5


def test_empty_scene_cloud(oracle, tables):
    from mpinets_b200 import scenes
    p = scenes.config_problems(2, 2)
    for k in ("cuboid_dims", "cylinder_radii", "cylinder_heights"):
        p[k][:] = 0
    cloud = oracle.build_cloud(p["q0"], p["target"], p, tables, seed=1)
    assert (cloud[:, 2048:6144, :3] == 0).all() and (cloud[:, 2048:6144, 3] == 1).all()
    assert np.isinf(oracle.sdf_points(p, cloud[:, :16, :3])).all()
    flags, first, _ = oracle.sweep_flags(p, np.repeat(p["q0"][:, None], 3, 1), tables)
    assert not flags.any() and (first == -1).all()


def test_cloud_layout_and_robot_rows(oracle, tables):
    from mpinets_b200 import scenes
    p = scenes.config_problems(2, 3)
    cloud = oracle.build_cloud(p["q0"], p["target"], p, tables, seed=5)
    assert cloud.shape == (3, 6272, 4)
    assert (cloud[:, :2048, 3] == 0).all() and (cloud[:, 2048:6144, 3] == 1).all() and (cloud[:, 6144:, 3] == 2).all()
    # robot rows = FK-transformed canonical points of a keyed subset without repetition
    frames, _ = oracle.fk(p["q0"])
    fr = frames[:, tables.link_ids]                                 # [B,P,3,4]
    world = np.einsum("bpij,pj->bpi", fr[..., :3], tables.link_points) + fr[..., 3]
    for b in range(3):
        d = np.abs(cloud[b, :2048, None, :3] - world[b][None]).max(-1)
        nearest = d.argmin(1)
        assert d.min(1).max() < 1e-6 and len(set(nearest.tolist())) == 2048
    # target rows are the gripper points under the target pose
    tw = np.einsum("bij,pj->bpi", p["target"][:, :, :3], tables.ee_points) + p["target"][:, None, :, 3]
    for b in range(3):
        d = np.abs(cloud[b, 6144:, None, :3] - tw[b][None]).max(-1)
        assert d.min(1).max() < 1e-6
    # resampling at another step changes the subset but not the surface
    c2 = oracle.sample_robot(p["q0"], tables, 2048, 5, 9)
    assert not np.array_equal(c2[:, :, :3], cloud[:, :2048, :3])


# ----------------------------------------------------------------------------- pointnet2_ops restatement
def _fps_naive(xyz, m):
    """greedy farthest point sampling with first-index ties (no tree, no skip): a weaker independent statement"""
    N = len(xyz)
    d = np.full(N, 1e10, np.float32)
    out = [0]
    for _ in range(1, m):
        diff = (xyz - xyz[out[-1]]).astype(np.float32)
        dd = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1] + diff[:, 2] * diff[:, 2]).astype(np.float32)
        d = np.minimum(d, dd)
        out.append(int(d.argmax()))
    return out


def test_fps_is_a_greedy_farthest_sequence(oracle):
    rng = np.random.RandomState(3)
    xyz = (rng.uniform(-1, 1, size=(2, 700, 3)) + 2.0).astype(np.float32)   # away from the origin: no skipped points
    idx = oracle.fps(xyz, 64)
    for b in range(2):
        assert len(set(idx[b].tolist())) == 64 and idx[b, 0] == 0
        # each pick maximises the distance to the already-picked set (values compared, index may differ on fp ties)
        picked = [0]
        for j in range(1, 64):
            d = ((xyz[b][:, None] - xyz[b][picked][None]) ** 2).sum(-1).min(1)
            assert d[idx[b, j]] >= d.max() * (1 - 1e-5)
            picked.append(idx[b, j])


def test_fps_skips_points_near_origin(oracle):
    rng = np.random.RandomState(4)
    xyz = rng.uniform(-1, 1, size=(1, 600, 3)).astype(np.float32)
    xyz[0, 100:140] *= 0.01          # |p|^2 <= 1e-3 -> never selected (sampling_gpu.cu skip rule)
    idx = oracle.fps(xyz, 300)
    assert not set(range(100, 140)) & set(idx[0, 1:].tolist())


def _origin_boundary_cloud():
    """a cloud whose point 7 has |p|^2 == 0.001f EXACTLY (fma chain of the spec) and point 9 the float just below it"""
    x, y = np.float32(0.020976269617676735), np.float32(0.023664237931370735)
    mag = np.float32(np.float64(y) * np.float64(y) + np.float64(np.float32(x * x)))
    assert mag == np.float32(0.001)
    rng = np.random.RandomState(11)
    xyz = (0.0005 * rng.uniform(-1, 1, size=(1, 40, 3))).astype(np.float32)   # everything else deep inside the skip radius
    xyz[0, 0] = [0.5, 0.5, 0.5]
    xyz[0, 7] = [x, y, 0.0]
    y_lo = np.nextafter(y, np.float32(0))
    while np.float32(np.float64(y_lo) * np.float64(y_lo) + np.float64(np.float32(x * x))) >= np.float32(0.001):
        y_lo = np.nextafter(y_lo, np.float32(0))
    xyz[0, 9] = [-x, -y_lo, 0.0]
    return xyz


def test_fps_origin_skip_compares_against_the_double_literal(oracle):
    """sampling_gpu.cu writes `mag <= 1e-3` with a DOUBLE literal: (double)0.001f = 0.0010000000475 > 0.001, so a point whose
    squared norm is exactly the float 0.001f is NOT skipped, while the next float below is"""
    xyz = _origin_boundary_cloud()
    idx = oracle.fps(xyz, 3)
    assert idx[0, 0] == 0 and idx[0, 1] == 7          # the only other selectable point
    assert idx[0, 2] == 0                             # nothing else is selectable: best = -1 / besti = 0 of the reference kernel


def test_fps_tie_break_is_tree_order(oracle):
    # 4 points at equal distance from point 0 -> winner decided by the shared-memory tree, not by lowest index
    xyz = np.zeros((1, 8, 3), np.float32)
    xyz[0, :, 0] = 5.0
    for k, (dy, dz) in {1: (1, 0), 2: (-1, 0), 3: (0, 1), 6: (0, -1)}.items():
        xyz[0, k, 1], xyz[0, k, 2] = dy, dz
    idx = oracle.fps(xyz, 2)
    # block = 8 threads, one point each; ties: stride-4 stage keeps lower slot, ... final winner has the smallest
    # bit-reversed thread id among {1,2,3,6} = {100b,010b,110b,011b} -> thread 2 (bitrev 010b=2) vs 1 (100b=4): 2 wins
    assert idx[0, 1] == 2


def test_ball_query_semantics(oracle):
    rng = np.random.RandomState(5)
    xyz = rng.uniform(0, 1, size=(2, 900, 3)).astype(np.float32)
    new_xyz = xyz[:, :50].copy()
    idx = oracle.ball_query(0.2, 16, xyz, new_xyz)
    for b in range(2):
        for j in range(50):
            d2 = ((xyz[b] - new_xyz[b, j]) ** 2).sum(-1)
            inside = np.nonzero(d2 < 0.2 * 0.2 * (1 - 1e-5))[0]
            got = idx[b, j]
            n = min(len(inside), 16)
            assert (np.diff(got[:n]) > 0).all()                       # index order
            assert set(got.tolist()) <= set(np.nonzero(d2 < 0.2 * 0.2 * (1 + 1e-5))[0].tolist())
            if len(inside) < 16:
                assert (got[len(set(got.tolist())):] == got[0]).all()  # padded with the first hit
    far = oracle.ball_query(0.01, 4, xyz, np.full((2, 1, 3), 9.0, np.float32))
    assert (far == 0).all()                                            # no hit -> zero-initialised indices


# ----------------------------------------------------------------------------- network restatement vs torch modules
def test_network_restatement_matches_torch_modules(oracle, state_dict):
    """The oracle's dense/GroupNorm/LeakyReLU chain vs an nn.Sequential built exactly like model.py:47-66,385-393;
    the shared MLP vs nn.Conv2d(k=1)+ReLU+max_pool2d on an explicitly grouped tensor (pointnet2 build_shared_mlp)."""
    sd = state_dict
    torch.manual_seed(0)
    fc = torch.nn.Sequential(torch.nn.Linear(1024, 4096), torch.nn.GroupNorm(16, 4096), torch.nn.LeakyReLU(),
                             torch.nn.Linear(4096, 2048), torch.nn.GroupNorm(16, 2048), torch.nn.LeakyReLU(),
                             torch.nn.Linear(2048, 2048))
    fc.load_state_dict({k[len("point_cloud_encoder.fc_layer."):]: v for k, v in sd.items() if "fc_layer" in k})
    x = torch.randn(3, 1024)
    p = "point_cloud_encoder.fc_layer."
    y = oracle._dense(x, sd[p + "0.weight"], sd[p + "0.bias"], False, torch.float32)
    y = F.leaky_relu(F.group_norm(y, 16, sd[p + "1.weight"], sd[p + "1.bias"], eps=1e-5), 0.01)
    y = oracle._dense(y, sd[p + "3.weight"], sd[p + "3.bias"], False, torch.float32)
    y = F.leaky_relu(F.group_norm(y, 16, sd[p + "4.weight"], sd[p + "4.bias"], eps=1e-5), 0.01)
    y = oracle._dense(y, sd[p + "6.weight"], sd[p + "6.bias"], False, torch.float32)
    assert torch.allclose(y, fc(x), atol=1e-5, rtol=1e-5)

    rng = np.random.RandomState(0)
    xyz = (rng.uniform(-0.3, 0.3, size=(2, 300, 3)) + 0.5).astype(np.float32)
    feats = torch.from_numpy(rng.normal(size=(2, 300, 64)).astype(np.float32))
    spec = dict(npoint=16, radius=0.3, nsample=128, mlp=(67, 128, 128, 256))
    ws = [(sd[f"point_cloud_encoder.SA_modules.1.mlps.0.{2 * l}.weight"], sd[f"point_cloud_encoder.SA_modules.1.mlps.0.{2 * l}.bias"]) for l in range(3)]
    new_xyz, out, aux = oracle.sa_module(xyz, feats, spec, ws, return_aux=True)
    # explicit pointnet2 formulation: [B,C,N] features, grouping_operation, Conv2d 1x1
    f_cn = feats.permute(0, 2, 1).contiguous().numpy()
    g_xyz = oracle.grouping_operation(np.ascontiguousarray(xyz.transpose(0, 2, 1)), aux["ball_idx"]) - new_xyz.transpose(0, 2, 1)[..., None]
    g = torch.from_numpy(np.concatenate([g_xyz, oracle.grouping_operation(f_cn, aux["ball_idx"])], axis=1))  # [B,67,m,ns]
    for w, b in ws:
        g = F.relu(F.conv2d(g, w, b))
    ref = F.max_pool2d(g, kernel_size=[1, g.size(3)]).squeeze(-1)   # [B,256,m]
    assert torch.allclose(out.permute(0, 2, 1), ref, atol=1e-5, rtol=1e-5)


def test_state_dict_layout(state_dict):
    n = sum(v.numel() for v in state_dict.values())
    assert n == 19068103                                              # parameter count
    assert state_dict["point_cloud_encoder.SA_modules.0.mlps.0.0.weight"].shape == (64, 4, 1, 1)
    assert state_dict["decoder.0.weight"].shape == (512, 2112)


# ----------------------------------------------------------------------------- C ABI surface (no GPU needed)
def test_shared_library_exports_every_declared_symbol():
    from mpinets_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from mpinets_b200 import build
        build.build()
    header = open(os.path.join(ROOT, "include", "mpinets_b200.h")).read()
    declared = set(re.findall(r"\b(mpn_[a-z0-9_]+)\s*\(", header))
    declared -= {"mpn_ctx", "mpn_scene", "mpn_config", "mpn_status", "mpn_precision"}
    lib = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/mpinets_b200.h but not exported"
    assert set(_lib.EXPORTS) == declared


def test_library_reports_errors_without_gpu():
    from mpinets_b200 import _lib
    lib = _lib.load()
    assert b"mpinets_b200" in lib.mpn_version()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = C.c_void_p()
    cfg = _lib.MpnConfig(2048, 4096, 128, 40, 40, 1, 0)
    rc = lib.mpn_ctx_create(0, C.byref(cfg), C.byref(ctx))
    assert rc != 0 and len(lib.mpn_last_error()) > 0                 # fails loudly, no CPU fallback
    with pytest.raises(_lib.MpnError):
        from mpinets_b200.engine import Engine
        Engine()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mpinets_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"
                assert "mpn_oracle" not in src, f"{f} references the oracle library"


# ----------------------------------------------------------------------------- Evaluator subset (metrics.py:311-523)
from mpinets_b200 import scenes  # noqa: E402
from mpinets_b200.franka import fk_reference_f64  # noqa: E402


def _eval_inputs(B=10, T1=21, seed=5):
    """joint-space interpolation q0 -> q_goal (reaches the target exactly) with a few rows perturbed"""
    p = scenes.config_problems(4, B)
    w = np.linspace(0.0, 1.0, T1, dtype=np.float32)[None, :, None]
    traj = (p["q0"][:, None, :] * (1 - w) + p["q_goal"][:, None, :] * w).astype(np.float32)
    rng = np.random.default_rng(seed)
    traj[1, -1] += 0.2                       # misses the target
    traj[2, 5, 3] = 0.5                      # joint 4 upper limit is -0.0698 -> violation
    traj[3, :, :] = np.array([1.638, 1.227, 0.041, -3.039, 0.047, 1.604, 0.314], np.float32) + 0.01 * rng.standard_normal((T1, 7)).astype(np.float32)
    return p, np.ascontiguousarray(traj)


def _angle_deg_f64(A, B):
    R = A[:3, :3].astype(np.float64) @ B[:3, :3].astype(np.float64).T
    return np.degrees(np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1)))


def test_oracle_evaluate_matches_float64_numpy(oracle, tables):
    p, traj = _eval_inputs()
    B, T1, _ = traj.shape
    num = np.full(B, T1, np.int32); num[4] = 7
    tv = dict(cuboid_centers=p["target"][:, None, :3, 3].copy(), cuboid_dims=np.full((B, 1, 3), 0.1, np.float32),
              cuboid_quats=np.tile(np.array([1, 0, 0, 0], np.float32), (B, 1, 1)))
    tv["cuboid_centers"][5, 0, 0] += 1.0     # final pose is outside its target volume
    nv = dict(cuboid_centers=p["target"][:, None, :3, 3].copy() + np.array([0.3, 0, 0], np.float32),
              cuboid_dims=np.full((B, 1, 3), 0.2, np.float32), cuboid_quats=tv["cuboid_quats"].copy())
    nv["cuboid_dims"][6] = 1.0               # contains the target itself -> dropped by metrics.py:497-499
    ev = oracle.evaluate(p, traj, p["target"], tables, num_poses=num, target_volume=tv, negative_volumes=nv)
    col = {k: i for i, k in enumerate(("collision", "joint_limit_violation", "self_collision", "physical_violations",
                                       "position_error", "orientation_error", "eff_position_path_length",
                                       "eff_orientation_path_length", "correct_final_region", "success", "num_steps",
                                       "first_collision_step", "config_path_length", "max_collision_depth"))}
    flags, first, _ = oracle.sweep_flags(p, traj, tables)
    lim = tables.joint_limits
    for b in range(B):
        n = int(num[b])
        eef = np.stack([fk_reference_f64(traj[b, t].astype(np.float64))[1] for t in range(n)])
        tg = np.eye(4); tg[:3] = p["target"][b]
        assert abs(ev[b, col["position_error"]] - 100 * np.linalg.norm(eef[-1, :3, 3] - tg[:3, 3])) < 2e-3
        assert abs(ev[b, col["orientation_error"]] - _angle_deg_f64(eef[-1], tg)) < 5e-2
        steps = np.linalg.norm(np.diff(eef[:, :3, 3], axis=0), axis=1).sum()
        assert abs(ev[b, col["eff_position_path_length"]] - steps) < 1e-4
        ori = sum(_angle_deg_f64(eef[t + 1], eef[t]) for t in range(n - 1))
        assert abs(ev[b, col["eff_orientation_path_length"]] - ori) < 0.05 * max(1, n / 10)
        cfg = np.linalg.norm(np.diff(traj[b, :n].astype(np.float64), axis=0), axis=1).sum()
        assert abs(ev[b, col["config_path_length"]] - cfg) < 1e-4
        jl = bool(((traj[b, :n] < lim[:, 0]) | (traj[b, :n] > lim[:, 1])).any())
        assert bool(ev[b, col["joint_limit_violation"]]) == jl
        assert ev[b, col["num_steps"]] == n
        if n == T1:
            assert bool(ev[b, col["collision"]]) == bool(flags[b]) and int(ev[b, col["first_collision_step"]]) == int(first[b])
        phys = ev[b, col["collision"]] or ev[b, col["joint_limit_violation"]] or ev[b, col["self_collision"]]
        assert bool(ev[b, col["physical_violations"]]) == bool(phys)
        succ = (ev[b, col["position_error"]] < 1 and ev[b, col["orientation_error"]] < 15 and ev[b, col["correct_final_region"]]
                and not phys)
        assert bool(ev[b, col["success"]]) == bool(succ)
    assert ev[2, col["joint_limit_violation"]] == 1 and ev[0, col["joint_limit_violation"]] == 0
    assert ev[0, col["position_error"]] < 1e-2 and ev[1, col["position_error"]] > 1.0
    assert ev[5, col["correct_final_region"]] == 0 and ev[0, col["correct_final_region"]] == 1
    assert ev[6, col["correct_final_region"]] == 1       # the negative volume that swallows the target is ignored
    assert ev[3, col["self_collision"]] == 1             # elbow fully flexed: gripper spheres 7.5 cm inside the link-2 spheres
    # max_collision_depth is the deepest sphere penetration r - sdf over the trajectory (0 when never in contact)
    assert ((ev[:, col["max_collision_depth"]] >= 0) & ((ev[:, col["max_collision_depth"]] > 0) | (ev[:, col["collision"]] == 0))).all()


def test_oracle_evaluate_region_against_reference_sdf(oracle, tables):
    """the region test's cuboid SDF equals the real TorchCuboids.sdf for yaw-only volumes (fixture from geometry.py)"""
    p, traj = _eval_inputs(B=4, T1=8)
    B = 4
    rng = np.random.default_rng(0)
    for trial in range(8):
        yaw = rng.uniform(-np.pi, np.pi, B)
        quat = np.stack([np.cos(yaw / 2), 0 * yaw, 0 * yaw, np.sin(yaw / 2)], -1).astype(np.float32)[:, None]
        dims = rng.uniform(0.05, 0.4, (B, 1, 3)).astype(np.float32)
        fin = np.stack([fk_reference_f64(traj[b, -1].astype(np.float64))[1][:3, 3] for b in range(B)]).astype(np.float32)
        cen = (fin + rng.uniform(-0.2, 0.2, (B, 3))).astype(np.float32)[:, None]
        tv = dict(cuboid_centers=cen, cuboid_dims=dims, cuboid_quats=quat)
        ev = oracle.evaluate(p, traj, p["target"], tables, target_volume=tv)
        sc = dict(p); sc.update(cuboid_centers=cen, cuboid_dims=dims, cuboid_quats=quat)
        sc["cylinder_centers"] = np.zeros((B, 1, 3), np.float32); sc["cylinder_radii"] = np.zeros((B, 1, 1), np.float32)
        sc["cylinder_heights"] = np.zeros((B, 1, 1), np.float32); sc["cylinder_quats"] = np.tile(np.array([1, 0, 0, 0], np.float32), (B, 1, 1))
        fin_spec = oracle.fk(traj[:, -1])[1][:, :3, 3]
        sdf = oracle.sdf_points(sc, fin_spec[:, None, :], quirk=False, which=1)[:, 0]
        assert ((sdf <= 0) == (ev[:, 8] > 0)).all()


# ----------------------------------------------------------------------------- losses (loss.py:31-166)
def test_oracle_losses_match_reference_loss_py(oracle):
    """collision_loss / point_match_loss values and autograd gradients produced by the REAL mpinets/loss.py
    (tests/golden/make_golden.py -> loss_reference.npz)"""
    g = np.load(os.path.join(HERE, "golden", "loss_reference.npz"))
    for tag in ("yaw", "free"):
        sc = {k: g[f"{tag}_{k}"] for k in scenes.SCENE_KEYS}
        val, grad = oracle.collision_loss(sc, g[f"{tag}_points"], margin=0.03, quirk=True)
        assert abs(val - float(g[f"{tag}_collision_loss"])) < 2e-7
        ref = g[f"{tag}_collision_grad"]
        assert (ref != 0).any(axis=-1).sum() > 50                         # plenty of points inside the margin
        assert np.array_equal(grad != 0, ref != 0)                        # identical active set
        assert np.abs(grad - ref).max() < 1e-5 * np.abs(ref).max()      # fp32 autograd vs closed form
        val, grad = oracle.point_match_loss(g[f"{tag}_points"], g[f"{tag}_other"])
        assert abs(val - float(g[f"{tag}_point_match_loss"])) < 1e-7
        assert np.abs(grad - g[f"{tag}_point_match_grad"]).max() < 1e-9


def test_oracle_fixed_robot_cloud_properties(oracle, tables):
    """FrankaSampler(num_fixed_points=1024, with_base_link=False) stand-in: fixed, duplicate-free, no base-link rows,
    and equal to FK applied to those table rows"""
    q = np.array([[0.1, -0.5, 0.2, -2.0, 0.3, 1.5, 0.7], [-1.0, 0.4, 1.1, -1.2, -0.6, 2.2, -0.3]], np.float32)
    pts, idx = oracle.fixed_robot_points(q, tables, 1024, 0x4D50694E)
    pts2, idx2 = oracle.fixed_robot_points(q[::-1].copy(), tables, 1024, 0x4D50694E)
    assert np.array_equal(idx, idx2) and np.array_equal(pts[0], pts2[1])
    assert len(set(idx.tolist())) == 1024 and (tables.link_ids[idx] != 0).all()
    frames, _ = oracle.fk(q)
    for b in range(2):
        F = frames[b][tables.link_ids[idx]]
        ref = np.einsum("nij,nj->ni", F[:, :, :3].astype(np.float64), tables.link_points[idx].astype(np.float64)) + F[:, :, 3]
        assert np.abs(pts[b] - ref).max() < 1e-6


def test_oracle_bc_collision_losses_gradient(oracle, tables):
    """CollisionAndBCLossContainer (loss.py:111-166): values equal the two standalone losses on the fixed clouds; the analytic
    gradient w.r.t. input_normalized matches central finite differences of the loss itself"""
    p = scenes.config_problems(4, 6)
    seed = 0x4D50694E
    rng = np.random.default_rng(3)
    qi = rng.uniform(-0.9, 0.9, (6, 7)).astype(np.float32)
    qt = np.clip(qi + rng.normal(scale=0.05, size=qi.shape), -1, 1).astype(np.float32)
    losses, grad = oracle.bc_collision_losses(p, qi, qt, tables, seed, w_collision=5.0, w_bc=1.0)
    xi, _ = oracle.fixed_robot_points(oracle.unnormalize(qi, tables.joint_limits), tables, 1024, seed)
    xt, _ = oracle.fixed_robot_points(oracle.unnormalize(qt, tables.joint_limits), tables, 1024, seed)
    assert abs(losses[0] - oracle.collision_loss(p, xi)[0]) < 1e-7
    assert abs(losses[1] - oracle.point_match_loss(xi, xt)[0]) < 1e-7

    def total(qn):
        l, _ = oracle.bc_collision_losses(p, qn.astype(np.float32), qt, tables, seed, w_collision=5.0, w_bc=1.0)
        return 5.0 * float(l[0]) + float(l[1])
    eps = 5e-4
    fd = np.zeros((6, 7))
    for b in range(6):
        for j in range(7):
            d = np.zeros_like(qi); d[b, j] = eps
            fd[b, j] = (total(qi + d) - total(qi - d)) / (2 * eps)
    rel = np.linalg.norm(fd - grad, axis=1) / np.linalg.norm(grad, axis=1)
    assert rel.max() < 0.02          # fp32 losses with |.| / hinge kinks differenced at 5e-4: < 1 % noise; a wrong Jacobian term is O(1)
    # the collision part on its own is non-trivial for at least one of the problems
    _, gc = oracle.bc_collision_losses(p, qi, qt, tables, seed, w_collision=1.0, w_bc=0.0)
    assert (np.abs(gc).max(axis=1) > 1e-3).sum() >= 1


# ----------------------------------------------------------------------------- PlanningProblem -> SoA (mpinets_types.py:34-48)
def test_planning_problem_records_to_soa():
    from mpinets_b200 import mpinets_types as T
    p = scenes.config_problems(4, 7)
    probs = T.soa_to_problems(p)
    assert all(isinstance(x, T.PlanningProblem) for x in probs) and len(probs[0].obstacles) > 0
    probs[2].target_negative_volumes = [T.Cuboid([0.5, 0, 0.3], [0.2, 0.2, 0.2]), T.Cylinder([0.4, 0.1, 0.2], 0.05, 0.3)]
    probs[3].target_volume = T.Cylinder(probs[3].target.xyz, 0.1, 0.2)
    s = T.problems_to_soa(probs)
    assert np.array_equal(s["q0"], p["q0"]) and np.abs(s["target"] - p["target"]).max() < 1e-6
    for b in range(7):   # the valid primitives survive in order; padding rows are zero-volume with unit quaternions
        for fam, dimkey in (("cuboid", "cuboid_dims"), ("cylinder", "cylinder_radii")):
            keep_a = ~np.isclose(p[dimkey][b].reshape(p[dimkey].shape[1], -1), 0).any(axis=1)
            keep_b = ~np.isclose(s[dimkey][b].reshape(s[dimkey].shape[1], -1), 0).any(axis=1)
            if fam == "cylinder":
                keep_a &= ~np.isclose(p["cylinder_heights"][b, :, 0], 0); keep_b &= ~np.isclose(s["cylinder_heights"][b, :, 0], 0)
            assert np.allclose(p[f"{fam}_centers"][b][keep_a], s[f"{fam}_centers"][b][keep_b])
            assert np.allclose(p[f"{fam}_quats"][b][keep_a], s[f"{fam}_quats"][b][keep_b], atol=1e-6)
            assert (s[f"{fam}_quats"][b][~keep_b] == np.array([1, 0, 0, 0], np.float32)).all()
    assert s["negative_volumes"]["cuboid_dims"].shape == (7, 1, 3) and s["negative_volumes"]["cylinder_radii"].shape == (7, 1, 1)
    assert s["negative_volumes"]["cylinder_radii"][2, 0, 0] == np.float32(0.05) and s["negative_volumes"]["cuboid_dims"][0].max() == 0
    assert s["target_volume"]["cylinder_radii"][3, 0, 0] == np.float32(0.1) and s["target_volume"]["cuboid_dims"][3].max() == 0
    with pytest.raises(ValueError):
        T.primitives_to_soa([[T.Cuboid([0, 0, 0], [1, 1, 1])] * 3], 2, 2)
    ps = {"tabletop": {"task_oriented": probs[:3], "neutral_start": probs[3:5]}, "cubby": {"task_oriented": probs[5:]}}
    flat = T.flatten_problem_set(ps)
    assert [(e, k) for e, k, _ in flat] == [("tabletop", "task_oriented")] * 3 + [("tabletop", "neutral_start")] * 2 + [("cubby", "task_oriented")] * 2


# ----------------------------------------------------------------------------- SPARC (third_party/sparc.py)
def test_oracle_sparc_matches_reference(oracle):
    """the reference's own doctest value (sparc.py:87-91) and outputs of the real sparc() on rollout-shaped speed profiles"""
    g = np.load(os.path.join(HERE, "golden", "sparc_reference.npz"))
    assert "%.5f" % oracle.sparc(g["doctest_move"], 100.0) == "-1.41403"
    assert abs(oracle.sparc(g["doctest_move"], 100.0) - float(g["doctest_sal"])) < 1e-12
    for b in range(g["profiles"].shape[0]):
        n = int(g["num"][b])
        assert abs(oracle.sparc(g["profiles"][b, :n], float(g["fs"])) - g["sal"][b]) < 1e-12
    assert oracle.sparc(np.zeros(30), 12.5) == 0.0


def test_oracle_cloud_from_obstacle_points(oracle, tables):
    """make_point_cloud_from_problem (run_inference.py:58-90): the obstacle rows are a duplicate-free subset of the given
    cloud; robot and target rows are the ones of the primitive-based build"""
    p = scenes.config_problems(2, 3)
    rng = np.random.default_rng(0)
    counts = np.array([9000, 4096, 5000], np.int32)
    pts = rng.uniform(-1, 1, (3, 9000, 3)).astype(np.float32)
    seed = 0x4D50694E
    c = oracle.build_cloud_from_points(p["q0"], p["target"], pts, counts, tables, seed, problem0=5)
    ref = oracle.build_cloud(p["q0"], p["target"], p, tables, seed, problem0=5)
    assert np.array_equal(c[:, :2048], ref[:, :2048]) and np.array_equal(c[:, 6144:], ref[:, 6144:])
    assert (c[:, 2048:6144, 3] == 1).all()
    for b in range(3):
        rows = {tuple(r) for r in c[b, 2048:6144, :3]}
        pool = {tuple(r) for r in pts[b, : counts[b]]}
        assert len(rows) == 4096 and rows <= pool            # without replacement, only from the valid prefix
    assert not np.array_equal(c[0, 2048:6144], oracle.build_cloud_from_points(p["q0"], p["target"], pts, counts, tables, seed, problem0=6)[0, 2048:6144])


# ----------------------------------------------------------------------------- depth-camera clouds (run_inference.py:194-257)
def test_depth_render_known_answer_and_on_surface(oracle):
    """unit cube 2 m in front of a camera at the origin: its front face is at depth 2; every hit lies on a primitive surface
    (|sdf| ~ 0 with the SAME frames as the SDF), inside the frustum, and hits are in pixel order"""
    shapes = dict(cuboid_centers=(1, 40, 3), cuboid_dims=(1, 40, 3), cuboid_quats=(1, 40, 4), cylinder_centers=(1, 40, 3),
                  cylinder_radii=(1, 40), cylinder_heights=(1, 40), cylinder_quats=(1, 40, 4))
    p = {k: np.zeros(s, np.float32) for k, s in shapes.items()}
    p["cuboid_quats"][..., 0] = 1; p["cylinder_quats"][..., 0] = 1
    p["cuboid_centers"][0, 0] = [0, 0, 2.5]; p["cuboid_dims"][0, 0] = [1, 1, 1]
    p["cylinder_centers"][0, 0] = [2, 0, 3]; p["cylinder_radii"][0, 0] = 0.5; p["cylinder_heights"][0, 0] = 1.0
    cam = np.diag([1.0, -1.0, -1.0, 1.0]).astype(np.float32)[:3]   # GL camera (y up, looks along -z) turned to look along +z
    W, H = 64, 48
    pts, cnt = oracle.render_depth_cloud(p, cam, W, H, 60.0)
    q = pts[0, :cnt[0]]
    assert 0 < cnt[0] < W * H
    assert abs(q[:, 2].min() - 2.0) < 1e-6                       # the cube's front face
    centre = q[(np.abs(q[:, 0]) < 0.4) & (np.abs(q[:, 1]) < 0.4)]
    assert len(centre) and np.abs(centre[:, 2] - 2.0).max() < 1e-6
    assert np.abs(oracle.sdf_points(p, q[None])).max() < 1e-5
    ty = np.tan(np.radians(30.0))
    assert (np.abs(q[:, 1] / q[:, 2]) <= ty + 1e-6).all() and (np.abs(q[:, 0] / q[:, 2]) <= ty * W / H + 1e-6).all()
    # pixel order: un-project to pixel indices, they must increase
    col = np.floor((q[:, 0] / q[:, 2] / (ty * W / H) + 1) * W / 2).astype(int)
    row = np.floor((q[:, 1] / q[:, 2] / ty + 1) * H / 2).astype(int)
    assert (np.diff(row * W + col) > 0).all()
    # nothing behind the far plane; an empty scene gives no points
    _, c2 = oracle.render_depth_cloud(p, cam, W, H, 60.0, far=1.5)
    assert c2[0] == 0


def test_depth_render_eval_cameras_see_the_scenes(oracle):
    """the reference's evaluation cameras (run_inference.py:215-243), GL convention, see the synthetic scenes of their type"""
    from mpinets_b200 import scenes
    from mpinets_b200.run_inference import eval_camera
    for config, env in ((2, "tabletop"), (3, "cubby")):
        p = scenes.config_problems(config, 3)
        pts, cnt = oracle.render_depth_cloud(p, eval_camera(env), 80, 60)
        assert (cnt > 200).all(), (env, cnt)
        for b in range(3):
            assert np.abs(oracle.sdf_points({k: v[b:b + 1] for k, v in p.items() if k in scenes.SCENE_KEYS}, pts[b:b + 1, :cnt[b]])).max() < 1e-4


def test_packed_tile_rule_of_the_sa_kernels(oracle):
    """the first-fit rule by which the fused set-abstraction kernels pack the distinct rows of a round's groups into 128-row tiles
    (oracle.packed_tile_count mirrors tc_common.cuh; the GPU tests compare the device's tile counters against it)"""
    assert oracle.packed_tile_count([5, 9, 31, 32]) == 1                      # four quarters
    assert oracle.packed_tile_count([33, 1, 1, 1]) == 2                       # 2 + 1 + 1 quarters, the fourth group opens a tile
    assert oracle.packed_tile_count([128, 128, 128, 128]) == 4                # full groups: the reference formulation
    assert oracle.packed_tile_count([65, 64, 1, 1]) == 2                      # 3 | 2 + 1 + 1
    assert oracle.packed_tile_count([1] * 8) == 2                             # a round never spans more than four centroids ...
    assert oracle.packed_tile_count([1] * 8, per_round=8, gran=16) == 1       # ... eight at 16-row granularity
    assert oracle.packed_tile_count([17, 16, 100, 3, 40, 40, 40, 40], per_round=8, gran=16) == 4   # eighths: 2+1 | 7+1 | 3+3 | 3+3
    idx = np.zeros((1, 2, 128), np.int32)
    idx[0, 0, :5] = [3, 7, 9, 11, 400]; idx[0, 0, 5:] = 3                      # 5 hits padded with the first
    idx[0, 1] = np.arange(128)
    assert oracle.distinct_neighbour_counts(idx).tolist() == [[5, 128]]
