"""GPU parity tests (-m gpu): every call goes through the C ABI (mpinets_b200.engine.Engine -> libmpinets_b200.so)
and is compared with the CPU oracle on identical seeded inputs.

Bars: bit-exact for geometry (FK frames, sphere centres, cloud coordinates, SDF values), FPS / ball-query indices and
collision flags; 1e-5 for delta-q in the fp32 mode (BASELINE.json north_star)."""
import os

import numpy as np
import pytest
import torch

from conftest import to_dev

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

DQ_TOL = 1e-5   # north_star: "delta-q within 1e-5 fp32"


def _problems(config, B, seed=0x4D50694E, problem0=0):
    from mpinets_b200 import scenes
    return scenes.config_problems(config, B, seed, problem0)


def _rand_q(tables, n, seed):
    rng = np.random.RandomState(seed)
    lim = tables.joint_limits
    return rng.uniform(lim[:, 0], lim[:, 1], size=(n, 7)).astype(np.float32)


# ----------------------------------------------------------------------------- robofin replacements
def test_fk_bit_exact(engine, oracle, tables):
    q = _rand_q(tables, 4099, 0)
    q[0] = np.load(os.path.join(HERE, "golden", "fk_reference.npz"))["q"]
    frames, eef = engine.fk(torch.from_numpy(q).cuda())
    of, oe = oracle.fk(q)
    assert np.array_equal(frames.cpu().numpy(), of)
    assert np.array_equal(eef.cpu().numpy(), oe)
    g = np.load(os.path.join(HERE, "golden", "fk_reference.npz"))
    assert np.abs(eef[0, :, 3].cpu().numpy() - g["xyz"]).max() < 2e-7   # reference known-answer pair


def test_joint_normalisation_bit_exact(engine, oracle, tables):
    q = _rand_q(tables, 1000, 1)
    qn = engine.normalize(torch.from_numpy(q).cuda())
    assert np.array_equal(qn.cpu().numpy(), oracle.normalize(q, tables.joint_limits))
    back = engine.unnormalize(qn)
    assert np.array_equal(back.cpu().numpy(), oracle.unnormalize(qn.cpu().numpy(), tables.joint_limits))


def test_collision_spheres_bit_exact(engine, oracle, tables):
    q = _rand_q(tables, 513, 2)
    got = engine.compute_spheres(torch.from_numpy(q).cuda()).cpu().numpy()
    assert np.array_equal(got, oracle.spheres(q, tables))


def test_sample_robot_bit_exact(engine, oracle, tables):
    q = _rand_q(tables, 65, 3)
    for step, n in ((0, 2048), (17, 2048), (3, 1024)):
        got = engine.sample_robot(torch.from_numpy(q).cuda(), n, step).cpu().numpy()
        exp = oracle.sample_robot(q, tables, n, engine.cfg.seed, step)
        assert np.array_equal(got, exp)


# ----------------------------------------------------------------------------- geometry.py replacements
@pytest.mark.parametrize("tag", ["yaw", "free"])
def test_sdf_matches_real_reference_fixture(engine, tag):
    """CUDA vs the values produced by the real mpinets/geometry.py (tests/golden/make_golden.py)."""
    from mpinets_b200.engine import Engine
    g = np.load(os.path.join(HERE, "golden", "sdf_reference.npz"))
    s = {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + "_")}
    e = Engine(max_cuboids=s["cuboid_dims"].shape[1], max_cylinders=s["cylinder_radii"].shape[1], tables=engine.tables)
    sc = to_dev(s, [k for k in s if k.startswith(("cuboid", "cylinder"))])
    pts = torch.from_numpy(s["points"]).cuda()
    for which, key in ((1, "sdf_cuboids"), (2, "sdf_cylinders")):
        got = e.sdf_points(sc, pts, which).cpu().numpy()
        fin = np.isfinite(s[key])
        assert (np.isfinite(got) == fin).all()
        assert np.abs(got[fin] - s[key][fin]).max() < 2e-6
    B, T, NS, _ = s["seq"].shape
    got = e.sdf_points(sc, torch.from_numpy(s["seq"].reshape(B, T * NS, 3)).cuda(), 0).cpu().numpy()
    assert ((got <= 0.06).any(-1) == s["has_collision_r006"]).all()
    e.close()


@pytest.mark.parametrize("config", [2, 3, 4])
def test_sdf_points_bit_exact(engine, oracle, config):
    p = _problems(config, 64)
    rng = np.random.RandomState(config)
    pts = rng.uniform(-1.0, 1.5, size=(64, 777, 3)).astype(np.float32)
    for which in (0, 1, 2):
        got = engine.sdf_points(to_dev(p), torch.from_numpy(pts).cuda(), which).cpu().numpy()
        exp = oracle.sdf_points(p, pts, quirk=True, which=which)
        assert np.array_equal(got, exp)


@pytest.mark.parametrize("config,B,T", [(2, 512, 50), (3, 512, 70), (4, 1024, 70)])
def test_collision_flags_bit_exact(engine, oracle, tables, config, B, T):
    p = _problems(config, B)
    rng = np.random.RandomState(10 + config)
    # config-1 style poses: linear interpolation in joint space between two in-limit configurations
    a, b = _rand_q(tables, B, 20 + config), _rand_q(tables, B, 30 + config)
    w = np.linspace(0, 1, T, dtype=np.float32)[None, :, None]
    traj = (a[:, None] * (1 - w) + b[:, None] * w).astype(np.float32)
    flags, first = engine.sweep_flags(to_dev(p), torch.from_numpy(traj).cuda())
    oflags, ofirst, margin = oracle.sweep_flags(p, traj, tables)
    assert np.array_equal(flags.cpu().numpy(), oflags)                 # collision-flag match == 1.0
    assert np.array_equal(first.cpu().numpy(), ofirst)
    assert 0.02 < oflags.mean() < 0.999                                 # the test exercises both outcomes


def test_collision_flags_edge_cases(engine, oracle, tables):
    p = _problems(2, 8)
    p["cuboid_dims"][0] = 0; p["cylinder_radii"][0] = 0                 # empty scene -> never in collision
    p["cuboid_dims"][1, :, 1] = 0                                        # every cuboid zero-volume
    p["cylinder_heights"][2] = 0
    p["cuboid_dims"][3] = 1e-9                                           # below isclose atol -> masked
    p["cuboid_centers"][4, 0] = 0; p["cuboid_dims"][4, 0] = 3.0          # robot fully inside a box -> always colliding
    traj = np.repeat(_rand_q(tables, 8, 5)[:, None], 4, 1)
    flags, first = engine.sweep_flags(to_dev(p), torch.from_numpy(traj).cuda())
    oflags, ofirst, _ = oracle.sweep_flags(p, traj, tables)
    assert np.array_equal(flags.cpu().numpy(), oflags) and np.array_equal(first.cpu().numpy(), ofirst)
    assert flags[0].item() == 0 and flags[4].item() == 1 and first[4].item() == 0
    # T = 1 and a maximum-length rollout (150 + start, run_inference.py:55)
    for T in (1, 151):
        tr = np.repeat(_rand_q(tables, 8, 6)[:, None], T, 1)
        f, _ = engine.sweep_flags(to_dev(p), torch.from_numpy(tr).cuda())
        assert np.array_equal(f.cpu().numpy(), oracle.sweep_flags(p, tr, tables)[0])


@pytest.mark.parametrize("config", [2, 3, 4])
def test_build_cloud_bit_exact(engine, oracle, tables, config):
    p = _problems(config, 48, problem0=1000)
    cloud = engine.build_cloud(to_dev(p), torch.from_numpy(p["q0"]).cuda(), torch.from_numpy(p["target"]).cuda(), problem0=1000)
    exp = oracle.build_cloud(p["q0"], p["target"], p, tables, engine.cfg.seed, problem0=1000)
    assert cloud.shape == (48, 6272, 4)
    assert np.array_equal(cloud.cpu().numpy(), exp)


def test_build_cloud_empty_scene(engine, oracle, tables):
    p = _problems(2, 3)
    for k in ("cuboid_dims", "cylinder_radii", "cylinder_heights"):
        p[k][1] = 0
    cloud = engine.build_cloud(to_dev(p), torch.from_numpy(p["q0"]).cuda(), torch.from_numpy(p["target"]).cuda())
    assert np.array_equal(cloud.cpu().numpy(), oracle.build_cloud(p["q0"], p["target"], p, tables, engine.cfg.seed))


# ----------------------------------------------------------------------------- pointnet2_ops replacements
def _clouds(engine, oracle, tables, B, config=4):
    p = _problems(config, B)
    return oracle.build_cloud(p["q0"], p["target"], p, tables, engine.cfg.seed), p


def test_fps_bit_exact_on_scene_clouds(engine, oracle, tables):
    cloud, _ = _clouds(engine, oracle, tables, 16)
    idx, new_xyz = engine.fps(torch.from_numpy(cloud).cuda(), 512, return_xyz=True)
    exp = oracle.fps(cloud, 512)
    assert np.array_equal(idx.cpu().numpy(), exp)                       # FPS indices bit-exact
    gathered = np.stack([cloud[b, exp[b], :3] for b in range(16)])
    assert np.array_equal(new_xyz.cpu().numpy(), gathered)
    # second level: 512 -> 128 on xyz-only rows (stride 3)
    idx2 = engine.fps(torch.from_numpy(gathered).cuda(), 128)
    assert np.array_equal(idx2.cpu().numpy(), oracle.fps(gathered, 128))


@pytest.mark.parametrize("N,m", [(1, 1), (5, 3), (31, 31), (33, 8), (100, 64), (129, 100), (300, 300), (512, 128), (512, 512), (513, 200),
                                 (1000, 512), (4097, 64), (6272, 512), (8192, 16)])
def test_fps_bit_exact_sizes_ties_and_skips(engine, oracle, N, m):
    rng = np.random.RandomState(N)
    xyz = rng.uniform(-1, 1, size=(3, N, 3)).astype(np.float32)
    xyz[1] = np.round(xyz[1] * 4) / 4                                   # heavy ties: winner decided by tree order
    xyz[2, : N // 3] *= 0.02                                            # |p|^2 <= 1e-3 skip rule
    for stride in (3, 4):
        a = xyz if stride == 3 else np.concatenate([xyz, np.ones((3, N, 1), np.float32)], -1)
        got = engine.fps(torch.from_numpy(np.ascontiguousarray(a)).cuda(), m).cpu().numpy()
        assert np.array_equal(got, oracle.fps(a, m))


@pytest.mark.parametrize("N,m", [(129, 100), (300, 300), (512, 128), (512, 512)])
def test_fps_warp_per_problem_kernel(engine, oracle, N, m):
    """the warp-per-problem FPS of small clouds (taken for batches >= 1024; forced here with MPN_FPS_WARP=1) on the same tie / skip cases"""
    rng = np.random.RandomState(N + 1)
    xyz = rng.uniform(-1, 1, size=(3, N, 3)).astype(np.float32)
    xyz[1] = np.round(xyz[1] * 4) / 4
    xyz[2, : N // 3] *= 0.02
    os.environ["MPN_FPS_WARP"] = "1"
    try:
        for stride in (3, 4):
            a = xyz if stride == 3 else np.concatenate([xyz, np.ones((3, N, 1), np.float32)], -1)
            got = engine.fps(torch.from_numpy(np.ascontiguousarray(a)).cuda(), m).cpu().numpy()
            assert np.array_equal(got, oracle.fps(a, m))
    finally:
        os.environ.pop("MPN_FPS_WARP", None)


def test_fps_origin_skip_boundary(engine, oracle):
    """|p|^2 exactly 0.001f is selectable, the float below is skipped (upstream compares against the double literal 1e-3)"""
    from test_oracle_cpu import _origin_boundary_cloud
    xyz = _origin_boundary_cloud()
    got = engine.fps(torch.from_numpy(xyz).cuda(), 3).cpu().numpy()
    assert np.array_equal(got, oracle.fps(xyz, 3)) and got[0, 1] == 7
    big = np.concatenate([xyz, np.tile(xyz[:, 1:2], (1, 6272 - 40, 1))], axis=1)      # the pruned large-cloud kernel
    got = engine.fps(torch.from_numpy(np.ascontiguousarray(big)).cuda(), 64).cpu().numpy()
    assert np.array_equal(got, oracle.fps(big, 64)) and got[0, 1] == 7


@pytest.mark.parametrize("variant", ["0", "1", "2", "3", "4", "5"])
def test_fps_kernel_variants_agree(engine, oracle, tables, variant):
    """every variant of the large-cloud FPS (MPN_FPS_VARIANT: thread / points-per-thread splits of the register-resident pruned kernel,
    4 = coordinates in shared memory with two problems per SM, 5 = director warp + workers sleeping on mbarriers) against the oracle
    on scene clouds, a tie-heavy lattice cloud and a cloud with skipped near-origin points: the index sequence is a property of the
    (distance, tie word) total order, not of the reduction shape.  3 is the default (fastest measured: DESIGN.md section 13)."""
    cloud, _ = _clouds(engine, oracle, tables, 6)
    rng = np.random.RandomState(3)
    extra = rng.uniform(-1, 1, size=(2, 6272, 4)).astype(np.float32)
    extra[0, :, :3] = np.round(extra[0, :, :3] * 6) / 6
    extra[1, :2000, :3] *= 0.02
    allc = np.ascontiguousarray(np.concatenate([cloud, extra]))
    os.environ["MPN_FPS_VARIANT"] = variant
    try:
        idx, new_xyz = engine.fps(torch.from_numpy(allc).cuda(), 512, return_xyz=True)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("MPN_FPS_VARIANT", None)
    assert not engine.tc_error()
    exp = oracle.fps(allc, 512)
    assert np.array_equal(idx.cpu().numpy(), exp)
    assert np.array_equal(new_xyz.cpu().numpy(), np.stack([allc[b, exp[b], :3] for b in range(len(allc))]))


def test_ball_query_bit_exact(engine, oracle, tables):
    cloud, _ = _clouds(engine, oracle, tables, 8)
    idx = oracle.fps(cloud, 512)
    new_xyz = np.stack([cloud[b, idx[b], :3] for b in range(8)])
    got = engine.ball_query(0.05, 128, torch.from_numpy(cloud).cuda(), torch.from_numpy(new_xyz).cuda())
    assert np.array_equal(got.cpu().numpy(), oracle.ball_query(0.05, 128, cloud, new_xyz))
    idx2 = oracle.fps(new_xyz, 128)
    nx2 = np.stack([new_xyz[b, idx2[b]] for b in range(8)])
    got = engine.ball_query(0.3, 128, torch.from_numpy(new_xyz).cuda(), torch.from_numpy(nx2).cuda())
    assert np.array_equal(got.cpu().numpy(), oracle.ball_query(0.3, 128, new_xyz, nx2))
    # more than nsample hits, no hits, ragged N
    rng = np.random.RandomState(0)
    xyz = rng.uniform(0, 0.2, size=(2, 777, 3)).astype(np.float32)
    q = np.concatenate([xyz[:, :30], np.full((2, 3, 3), 5.0, np.float32)], 1)
    got = engine.ball_query(0.1, 16, torch.from_numpy(xyz).cuda(), torch.from_numpy(q).cuda())
    assert np.array_equal(got.cpu().numpy(), oracle.ball_query(0.1, 16, xyz, q))


def test_gather_and_group(engine, oracle):
    rng = np.random.RandomState(0)
    feat = rng.normal(size=(3, 5, 200)).astype(np.float32)
    idx = rng.randint(0, 200, size=(3, 40)).astype(np.int32)
    gidx = rng.randint(0, 200, size=(3, 10, 16)).astype(np.int32)
    assert np.array_equal(engine.gather(torch.from_numpy(feat).cuda(), torch.from_numpy(idx).cuda()).cpu().numpy(),
                          oracle.gather_operation(feat, idx))
    assert np.array_equal(engine.group(torch.from_numpy(feat).cuda(), torch.from_numpy(gidx).cuda()).cpu().numpy(),
                          oracle.grouping_operation(feat, gidx))


def test_ops_reject_bad_tensors(engine):
    with pytest.raises(RuntimeError):
        engine.fps(torch.zeros(1, 10, 3), 2)                            # CPU tensor (pointnet2_ops: CUDA only)
    with pytest.raises(RuntimeError):
        engine.fps(torch.zeros(1, 10, 3, dtype=torch.float64).cuda(), 2)
    with pytest.raises(RuntimeError):
        engine.fps(torch.zeros(1, 10, 6).cuda()[:, :, :3], 2)           # non-contiguous
    with pytest.raises(RuntimeError):
        engine.fps(torch.zeros(1, 10, 3).cuda(), 11)                    # npoint > N -> library error surfaces


# ----------------------------------------------------------------------------- set abstraction + network, fp32 mode
def test_sa_modules_fp32(engine_w, oracle, tables, state_dict):
    cloud, _ = _clouds(engine_w, oracle, tables, 4)
    xyz = np.ascontiguousarray(cloud[..., :3])
    feats = torch.from_numpy(np.ascontiguousarray(cloud[..., 3:]))
    d_xyz, d_feats = torch.from_numpy(cloud).cuda(), torch.from_numpy(np.ascontiguousarray(cloud[..., 3:])).cuda()
    for m, spec in enumerate(oracle.SA_SPECS):
        ws = [(state_dict[f"point_cloud_encoder.SA_modules.{m}.mlps.0.{2 * l}.weight"],
               state_dict[f"point_cloud_encoder.SA_modules.{m}.mlps.0.{2 * l}.bias"]) for l in range(3)]
        o_xyz, o_feats, aux = oracle.sa_module(xyz, feats, spec, ws, return_aux=True)
        res = engine_w.sa_forward(m, d_xyz, d_feats, debug=True)
        if m < 2:
            assert np.array_equal(res[2].cpu().numpy(), aux["fps_idx"])
            assert np.array_equal(res[3].cpu().numpy(), aux["ball_idx"])
            assert np.array_equal(res[0].cpu().numpy(), o_xyz)
        got = res[1].cpu()
        scale = o_feats.abs().max().item()
        assert (got - o_feats).abs().max().item() <= 2e-6 * max(1.0, scale) + 2e-6
        xyz, feats = o_xyz, o_feats
        d_xyz = torch.from_numpy(o_xyz).cuda() if o_xyz is not None else None
        d_feats = o_feats.contiguous().cuda()


def test_policy_forward_fp32_within_1e5(engine_w, oracle, tables, state_dict):
    cloud, p = _clouds(engine_w, oracle, tables, 6)
    qn = oracle.normalize(p["q0"], tables.joint_limits)
    dq = engine_w.policy_forward(torch.from_numpy(cloud).cuda(), torch.from_numpy(qn).cuda()).cpu()
    exp = oracle.policy_forward(state_dict, cloud, qn)
    exp64 = oracle.policy_forward(state_dict, cloud, qn, dtype=torch.float64).float()
    err = (dq - exp).abs().max().item()
    print("delta-q max-abs-err vs fp32 oracle:", err, " vs fp64 shadow:", (dq - exp64).abs().max().item(),
          " |dq| max:", exp.abs().max().item())
    assert err <= DQ_TOL
    enc = engine_w.encoder_forward(torch.from_numpy(cloud).cuda()).cpu()
    oenc = oracle.encoder_forward(state_dict, cloud)
    assert (enc - oenc).abs().max().item() <= 1e-5 * max(1.0, oenc.abs().max().item())


def test_rollout_fp32(engine_w, oracle, tables, state_dict):
    """TrainingMotionPolicyNetwork.rollout + validation sweep (model.py:128-183,293-314), 3 lock-step steps."""
    T = 3
    p = _problems(4, 6)
    sc = to_dev(p)
    q0, tg = torch.from_numpy(p["q0"]).cuda(), torch.from_numpy(p["target"]).cuda()
    cloud = engine_w.build_cloud(sc, q0, tg)
    ocloud = cloud.cpu().numpy().copy()
    traj, metrics = engine_w.rollout(sc, cloud, q0, tg, T)
    torch.cuda.synchronize()
    otraj = oracle.rollout(state_dict, ocloud, oracle.normalize(p["q0"], tables.joint_limits), tables, T, engine_w.cfg.seed)
    traj_h = traj.cpu().numpy()
    assert np.array_equal(traj_h[:, 0], p["q0"])
    # step 1 starts from identical inputs: 1e-5 in normalised units -> scaled by the joint range after unnormalising
    rng_ = (tables.joint_limits[:, 1] - tables.joint_limits[:, 0]) / 2
    assert (np.abs(traj_h[:, 1] - otraj[:, 1]) / rng_).max() <= DQ_TOL
    assert (np.abs(traj_h - otraj) / rng_).max() <= 1e-3               # later steps: bounded drift (FPS picks may differ)
    # the cloud was updated in place with the robot at the last configuration (model.py:181)
    exp_rows = oracle.sample_robot(traj_h[:, -1], tables, 2048, engine_w.cfg.seed, T)
    assert np.array_equal(cloud[:, :2048].cpu().numpy(), exp_rows)
    assert np.array_equal(cloud[:, 2048:].cpu().numpy(), ocloud[:, 2048:])   # obstacle / target rows untouched
    # collision flags of the GPU trajectory: bit-exact against the oracle sweep of the same trajectory
    oflags, ofirst, _ = oracle.sweep_flags(p, traj_h, tables)
    m = metrics.cpu().numpy()
    assert np.array_equal(m[:, 0].astype(np.uint8), oflags)
    assert np.array_equal(m[:, 1].astype(np.int32), ofirst)
    assert (m[:, 2] == T).all()
    # final position error column = |FK(q_T).xyz - target.xyz|
    _, eef = oracle.fk(traj_h[:, -1])
    assert np.abs(m[:, 3] - np.linalg.norm(eef[:, :, 3] - p["target"][:, :, 3], axis=1)).max() < 1e-5
    # per-step checking (config 3) gives the same flags
    cloud2 = engine_w.build_cloud(sc, q0, tg)
    traj2, metrics2 = engine_w.rollout(sc, cloud2, q0, tg, T, check_every_step=True)
    assert torch.equal(traj2, traj) and torch.equal(metrics2[:, :2], metrics[:, :2])


def test_rollout_early_exit_mask(engine_w, oracle, tables):
    """run_inference.rollout_until_success (run_inference.py:171-187): a problem whose start already satisfies the
    1 cm / 15 deg test stops after its first step and is frozen afterwards."""
    p = _problems(2, 4)
    # make problem 0's target equal to where the first step lands: run one step, then use its pose as the target
    sc = to_dev(p)
    q0, tg = torch.from_numpy(p["q0"]).cuda(), torch.from_numpy(p["target"]).cuda()
    cloud = engine_w.build_cloud(sc, q0, tg)
    traj, _ = engine_w.rollout(sc, cloud.clone(), q0, tg, 1)
    _, eef = engine_w.fk(traj[:, 1].contiguous())
    tg2 = tg.clone(); tg2[0] = eef[0]
    cloud = engine_w.build_cloud(sc, q0, tg)    # same cloud (target rows differ from tg2 on purpose: same policy output)
    traj3, metrics = engine_w.rollout(sc, cloud, q0, tg2, 3, early_exit=True)
    m = metrics.cpu().numpy()
    assert m[0, 2] == 1 and m[0, 5] == 1                                 # stopped at step 1, reached
    assert torch.equal(traj3[0, 1], traj3[0, 2]) and torch.equal(traj3[0, 2], traj3[0, 3])
    assert (m[1:, 2] == 3).all()


def test_rollout_early_exit_stops_launching(engine_w, oracle, tables):
    """rollout_until_success breaks out of its loop (run_inference.py:180-187): once EVERY problem of the batch has stopped, the
    library stops enqueuing steps (host poll every 8 steps) and the remaining trajectory rows repeat the frozen configurations"""
    p = _problems(2, 3)
    sc = to_dev(p)
    q0, tg = torch.from_numpy(p["q0"]).cuda(), torch.from_numpy(p["target"]).cuda()
    cloud = engine_w.build_cloud(sc, q0, tg)
    traj1, _ = engine_w.rollout(sc, cloud.clone(), q0, tg, 1)
    _, eef = engine_w.fk(traj1[:, 1].contiguous())                 # every target = where the first step lands
    T = 40
    l0 = engine_w.launch_count
    traj, metrics = engine_w.rollout(sc, cloud.clone(), q0, eef.contiguous(), T, early_exit=True)
    per_step = (engine_w.launch_count - l0)
    l1 = engine_w.launch_count
    traj_full, metrics_full = engine_w.rollout(sc, cloud.clone(), q0, eef.contiguous(), T, early_exit=2)   # done mask only: all T steps
    full = engine_w.launch_count - l1
    assert per_step < 0.35 * full                                   # 8 of 40 steps were enqueued
    m = metrics.cpu().numpy()
    assert (m[:, 2] == 1).all() and (m[:, 5] == 1).all()
    assert torch.equal(traj, traj_full) and torch.equal(metrics, metrics_full)
    assert all(torch.equal(traj[:, t], traj[:, 1]) for t in range(2, T + 1))


def test_rollout_early_exit_compacts_live_problems(engine_w, oracle, tables):
    """rollout_until_success on a batch (run_inference.py:137-191): problems stop at different steps; once at most half of the current
    set is still running the library carries only the running ones on (state gathered into a compact set, results scattered back).
    120 of 160 problems stop at step 1, 20 more at step 12, 20 never: compaction at the polls of step 8 (160 -> 40) and 16 (40 -> 20).
    Trajectories, metrics and the final clouds must equal the uncompacted rollout (MPN_NO_LIVE_COMPACTION=1) bit for bit, with the
    per-step collision check on.  (The sets stay above 16 problems: at <= 16 the FC head switches to its fp32 weight-streaming
    kernels, DESIGN.md section 9c, and the two runs would differ by that kernel's rounding.)"""
    from mpinets_b200 import _lib
    B, T = 160, 30
    p = _problems(4, B)
    sc = to_dev(p)
    q0, tg = torch.from_numpy(p["q0"]).cuda(), torch.from_numpy(p["target"]).cuda()
    cloud0 = engine_w.build_cloud(sc, q0, tg)
    ref_traj, _ = engine_w.rollout(sc, cloud0.clone(), q0, tg, T, precision=_lib.PREC_BF16X3)      # no early exit: where every step lands
    _, eef1 = engine_w.fk(ref_traj[:, 1].contiguous())
    _, eef12 = engine_w.fk(ref_traj[:, 12].contiguous())
    tg2 = tg.clone()
    tg2[:120] = eef1[:120]          # reached after the first step
    tg2[120:140] = eef12[120:140]   # reached at step 12 (unless the arm passed through that pose before)
    outs = []
    for nocompact in (True, False):
        if nocompact:
            os.environ["MPN_NO_LIVE_COMPACTION"] = "1"
        try:
            c = cloud0.clone()
            traj, metrics = engine_w.rollout(sc, c, q0, tg2, T, early_exit=True, check_every_step=True, precision=_lib.PREC_BF16X3)
            torch.cuda.synchronize()
        finally:
            os.environ.pop("MPN_NO_LIVE_COMPACTION", None)
        assert not engine_w.tc_error()
        outs.append((traj.clone(), metrics.clone(), c))
    (t0, m0, c0), (t1, m1, c1) = outs
    steps = m0[:, 2].cpu().numpy()
    assert (steps[:120] == 1).all() and (steps[120:140] <= 12).all() and (steps[140:] == T).all()
    print("compacted vs uncompacted: max |traj diff|", float((t0 - t1).abs().max()), " metrics", float((m0 - m1).abs().max()),
          " cloud", float((c0 - c1).abs().max()))
    # a stopped problem's cloud keeps the robot rows of the last step it was carried through (the subset of surface points changes per
    # step even for a frozen pose), so the in-place cloud is compared for the problems that run to the end
    assert torch.equal(t0, t1) and torch.equal(m0, m1) and torch.equal(c0[140:], c1[140:])
    assert torch.equal(t1[140:], ref_traj[140:])                      # the problems that never stop follow the plain rollout


# ----------------------------------------------------------------------------- full-size properties (BASELINE configs)
def test_full_size_properties(engine, oracle, tables):
    """4096 problems (configs[1]): size-independent properties of the GPU path + spot parity on a subset."""
    B = 4096
    p = _problems(2, B)
    sc = to_dev(p)
    q0, tg = torch.from_numpy(p["q0"]).cuda(), torch.from_numpy(p["target"]).cuda()
    cloud = engine.build_cloud(sc, q0, tg)
    assert cloud.shape == (B, 6272, 4)
    c = cloud.cpu().numpy()
    assert (c[:, :2048, 3] == 0).all() and (c[:, 2048:6144, 3] == 1).all() and (c[:, 6144:, 3] == 2).all()
    assert np.isfinite(c).all()
    sub = np.arange(0, B, 257)
    for i, b in enumerate(sub):   # problem index enters the RNG counter
        e1 = oracle.build_cloud(p["q0"][b:b + 1], p["target"][b:b + 1], {k: v[b:b + 1] for k, v in p.items()}, tables,
                                engine.cfg.seed, problem0=int(b))
        assert np.array_equal(c[b], e1[0])
    # obstacle points lie on the scene surface: scene sdf <= ~0 (inside another primitive allowed) and never far outside
    sdf = engine.sdf_points(sc, cloud[:, 2048:6144, :3].contiguous()).cpu().numpy()
    assert sdf.max() < 1e-5
    # FPS: indices unique, start at 0, idempotent on re-run (determinism)
    idx = engine.fps(cloud, 512)
    idx_again = engine.fps(cloud, 512)
    assert torch.equal(idx, idx_again)
    ih = idx.cpu().numpy()
    assert (ih[:, 0] == 0).all()
    assert all(len(set(r.tolist())) == 512 for r in ih[::64])
    assert np.array_equal(ih[sub], oracle.fps(c[sub], 512))


# ----------------------------------------------------------------------------- tensor-core self-test
TC_MODE = int(os.environ.get("MPN_TC_MODE", "0"))   # smem-descriptor convention used by sa_tc.cu


@pytest.mark.parametrize("N,K", [(64, 64), (128, 128), (256, 128), (128, 256), (32, 64)])
def test_tcgen05_selftest_gemm(engine, N, K):
    """single-CTA tcgen05.mma GEMM (the descriptor / TMEM conventions of the fused kernels) vs torch fp32"""
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + K)
    a = torch.randn(128, K, generator=g).to(torch.bfloat16).cuda()
    b = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
    ref = a.float() @ b.float().t()
    report = {}
    for mode in range(4):
        d, timeout = engine.tc_selftest(a, b, mode)
        report[mode] = ("timeout" if timeout else float((d - ref).abs().max().item()))
    print(f"tcgen05 selftest N={N} K={K}: max-abs-err per descriptor mode: {report}")
    assert report[TC_MODE] != "timeout" and report[TC_MODE] < 1e-2 * K ** 0.5


@pytest.mark.parametrize("N,K", [(64, 64), (128, 128), (64, 256)])
def test_tcgen05_selftest_a_from_tmem(engine, N, K):
    """the same GEMM with the A operand written to tensor memory by tcgen05.st (row = lane, two bf16 per column) and read by
    tcgen05.mma from there (selftest mode 4) -- the operand path of kernels that keep activations in TMEM between layers"""
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + K + 7)
    a = torch.randn(128, K, generator=g).to(torch.bfloat16).cuda()
    b = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
    ref = a.float() @ b.float().t()
    d, timeout = engine.tc_selftest(a, b, 4)
    err = float((d - ref).abs().max().item())
    print(f"tcgen05 selftest (A from TMEM) N={N} K={K}: timeout={timeout} max-abs-err {err:.3e}")
    assert not timeout and err < 1e-2 * K ** 0.5


# ----------------------------------------------------------------------------- bf16 tensor-core mode
# Tolerances of the throughput mode, against the oracle run with the SAME bf16 operand rounding (fp64 accumulate):
# remaining differences are fp32-vs-fp64 accumulation order, which can flip a bf16 rounding of an intermediate
# activation by one ulp (2^-8 relative).
BF16_FEAT_RTOL = 2e-2


def _sa_weights(sd, m):
    return [(sd[f"point_cloud_encoder.SA_modules.{m}.mlps.0.{2 * l}.weight"], sd[f"point_cloud_encoder.SA_modules.{m}.mlps.0.{2 * l}.bias"])
            for l in range(3)]


def test_sa_modules_bf16_tensor_core(engine_w, oracle, tables, state_dict):
    from mpinets_b200 import _lib
    cloud, _ = _clouds(engine_w, oracle, tables, 5)
    d_cloud = torch.from_numpy(cloud).cuda()
    xyz = np.ascontiguousarray(cloud[..., :3])
    feats = torch.from_numpy(np.ascontiguousarray(cloud[..., 3:]))
    o_xyz1, o_f1, aux1 = oracle.sa_module(xyz, feats, oracle.SA_SPECS[0], _sa_weights(state_dict, 0), emulate_bf16=True,
                                          dtype=torch.float64, return_aux=True)
    nx, f1, fi1, bi1 = engine_w.sa_forward(0, d_cloud, d_cloud[..., 3:], precision=_lib.PREC_BF16, debug=True)
    assert not engine_w.tc_error()
    assert np.array_equal(nx.cpu().numpy(), o_xyz1)
    assert np.array_equal(fi1.cpu().numpy(), aux1["fps_idx"])
    assert np.array_equal(bi1.cpu().numpy(), aux1["ball_idx"])          # hash-grid ball query: index lists bit-exact
    e1 = (f1.cpu().double() - o_f1).abs().max().item() / o_f1.abs().max().item()
    print("SA1 bf16 tensor-core vs bf16-emulating oracle: rel err", e1)
    assert e1 < BF16_FEAT_RTOL
    # SA2 fed with the oracle's (bf16-representable) SA1 output
    o_xyz2, o_f2, aux2 = oracle.sa_module(o_xyz1, o_f1.float(), oracle.SA_SPECS[1], _sa_weights(state_dict, 1), emulate_bf16=True,
                                          dtype=torch.float64, return_aux=True)
    nx2, f2, fi2, bi2 = engine_w.sa_forward(1, torch.from_numpy(o_xyz1).cuda(), o_f1.float().contiguous().cuda(),
                                            precision=_lib.PREC_BF16, debug=True)
    assert not engine_w.tc_error()
    assert np.array_equal(nx2.cpu().numpy(), o_xyz2)
    assert np.array_equal(bi2.cpu().numpy(), aux2["ball_idx"])
    e2 = (f2.cpu().double() - o_f2).abs().max().item() / o_f2.abs().max().item()
    print("SA2 bf16 tensor-core vs bf16-emulating oracle: rel err", e2)
    assert e2 < BF16_FEAT_RTOL
    # and against the fp32 path of the same module (sanity of the whole mode, not a parity bar)
    _, f1_32 = engine_w.sa_forward(0, d_cloud, torch.from_numpy(np.ascontiguousarray(cloud[..., 3:])).cuda())
    print("SA1 bf16 vs fp32 path: rel err", ((f1 - f1_32).abs().max() / f1_32.abs().max()).item())


def test_sa1_shared_memory_operand_kernel(engine_w, oracle, tables, state_dict):
    """the predecessor / fallback SA1 kernel (operands in shared memory, MPN_SA1_SS=1; also taken when a cloud does not fit in
    shared memory next to the weights) against the same oracle, and against the default kernel (activations in TMEM): identical
    ball-query indices, features equal up to the accumulation order inside the tensor core"""
    from mpinets_b200 import _lib
    cloud, _ = _clouds(engine_w, oracle, tables, 3)
    d_cloud = torch.from_numpy(cloud).cuda()
    xyz = np.ascontiguousarray(cloud[..., :3])
    feats = torch.from_numpy(np.ascontiguousarray(cloud[..., 3:]))
    _, o_f1, aux1 = oracle.sa_module(xyz, feats, oracle.SA_SPECS[0], _sa_weights(state_dict, 0), emulate_bf16=True, dtype=torch.float64,
                                     return_aux=True)
    _, f_t, _, bi_t = engine_w.sa_forward(0, d_cloud, d_cloud[..., 3:], precision=_lib.PREC_BF16, debug=True)
    os.environ["MPN_SA1_SS"] = "1"
    try:
        _, f_s, _, bi_s = engine_w.sa_forward(0, d_cloud, d_cloud[..., 3:], precision=_lib.PREC_BF16, debug=True)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("MPN_SA1_SS", None)
    assert not engine_w.tc_error()
    assert np.array_equal(bi_s.cpu().numpy(), aux1["ball_idx"]) and torch.equal(bi_s, bi_t)
    scale = o_f1.abs().max().item()
    assert (f_s.cpu().double() - o_f1).abs().max().item() / scale < BF16_FEAT_RTOL
    assert (f_s - f_t).abs().max().item() / scale < 1e-2


def test_policy_forward_bf16(engine_w, oracle, tables, state_dict):
    from mpinets_b200 import _lib
    cloud, p = _clouds(engine_w, oracle, tables, 6)
    qn = oracle.normalize(p["q0"], tables.joint_limits)
    dq = engine_w.policy_forward(torch.from_numpy(cloud).cuda(), torch.from_numpy(qn).cuda(), _lib.PREC_BF16).cpu()
    assert not engine_w.tc_error()
    exp32 = oracle.policy_forward(state_dict, cloud, qn)
    expbf = oracle.policy_forward(state_dict, cloud, qn, emulate_bf16=True, dtype=torch.float64).float()
    print("bf16 mode delta-q: max-abs-err vs bf16-emulating oracle", (dq - expbf).abs().max().item(),
          " vs fp32 oracle", (dq - exp32).abs().max().item(), " |dq| max", exp32.abs().max().item())
    assert (dq - exp32).abs().max().item() < 2e-2 * max(1.0, exp32.abs().max().item())


def test_tensor_core_ball_query_dense_and_fallback(engine_w, oracle):
    """SA1's hash-grid ball query against pointnet2 semantics on adversarial clouds: > 128 hits (first-128-in-index-order
    selection by rank), > 512 hits (linear-scan fallback), points exactly on cell boundaries, far-away outliers."""
    from mpinets_b200 import _lib
    rng = np.random.RandomState(7)
    N = 6272
    cloud = np.zeros((3, N, 4), np.float32)
    cloud[..., :3] = rng.uniform(-1.0, 1.5, size=(3, N, 3))
    # FPS always starts at row 0, so clusters that contain row 0 are guaranteed to hold a centroid
    cloud[0, 0:300, :3] = 0.5 + rng.uniform(-0.02, 0.02, size=(300, 3))        # 300 points in a 4 cm cube  (> 128 hits)
    cloud[1, 0:1500, :3] = -0.3 + rng.uniform(-0.015, 0.015, size=(1500, 3))    # 1500 points in a 3 cm cube (> 512 hits)
    cloud[2, :, :3] = np.round(cloud[2, :, :3] / 0.0501) * 0.0501               # everything on grid-cell corners
    cloud[2, 5, :3] = [7.5, -7.9, 30.0]                                          # outliers far outside the scene
    cloud[..., 3] = rng.randint(0, 3, size=(3, N))
    idx = oracle.fps(cloud, 512)
    new_xyz = np.stack([cloud[b, idx[b], :3] for b in range(3)])
    exp = oracle.ball_query(0.05, 128, cloud, new_xyz)
    d = torch.from_numpy(cloud).cuda()
    nx, f1, fi, bi = engine_w.sa_forward(0, d, d[..., 3:], precision=_lib.PREC_BF16, debug=True)
    assert not engine_w.tc_error()
    assert np.array_equal(fi.cpu().numpy(), idx)
    assert np.array_equal(bi.cpu().numpy(), exp)
    counts = np.array([[len(set(r.tolist())) for r in exp[b]] for b in range(3)])
    assert counts[0].max() == 128 and counts[1].max() == 128                      # the dense cases were exercised


# ----------------------------------------------------------------------------- Evaluator subset (metrics.py:311-523)
def _eval_case(B=24, T1=21, seed=5):
    p = _problems(4, B)
    w = np.linspace(0.0, 1.0, T1, dtype=np.float32)[None, :, None]
    traj = (p["q0"][:, None, :] * (1 - w) + p["q_goal"][:, None, :] * w).astype(np.float32)
    rng = np.random.default_rng(seed)
    traj[1, -1] += 0.2
    traj[2, 5, 3] = 0.5
    traj[3, :, :] = np.array([1.638, 1.227, 0.041, -3.039, 0.047, 1.604, 0.314], np.float32) + 0.01 * rng.standard_normal((T1, 7)).astype(np.float32)
    num = np.full(B, T1, np.int32); num[4] = 7; num[7] = 1
    unit = np.tile(np.array([1, 0, 0, 0], np.float32), (B, 1, 1))
    yaw = rng.uniform(-np.pi, np.pi, B)
    quat = np.stack([np.cos(yaw / 2), 0 * yaw, 0 * yaw, np.sin(yaw / 2)], -1).astype(np.float32)[:, None]
    tv = dict(cuboid_centers=p["target"][:, None, :3, 3].copy(), cuboid_dims=np.full((B, 1, 3), 0.1, np.float32), cuboid_quats=quat)
    tv["cuboid_centers"][5, 0, 0] += 1.0
    nv = dict(cuboid_centers=np.repeat(p["target"][:, None, :3, 3], 2, axis=1) + np.array([[0.3, 0, 0], [0, 0.02, 0]], np.float32),
              cuboid_dims=np.full((B, 2, 3), 0.2, np.float32), cuboid_quats=np.repeat(unit, 2, axis=1),
              cylinder_centers=p["target"][:, None, :3, 3] + np.array([0, 0, 0.12], np.float32),
              cylinder_radii=np.full((B, 1, 1), 0.05, np.float32), cylinder_heights=np.full((B, 1, 1), 0.1, np.float32),
              cylinder_quats=unit.copy())
    nv["cuboid_dims"][6, 0] = 1.0
    nv["cuboid_dims"][::2, 1] = 0.0       # padding rows
    nv["cylinder_heights"][8] = 0.3       # now contains the target itself -> dropped (metrics.py:497-499)
    from mpinets_b200.franka import fk_reference_f64
    nv["cuboid_centers"][1, 0] = fk_reference_f64(traj[1, -1].astype(np.float64))[1][:3, 3]   # final pose of the miss
    nv["cuboid_dims"][1, 0] = 0.05        # -> final xyz inside a negative volume that does not contain the target
    return p, np.ascontiguousarray(traj), num, tv, nv


def test_evaluate_matches_oracle(engine, oracle, tables):
    p, traj, num, tv, nv = _eval_case()
    sc = to_dev(p)
    dev = lambda d: {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in d.items()}  # noqa: E731
    tg = torch.from_numpy(p["target"]).cuda()
    for kw_g, kw_o in (
        (dict(num_poses=torch.from_numpy(num).cuda(), target_volume=dev(tv), negative_volumes=dev(nv)),
         dict(num_poses=num, target_volume=tv, negative_volumes=nv)),
        (dict(), dict()),
        (dict(target_volume=dev(tv)), dict(target_volume=tv)),
        (dict(negative_volumes=dev(nv)), dict(negative_volumes=nv)),
    ):
        got = engine.evaluate(sc, torch.from_numpy(traj).cuda(), tg, **kw_g).cpu().numpy()
        exp = oracle.evaluate(p, traj, p["target"], tables, **kw_o)
        flag_cols = [0, 1, 2, 3, 8, 9, 10, 11]
        assert np.array_equal(got[:, flag_cols], exp[:, flag_cols])          # flags / counts: bit-exact
        assert np.array_equal(got[:, [4, 6, 12, 13]], exp[:, [4, 6, 12, 13]])  # spec-arithmetic lengths: bit-exact
        assert np.abs(got[:, [5, 7]] - exp[:, [5, 7]]).max() < 1e-3            # atan2f: library ulps
        assert (got[:, 14:] == 0).all()
    assert got[1, 8] == 0 and got[0, 8] == 1 and got[8, 8] == 1   # negative-volume-only run: row 1 ends in the wrong region
    assert got.shape == (traj.shape[0], 16)


def test_evaluate_after_rollout(engine_w, oracle, tables):
    """mpn_rollout's trajectory buffer feeds mpn_evaluate directly; its collision / position columns agree with the
    rollout's own metrics table."""
    T = 3
    p = _problems(3, 8)
    sc = to_dev(p)
    q0, tg = torch.from_numpy(p["q0"]).cuda(), torch.from_numpy(p["target"]).cuda()
    cloud = engine_w.build_cloud(sc, q0, tg)
    traj, metrics = engine_w.rollout(sc, cloud, q0, tg, T)
    ev = engine_w.evaluate(sc, traj, tg).cpu().numpy()
    m = metrics.cpu().numpy()
    assert np.array_equal(ev[:, 0], m[:, 0]) and np.array_equal(ev[:, 11], m[:, 1])
    assert np.abs(ev[:, 4] - 100 * m[:, 3]).max() < 1e-3
    assert np.abs(ev[:, 5] - m[:, 4]).max() < 0.1
    assert (ev[:, 10] == T + 1).all()
    exp = oracle.evaluate(p, traj.cpu().numpy(), p["target"], tables)
    assert np.array_equal(ev[:, [0, 1, 2, 3, 8, 9, 10, 11]], exp[:, [0, 1, 2, 3, 8, 9, 10, 11]])


# ----------------------------------------------------------------------------- losses (loss.py:31-166)
def test_collision_and_point_match_loss_match_reference_fixture(engine, oracle):
    """CUDA collision_loss / point_match_loss against the values and autograd gradients of the REAL mpinets/loss.py
    (tests/golden/loss_reference.npz) and against the oracle"""
    from mpinets_b200 import scenes
    from mpinets_b200.engine import Engine
    g = np.load(os.path.join(HERE, "golden", "loss_reference.npz"))
    M1, M2 = g["yaw_cuboid_centers"].shape[1], g["yaw_cylinder_centers"].shape[1]
    eng = Engine(max_cuboids=M1, max_cylinders=M2)       # the fixture's scenes have 6 + 4 primitive rows
    for tag in ("yaw", "free"):
        sc = {k: g[f"{tag}_{k}"] for k in scenes.SCENE_KEYS}
        pts = g[f"{tag}_points"]
        loss, grad = eng.collision_loss(to_dev(sc), torch.from_numpy(pts).cuda(), 0.03, need_grad=True)
        ref_g = g[f"{tag}_collision_grad"]
        assert abs(float(loss) - float(g[f"{tag}_collision_loss"])) < 5e-7
        gh = grad.cpu().numpy()
        assert np.array_equal(gh != 0, ref_g != 0)                       # same points inside the margin
        assert np.abs(gh - ref_g).max() < 1e-5 * np.abs(ref_g).max()
        oval, ograd = oracle.collision_loss(sc, pts)
        assert abs(float(loss) - oval) < 5e-7 and np.abs(gh - ograd).max() < 1e-5 * np.abs(ograd).max()
        loss2, _ = eng.collision_loss(to_dev(sc), torch.from_numpy(pts).cuda(), 0.03, need_grad=False)
        assert torch.equal(loss, loss2)                                  # deterministic reduction
        pl, pg = eng.point_match_loss(torch.from_numpy(pts).cuda(), torch.from_numpy(g[f"{tag}_other"]).cuda(), need_grad=True)
        assert abs(float(pl) - float(g[f"{tag}_point_match_loss"])) < 1e-6
        assert np.abs(pg.cpu().numpy() - g[f"{tag}_point_match_grad"]).max() < 1e-9


def test_bc_collision_losses_match_oracle(engine, oracle, tables):
    """CollisionAndBCLossContainer.__call__ (loss.py:111-166) on config-4 scenes: both losses and the joint-space gradient"""
    B = 48
    p = _problems(4, B)
    rng = np.random.default_rng(3)
    qi = rng.uniform(-0.9, 0.9, (B, 7)).astype(np.float32)
    qi[: B // 2] = oracle.normalize(p["q0"][: B // 2], tables.joint_limits)      # half the batch at the scene's own start poses
    qt = np.clip(qi + rng.normal(scale=0.05, size=qi.shape), -1, 1).astype(np.float32)
    for wc, wb in ((5.0, 1.0), (1.0, 0.0), (0.0, 1.0)):
        losses, grad = engine.bc_collision_losses(to_dev(p), torch.from_numpy(qi).cuda(), torch.from_numpy(qt).cuda(), 1024, 0.03, wc, wb,
                                                  need_grad=True)
        ol, og = oracle.bc_collision_losses(p, qi, qt, tables, engine.cfg.seed, 1024, 0.03, wc, wb)
        assert np.abs(losses.cpu().numpy() - ol).max() < 1e-6 * max(1.0, float(np.abs(ol).max())) + 2e-7
        gh = grad.cpu().numpy()
        assert np.abs(gh - og).max() < 2e-5 * np.abs(og).max() + 1e-8
    # the fixed cloud the kernel uses is the documented subset: losses equal the standalone kernels on oracle-built clouds
    xi, _ = oracle.fixed_robot_points(oracle.unnormalize(qi, tables.joint_limits), tables, 1024, engine.cfg.seed)
    xt, _ = oracle.fixed_robot_points(oracle.unnormalize(qt, tables.joint_limits), tables, 1024, engine.cfg.seed)
    l_c, _ = engine.collision_loss(to_dev(p), torch.from_numpy(xi).cuda())
    l_p, _ = engine.point_match_loss(torch.from_numpy(xi).cuda(), torch.from_numpy(xt).cuda())
    l2, _ = engine.bc_collision_losses(to_dev(p), torch.from_numpy(qi).cuda(), torch.from_numpy(qt).cuda())
    assert abs(float(l_c) - float(l2[0])) < 1e-6 and abs(float(l_p) - float(l2[1])) < 1e-6


def test_sparc_matches_reference_fixture(engine, oracle):
    """mpn_sparc against outputs of the REAL third_party/sparc.py (tests/golden/sparc_reference.npz) and the oracle"""
    g = np.load(os.path.join(HERE, "golden", "sparc_reference.npz"))
    got = engine.sparc(torch.from_numpy(g["profiles"]).cuda(), float(g["fs"]), torch.from_numpy(g["num"]).cuda()).cpu().numpy()
    assert np.abs(got - g["sal"]).max() < 2e-4                       # fp32 DFT vs float64 numpy FFT
    doc = g["doctest_move"].astype(np.float32)[None]                  # sparc.py:87-91
    sal = engine.sparc(torch.from_numpy(np.ascontiguousarray(doc)).cuda(), 100.0)
    assert "%.4f" % float(sal[0]) == "-1.4140"
    rng = np.random.default_rng(1)
    prof = np.abs(rng.normal(size=(9, 70))).astype(np.float32); prof[3] = 0.0
    got = engine.sparc(torch.from_numpy(prof).cuda(), 12.5).cpu().numpy()
    exp = np.array([oracle.sparc(prof[b], 12.5) for b in range(9)])
    assert got[3] == 0.0 and np.abs(got - exp).max() < 5e-4


def test_build_cloud_from_obstacle_points_bit_exact(engine, oracle, tables):
    """mpn_build_cloud_from_points (run_inference.make_point_cloud_from_problem, run_inference.py:58-90)"""
    p = _problems(2, 5)
    rng = np.random.default_rng(0)
    counts = np.array([9000, 4096, 5000, 4097, 8191], np.int32)
    pts = rng.uniform(-1, 1, (5, 9000, 3)).astype(np.float32)
    got = engine.build_cloud_from_points(torch.from_numpy(p["q0"]).cuda(), torch.from_numpy(p["target"]).cuda(),
                                         torch.from_numpy(pts).cuda(), torch.from_numpy(counts).cuda(), problem0=11).cpu().numpy()
    exp = oracle.build_cloud_from_points(p["q0"], p["target"], pts, counts, tables, engine.cfg.seed, problem0=11)
    assert np.array_equal(got, exp)


def test_depth_render_bit_exact_and_feeds_cloud_build(engine, oracle, tables):
    """mpn_render_depth_cloud (run_inference.convert_primitive_problems_to_depth stand-in) vs the oracle ray caster: counts
    and hit coordinates bit-exact, shared and per-problem cameras; the result feeds make_point_cloud_from_problem"""
    from mpinets_b200.run_inference import eval_camera
    B, W, H = 6, 96, 72
    for config, env in ((2, "tabletop"), (3, "dresser")):
        p = _problems(config, B)
        cam = eval_camera(env)
        pts, cnt = engine.render_depth_cloud(to_dev(p), torch.from_numpy(cam).cuda(), W, H)
        opts, ocnt = oracle.render_depth_cloud(p, cam, W, H)
        assert np.array_equal(cnt.cpu().numpy(), ocnt)
        got = pts.cpu().numpy()
        for b in range(B):
            assert np.array_equal(got[b, :ocnt[b]], opts[b, :ocnt[b]])
    cams = np.stack([cam] * B)
    cams[:, :, 3] += np.linspace(0, 0.2, B, dtype=np.float32)[:, None]
    pts2, cnt2 = engine.render_depth_cloud(to_dev(p), torch.from_numpy(cams).cuda(), W, H)
    opts2, ocnt2 = oracle.render_depth_cloud(p, cams, W, H)
    assert np.array_equal(cnt2.cpu().numpy(), ocnt2)
    assert np.array_equal(pts2.cpu().numpy()[1, :ocnt2[1]], opts2[1, :ocnt2[1]])
    # depth clouds with >= 4096 points feed the cloud builder (run_inference.py:58-90)
    big, bcnt = engine.render_depth_cloud(to_dev(p), torch.from_numpy(cam).cuda(), 320, 240)
    assert int(bcnt.min()) >= 4096
    q0, tg = torch.from_numpy(p["q0"]).cuda(), torch.from_numpy(np.ascontiguousarray(p["target"])).cuda()
    cloud = engine.build_cloud_from_points(q0, tg, big, bcnt)
    obig, obcnt = oracle.render_depth_cloud(p, cam, 320, 240)
    ocloud = oracle.build_cloud_from_points(p["q0"], p["target"], obig, obcnt, tables, engine.cfg.seed)
    assert np.array_equal(cloud.cpu().numpy(), ocloud)
