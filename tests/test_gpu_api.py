"""GPU tests (-m gpu) of the reference-facing Python surface: the same call patterns the reference uses
(`/root/reference/mpinets/model.py`, `run_inference.py`, `loss.py`) against the shim modules in `mpinets_b200/`."""
import numpy as np
import pytest
import torch

from conftest import to_dev

pytestmark = pytest.mark.gpu


def _problems(B, config=4):
    from mpinets_b200 import scenes
    return scenes.config_problems(config, B)


def test_pointnet2_utils_surface(oracle):
    """model.py:27 imports PointnetSAModule from pointnet2_ops; its glue calls these functions with these layouts."""
    from mpinets_b200 import pointnet2_utils as pu
    rng = np.random.RandomState(0)
    xyz = torch.from_numpy((rng.uniform(-1, 1, size=(2, 900, 3)) + 2).astype(np.float32)).cuda()
    feats = torch.from_numpy(rng.normal(size=(2, 5, 900)).astype(np.float32)).cuda()
    idx = pu.furthest_point_sample(xyz, 64)
    assert idx.dtype == torch.int32 and idx.shape == (2, 64)
    assert np.array_equal(idx.cpu().numpy(), oracle.fps(xyz.cpu().numpy(), 64))
    new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()   # _PointnetSAModuleBase.forward
    assert new_xyz.shape == (2, 64, 3)
    bq = pu.ball_query(0.4, 16, xyz, new_xyz)
    assert np.array_equal(bq.cpu().numpy(), oracle.ball_query(0.4, 16, xyz.cpu().numpy(), new_xyz.cpu().numpy()))
    grouped = pu.QueryAndGroup(0.4, 16)(xyz, new_xyz, feats)
    assert grouped.shape == (2, 3 + 5, 64, 16)
    ref = np.concatenate([oracle.grouping_operation(xyz.cpu().numpy().transpose(0, 2, 1), bq.cpu().numpy())
                          - new_xyz.cpu().numpy().transpose(0, 2, 1)[..., None],
                          oracle.grouping_operation(feats.cpu().numpy(), bq.cpu().numpy())], axis=1)
    assert np.array_equal(grouped.cpu().numpy(), ref)
    assert pu.GroupAll()(xyz, None, feats).shape == (2, 8, 1, 900)
    with pytest.raises(RuntimeError):
        pu.furthest_point_sample(xyz.cpu(), 4)           # "CPU tensors not supported" (model.py:417)


def test_geometry_surface(oracle):
    """model.py:281-312 / loss.py:72-88: TorchCuboids / TorchCylinders built from the batch dict, .sdf and .sdf_sequence"""
    from mpinets_b200.geometry import TorchCuboids, TorchCylinders
    p = _problems(5)
    t = to_dev(p)
    cub = TorchCuboids(t["cuboid_centers"], t["cuboid_dims"], t["cuboid_quats"])
    cyl = TorchCylinders(t["cylinder_centers"], t["cylinder_radii"], t["cylinder_heights"], t["cylinder_quats"])
    rng = np.random.RandomState(1)
    pts = rng.uniform(-1, 1.5, size=(5, 300, 3)).astype(np.float32)
    sdf = torch.minimum(cub.sdf(torch.from_numpy(pts).cuda()), cyl.sdf(torch.from_numpy(pts).cuda()))   # loss.py:84
    assert np.array_equal(sdf.cpu().numpy(), oracle.sdf_points(p, pts))
    seq = torch.from_numpy(pts.reshape(5, 10, 30, 3)).cuda()
    sq = torch.minimum(cub.sdf_sequence(seq), cyl.sdf_sequence(seq))                                     # model.py:304-307
    assert sq.shape == (5, 10, 30) and np.array_equal(sq.cpu().numpy().reshape(5, 300), oracle.sdf_points(p, pts))
    empty = TorchCylinders(t["cylinder_centers"], torch.zeros_like(t["cylinder_radii"]), t["cylinder_heights"], t["cylinder_quats"])
    assert torch.isinf(empty.sdf(torch.from_numpy(pts).cuda())).all()                                    # geometry.py:465-468


def test_samplers_and_utils_surface(oracle, tables):
    """model.py:267-275,300-303; utils.py"""
    from mpinets_b200.robofin_shim import FrankaSampler, FrankaCollisionSampler
    from mpinets_b200 import utils
    rng = np.random.RandomState(2)
    lim = tables.joint_limits
    q = torch.from_numpy(rng.uniform(lim[:, 0], lim[:, 1], size=(9, 7)).astype(np.float32)).cuda()
    s = FrankaSampler("cuda:0", use_cache=True)
    pc = s.sample(q, 2048)
    assert pc.shape == (9, 2048, 3)
    pose = s.end_effector_pose(q)
    assert pose.shape == (9, 4, 4) and np.array_equal(pose[:, :3].cpu().numpy(), oracle.fk(q.cpu().numpy())[1])
    ee = s.sample_end_effector(pose, num_points=128)
    assert ee.shape == (9, 128, 3)
    # on the library (mpn_sample_end_effector): bit-exact against the oracle, and the same keyed subset as the cloud's target rows
    assert np.array_equal(ee.cpu().numpy(), oracle.sample_end_effector(pose[:, :3].cpu().numpy(), tables, 128, s.engine.cfg.seed, problem0=0))
    ee2 = s.sample_end_effector(pose, num_points=128)                  # a fresh subset per call, like robofin's np.random.choice
    assert not torch.equal(ee, ee2)
    assert np.array_equal(ee2.cpu().numpy(), oracle.sample_end_effector(pose[:, :3].cpu().numpy(), tables, 128, s.engine.cfg.seed, problem0=9))
    p9 = _problems(9, 2)
    sc9 = to_dev(p9)
    cloud9 = s.engine.build_cloud({k: sc9[k] for k in sc9 if k.startswith(("cuboid", "cylinder"))}, sc9["q0"], sc9["target"], problem0=0)
    assert torch.equal(s.engine.sample_end_effector(sc9["target"], 128, problem0=0), cloud9[:, 2048 + 4096:, :3])
    with pytest.raises(NotImplementedError):
        FrankaSampler("cuda:0", with_base_link=False)                  # the default link table carries panda_link0 points
    with pytest.raises(RuntimeError):
        FrankaCollisionSampler("cuda:0", with_base_link=True)          # the default sphere table has no panda_link0 sphere
    groups = FrankaCollisionSampler("cuda:0", with_base_link=False).compute_spheres(q)
    assert sum(c.shape[1] for _, c in groups) == tables.sphere_centers.shape[0]
    assert sorted(set(round(r, 4) for r, _ in groups)) == sorted(set(round(float(r), 4) for r in tables.sphere_radii))
    qn = utils.normalize_franka_joints(q)
    assert np.array_equal(qn.cpu().numpy(), oracle.normalize(q.cpu().numpy(), lim))
    assert np.array_equal(utils.unnormalize_franka_joints(qn).cpu().numpy(), oracle.unnormalize(qn.cpu().numpy(), lim))
    with pytest.raises(NotImplementedError):
        utils.normalize_franka_joints([0.0] * 7)             # utils.py:126-127


def test_model_surface(oracle, tables, state_dict):
    """MotionPolicyNetwork: reference state-dict keys load unchanged; forward / rollout keep the reference contracts."""
    from mpinets_b200.model import MotionPolicyNetwork
    from mpinets_b200.runtime import get_engine
    mdl = MotionPolicyNetwork(precision="fp32")
    assert set(mdl.state_dict().keys()) == set(state_dict.keys())
    mdl.load_state_dict(state_dict)
    mdl = mdl.cuda()
    p = _problems(4)
    eng = get_engine(torch.device("cuda", 0))
    sc = to_dev(p)
    q0 = torch.from_numpy(p["q0"]).cuda()
    tg = torch.from_numpy(p["target"]).cuda()
    xyz = eng.build_cloud({k: sc[k] for k in sc if k.startswith(("cuboid", "cylinder"))}, q0, tg)
    qn = eng.normalize(q0)
    dq = mdl(xyz, qn)                                                                              # model.py:75-91
    exp = oracle.policy_forward(state_dict, xyz.cpu().numpy(), qn.cpu().numpy())
    assert dq.shape == (4, 7) and (dq.cpu() - exp).abs().max().item() <= 1e-5
    batch = dict(xyz=xyz, configuration=qn, **{k: sc[k] for k in sc if k.startswith(("cuboid", "cylinder"))})
    before = xyz.clone()
    traj = mdl.rollout(batch, 2, sampler=None, unnormalize=True)                                   # model.py:128-183
    assert len(traj) == 3 and traj[0].shape == (4, 7)
    assert torch.equal(traj[0], eng.unnormalize(qn))
    assert not torch.equal(batch["xyz"][:, :2048], before[:, :2048])                               # in-place update (model.py:181)
    assert torch.equal(batch["xyz"][:, 2048:], before[:, 2048:])
    mdl.precision = "bf16"
    dq16 = mdl(before, qn)
    assert (dq16.cpu() - exp).abs().max().item() < 3e-4
    mdl.precision = "bf16x3"                                                                       # the default of inference modules
    assert MotionPolicyNetwork().precision == "bf16x3"
    assert (mdl(before, qn).cpu() - exp).abs().max().item() <= 1e-5
    # PointnetSAModule.forward honours its precision argument (fp32 / bf16 / bf16x3 kernels of the same module)
    from mpinets_b200 import _lib
    sa1 = mdl.point_cloud_encoder.SA_modules[0]
    xyz3, feat = before[..., :3].contiguous(), before[..., 3:].transpose(1, 2).contiguous()
    nx32, f32 = sa1(xyz3, feat)
    nx3, f3 = sa1(xyz3, feat, precision=_lib.PREC_BF16X3)
    nxb, fb = sa1(xyz3, feat, precision=_lib.PREC_BF16)
    assert torch.equal(nx3, nx32) and torch.equal(nxb, nx32)
    scale = f32.abs().max().item()
    assert (f3 - f32).abs().max().item() <= 2e-5 * scale and 1e-5 * scale < (fb - f32).abs().max().item() <= 2e-2 * scale


def test_two_models_share_one_engine_context_safely(oracle, tables, state_dict):
    """the per-device engine context holds ONE parameter set; a second module must never run on the first one's
    weights, and optimiser updates that live only in the context are pulled back before another module takes it over."""
    from mpinets_b200.model import MotionPolicyNetwork
    from mpinets_b200.runtime import get_engine
    a, b = MotionPolicyNetwork(precision="fp32"), MotionPolicyNetwork(precision="fp32")
    a.load_state_dict(state_dict)
    sd_b = {k: v.clone() for k, v in state_dict.items()}
    sd_b["decoder.6.bias"] = sd_b["decoder.6.bias"] + 0.25
    b.load_state_dict(sd_b)
    p = _problems(2)
    eng = get_engine(torch.device("cuda", 0))
    sc = to_dev(p)
    xyz = eng.build_cloud({k: sc[k] for k in sc if k.startswith(("cuboid", "cylinder"))}, sc["q0"], sc["target"])
    qn = eng.normalize(sc["q0"])
    da1 = a(xyz, qn)
    db = b(xyz, qn)
    da2 = a(xyz, qn)                                       # A again after B synced: must be A's weights, not B's
    assert torch.equal(da1, da2)
    assert ((db - da1) - 0.25).abs().max().item() < 1e-5


def test_evaluator_surface(tables, oracle):
    """metrics.py:Evaluator usage in run_inference.py:452-516: create_new_group -> evaluate_trajectory(...) per problem
    -> print_group_metrics; here one call evaluates the whole batch."""
    from mpinets_b200.metrics import Evaluator
    p = _problems(12)
    T1 = 11
    w = np.linspace(0.0, 1.0, T1, dtype=np.float32)[None, :, None]
    traj = torch.from_numpy((p["q0"][:, None] * (1 - w) + p["q_goal"][:, None] * w).astype(np.float32)).cuda().contiguous()
    ev = Evaluator()
    ev.create_new_group("tabletop")
    table = ev.evaluate_trajectories(traj, 0.1, torch.from_numpy(p["target"]).cuda(), to_dev(p))
    assert table.shape == (12, 16)
    g = ev.groups["tabletop"]
    assert len(g["success"]) == 12 and all(isinstance(v, bool) for v in g["success"])
    m = Evaluator.metrics(g)
    exp = oracle.evaluate(p, traj.cpu().numpy(), p["target"], tables)
    assert m["total"] == 12 and abs(m["success"] - 100 * np.count_nonzero(exp[:, 9]) / 12) < 1e-9
    assert abs(m["env collision"] - 100 * np.count_nonzero(exp[:, 0]) / 12) < 1e-9
    assert m["1 cm"] == 100.0                                  # joint-space interpolation ends exactly on the target
    # SPARC columns (metrics.py:387-409) against the oracle's restatement of sparc.py on float64 speed profiles
    from mpinets_b200.franka import fk_reference_f64
    th = traj.cpu().numpy().astype(np.float64)
    for b in (0, 5, 11):
        cfg = np.linalg.norm(np.diff(th[b], axis=0) / 0.1, axis=1)
        eff = np.linalg.norm(np.diff(np.stack([fk_reference_f64(q)[1][:3, 3] for q in th[b]]), axis=0) / 0.1, axis=1)
        assert abs(g["config_smoothness"][b] - oracle.sparc(cfg, 10.0)) < 2e-3
        assert abs(g["eff_smoothness"][b] - oracle.sparc(eff, 10.0)) < 2e-3
    ev.print_group_metrics()
    ev.print_overall_metrics()


def test_loss_surface(tables, oracle):
    """loss.py call patterns of model.py:222-236: container(y_hat, 7 scene tensors, supervision) -> two scalars that are
    weighted, summed and back-propagated into the network output."""
    from mpinets_b200 import loss as L
    from mpinets_b200.runtime import get_engine
    p = _problems(16)
    sc = to_dev(p)
    rng = np.random.default_rng(0)
    y_hat = torch.from_numpy(rng.uniform(-0.8, 0.8, (16, 7)).astype(np.float32)).cuda().requires_grad_(True)
    sup = (y_hat.detach() + 0.03).clamp(-1, 1)
    container = L.CollisionAndBCLossContainer()
    collision, point_match = container(y_hat, sc["cuboid_centers"], sc["cuboid_dims"], sc["cuboid_quats"], sc["cylinder_centers"],
                                       sc["cylinder_radii"], sc["cylinder_heights"], sc["cylinder_quats"], sup)
    assert collision.ndim == 0 and point_match.ndim == 0
    total = 5.0 * collision + 1.0 * point_match        # jobconfig.yaml:24-25
    total.backward()
    ol, og = oracle.bc_collision_losses(p, y_hat.detach().cpu().numpy(), sup.cpu().numpy(), tables, get_engine().cfg.seed, 1024, 0.03, 5.0, 1.0)
    assert abs(collision.item() - ol[0]) < 1e-6 and abs(point_match.item() - ol[1]) < 1e-6
    assert np.abs(y_hat.grad.cpu().numpy() - og).max() < 2e-5 * np.abs(og).max() + 1e-8
    # the standalone functions, differentiable w.r.t. the cloud (loss.py:31-94)
    pc = torch.from_numpy(rng.uniform(-0.5, 1.0, (16, 200, 3)).astype(np.float32)).cuda().requires_grad_(True)
    cl = L.collision_loss(pc, sc["cuboid_centers"], sc["cuboid_dims"], sc["cuboid_quats"], sc["cylinder_centers"], sc["cylinder_radii"],
                          sc["cylinder_heights"], sc["cylinder_quats"])
    pm = L.point_match_loss(pc, pc.detach() + 0.01)
    (cl + pm).backward()
    oc, ogc = oracle.collision_loss(p, pc.detach().cpu().numpy())
    opm, ogp = oracle.point_match_loss(pc.detach().cpu().numpy(), (pc.detach() + 0.01).cpu().numpy())
    assert abs(cl.item() - oc) < 1e-6 and abs(pm.item() - opm) < 1e-6
    assert np.abs(pc.grad.cpu().numpy() - (ogc + ogp)).max() < 1e-5 * np.abs(ogc + ogp).max()


def test_run_inference_surface(state_dict):
    """run_inference.calculate_metrics (run_inference.py:426-516) over a ProblemSet of PlanningProblem records: clouds,
    rollout_until_success semantics and evaluation for every group, on the device"""
    from mpinets_b200 import mpinets_types as T
    from mpinets_b200.model import MotionPolicyNetwork
    from mpinets_b200.run_inference import calculate_metrics, run_problems
    probs = T.soa_to_problems(_problems(6))
    mdl = MotionPolicyNetwork(precision="bf16")
    mdl.load_state_dict(state_dict)
    out = run_problems(mdl, probs, max_steps=4)
    assert out["trajectories"].shape == (6, 5, 7) and out["eval"].shape == (6, 16)
    assert (out["num_poses"] >= 2).all() and (out["num_poses"] <= 5).all()
    assert torch.equal(out["eval"][:, 10].to(torch.int32), out["num_poses"])
    ev = calculate_metrics(mdl, {"tabletop": {"task_oriented": probs[:4]}, "cubby": {"neutral_start": probs[4:]}}, max_steps=3)
    assert list(ev.groups) == ["tabletop, task_oriented", "cubby, neutral_start"]
    assert len(ev.groups["tabletop, task_oriented"]["success"]) == 4 and len(ev.groups["cubby, neutral_start"]["success"]) == 2
    ev.print_overall_metrics()
    # problems that carry a sensed obstacle cloud go through make_point_cloud_from_problem (run_inference.py:58-90)
    rng = np.random.default_rng(0)
    for q in probs:
        q.obstacle_point_cloud = rng.uniform(-1, 1, (int(rng.integers(4096, 6000)), 3)).astype(np.float32)
    out2 = run_problems(mdl, probs, max_steps=2)
    assert out2["trajectories"].shape == (6, 3, 7) and torch.isfinite(out2["eval"]).all()
