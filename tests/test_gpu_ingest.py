"""GPU tests (-m gpu) of the ingestion kernels (ingest.cu, mpn_build_cloud_ids) and the batched get_inputs replacement."""
import numpy as np
import pytest
import torch

from conftest import to_dev

pytestmark = pytest.mark.gpu


def test_augment_joints_matches_oracle(engine, oracle, tables):
    """data_loader.py:167-180 on the device: same Philox / Box-Muller streams as the oracle (libm vs CUDA logf / sincosf: tolerance)"""
    rng = np.random.RandomState(0)
    lim = tables.joint_limits
    q = rng.uniform(lim[:, 0], lim[:, 1], size=(3000, 7)).astype(np.float32)
    q[:50] = lim[:, 1]                                  # at the upper limits: about half of the noise is clamped away
    ids = rng.randint(0, 2 ** 31 - 1, size=3000).astype(np.int64)
    out, outn = engine.augment_joints(torch.from_numpy(q).cuda(), 0.015, sample0=torch.from_numpy(ids).cuda(), epoch=3)
    eo, eon = oracle.augment_joints(q, tables, 0.015, engine.cfg.seed, ids=ids, epoch=3)
    assert np.abs(out.cpu().numpy() - eo).max() <= 5e-7 and np.abs(outn.cpu().numpy() - eon).max() <= 5e-7   # one float32 ulp at |q| <= 4
    assert (out.cpu().numpy() <= lim[:, 1]).all() and (out.cpu().numpy() >= lim[:, 0]).all()
    assert np.array_equal(outn.cpu().numpy(), oracle.normalize(out.cpu().numpy(), lim))      # normalisation itself is bit-exact
    d, _ = engine.augment_joints(torch.from_numpy(q[:8]).cuda(), 0.015)                        # default ids = row index
    assert np.abs(d.cpu().numpy() - oracle.augment_joints(q[:8], tables, 0.015, engine.cfg.seed)[0]).max() <= 5e-7


def test_build_cloud_with_sample_ids(engine, oracle, tables):
    """mpn_build_cloud_ids: row b's streams are keyed by ids[b] -- equal to building problem ids[b] on its own, whatever the batch"""
    from mpinets_b200 import scenes
    p = scenes.config_problems(4, 6)
    sc = to_dev(p)
    scene = {k: sc[k] for k in scenes.SCENE_KEYS}
    ids = np.array([17, 3, 2 ** 31 + 5, 3, 900000, 0], np.int64)
    cloud = engine.build_cloud(scene, sc["q0"], sc["target"], problem_ids=torch.from_numpy(ids).cuda()).cpu().numpy()
    for b, i in enumerate(ids):
        exp = oracle.build_cloud(p["q0"][b:b + 1], p["target"][b:b + 1], {k: v[b:b + 1] for k, v in p.items()}, tables, engine.cfg.seed,
                                 problem0=int(i))
        assert np.array_equal(cloud[b], exp[0]), b
    again = engine.build_cloud(scene, sc["q0"], sc["target"], problem_ids=torch.from_numpy(ids).cuda(), epoch=1).cpu().numpy()
    assert not np.array_equal(again[:, 2048:6144], cloud[:, 2048:6144])       # another epoch: other surface samples
    assert np.array_equal(again[:, :2048], cloud[:, :2048])                   # the robot subset is keyed by the step, not the problem


def test_clean_point_cloud_bit_exact(engine, oracle):
    rng = np.random.RandomState(3)
    N = 307200                                            # a 640 x 480 depth image worth of points
    xyz = rng.uniform([-0.6, -0.8, -0.2], [1.6, 1.9, 0.6], size=(N, 3)).astype(np.float32)
    xyz[:7] = [[0.25, 1.0, 0.2], [1.35, 0, 0], [0.5, -0.3, 0.2], [0.5, 0, 0.35], [0.5, 0, -0.05], [-0.35, 0, 0], [0.1, 0.5, 0]]
    rgba = rng.uniform(size=(N, 4)).astype(np.float32)
    for cid in (0, 5):
        out, outc = engine.clean_point_cloud(torch.from_numpy(xyz).cuda(), torch.from_numpy(rgba).cuda(), 4096, cloud_id=cid)
        kept, eo, ec = oracle.clean_point_cloud(xyz, rgba, 4096, engine.cfg.seed, cloud_id=cid)
        assert np.array_equal(out.cpu().numpy(), eo) and np.array_equal(outc.cpu().numpy(), ec)
    out2, none = engine.clean_point_cloud(torch.from_numpy(xyz[:1025]).cuda(), None, 16)       # ragged N (1024 + 1), no colours
    assert none is None and np.array_equal(out2.cpu().numpy(), oracle.clean_point_cloud(xyz[:1025], None, 16, engine.cfg.seed)[1])
    with pytest.raises(ValueError):
        engine.clean_point_cloud(torch.from_numpy(xyz[:3000]).cuda(), None, 4096)              # np.random.choice would raise too


def test_batch_inputs_replaces_get_inputs(engine, oracle, tables):
    """problem_io.batch_inputs = PointCloudInstanceDataset.__getitem__ for a batch (data_loader.py:141-280,403-417) with the cloud
    built on the device: keys / shapes of the reference's batch dict, every tensor against the oracle"""
    from mpinets_b200 import scenes
    from mpinets_b200.problem_io import TrajectoryStore, batch_inputs
    n, T = 8, 50
    p = scenes.config_problems(4, n)
    w = np.linspace(0, 1, T, dtype=np.float32)[None, :, None]
    store = {"hybrid_solutions": (p["q0"][:, None] * (1 - w) + p["q_goal"][:, None] * w).astype(np.float32),
             "cuboid_centers": p["cuboid_centers"], "cuboid_dims": p["cuboid_dims"], "cuboid_quaternions": p["cuboid_quats"].copy(),
             "cylinder_centers": p["cylinder_centers"], "cylinder_radii": p["cylinder_radii"], "cylinder_heights": p["cylinder_heights"],
             "cylinder_quaternions": p["cylinder_quats"].copy()}
    store["cuboid_quaternions"][np.isclose(p["cuboid_dims"], 0).any(-1)] = 0         # absent primitives are stored as all-zero rows
    ts = TrajectoryStore(store, "hybrid_solutions")
    idx = np.array([3 * T + 7, 0 * T + 49, 5 * T + 0, 3 * T + 7], np.int64)
    item = batch_inputs(engine, ts, idx, train=True, random_scale=0.015, epoch=2)
    assert set(item) >= {"xyz", "configuration", "supervision", "target_position", "cuboid_centers", "cuboid_dims", "cuboid_quats",
                         "cylinder_centers", "cylinder_radii", "cylinder_heights", "cylinder_quats"}
    assert item["xyz"].shape == (4, 6272, 4) and item["configuration"].shape == (4, 7) and item["supervision"].shape == (4, 7)
    tr, ti = np.divmod(idx, T)
    q = store["hybrid_solutions"][tr, ti]
    qa, qan = oracle.augment_joints(q, tables, 0.015, engine.cfg.seed, ids=idx, epoch=2)
    assert np.abs(item["configuration"].cpu().numpy() - qan).max() <= 5e-7
    sup_t = np.clip(ti + 1, 0, T - 1)                                                # data_loader.py:405-409
    assert np.array_equal(item["supervision"].cpu().numpy(), oracle.normalize(store["hybrid_solutions"][tr, sup_t], tables.joint_limits))
    _, eef = oracle.fk(store["hybrid_solutions"][tr, -1])
    assert np.array_equal(item["target_position"].cpu().numpy(), eef[:, :, 3])
    assert torch.equal(item["xyz"][0], item["xyz"][3])                               # same dataset index -> same item, wherever it sits
    xyz = item["xyz"].cpu().numpy()
    q_used = engine.unnormalize(item["configuration"]).cpu().numpy()
    for b in range(4):   # obstacle + target rows exactly; robot rows at the device's (tolerance-level) augmented configuration
        one = {k: v[tr[b]:tr[b] + 1] for k, v in p.items() if isinstance(v, np.ndarray)}
        exp = oracle.build_cloud(qa[b:b + 1], eef[b:b + 1], one, tables, engine.cfg.seed, problem0=int((idx[b] + (2 << 20)) & 0xFFFFFFFF))
        assert np.array_equal(xyz[b, 2048:], exp[0, 2048:])
        assert np.abs(xyz[b, :2048] - exp[0, :2048]).max() < 1e-5
    val = batch_inputs(engine, ts, np.array([2, 6]), train=False, trajectory_dataset=True)    # PointCloudTrajectoryDataset (:331-341)
    assert "supervision" not in val
    assert np.array_equal(val["configuration"].cpu().numpy(), oracle.normalize(store["hybrid_solutions"][[2, 6], 0], tables.joint_limits))
