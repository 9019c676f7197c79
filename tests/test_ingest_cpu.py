"""CPU tests of the ingestion layer: the ProblemSet pickle reader, the HDF5-layout row store,
and the oracle's statements of the joint-noise augmentation and the sensed-cloud crop."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem_set(n=5):
    from mpinets_b200 import scenes
    from mpinets_b200.mpinets_types import soa_to_problems
    ps = {}
    for env, cfg in (("tabletop", 2), ("dresser", 3)):
        probs = soa_to_problems(scenes.config_problems(cfg, 2 * n))
        probs[0].obstacle_point_cloud = np.arange(30, dtype=np.float64).reshape(10, 3)
        ps[env] = {"task_oriented": probs[:n], "neutral_start": probs[n:]}
    return ps


@pytest.mark.parametrize("private_names", [False, True])
def test_problem_set_pickle_round_trip(tmp_path, private_names):
    """a pickle that names the reference's classes (mpinets.mpinets_types.PlanningProblem, geometrout.primitive.*, geometrout.transform.*,
    pyquaternion Quaternion) is read back WITHOUT those packages being importable, in both attribute layouts the reader accepts"""
    from mpinets_b200 import problem_io as io_
    from mpinets_b200.mpinets_types import flatten_problem_set, problems_to_soa
    ps = _problem_set()
    path = str(tmp_path / "problems.pkl")
    io_.dump_problem_set(ps, path, private_names=private_names)
    raw = open(path, "rb").read()
    for token in (b"mpinets.mpinets_types", b"PlanningProblem", b"geometrout.primitive", b"Cuboid", b"geometrout.transform", b"SE3"):
        assert token in raw
    if private_names:
        assert b"pyquaternion" in raw
    for mod in ("geometrout", "pyquaternion", "mpinets"):
        assert mod not in sys.modules
    back = io_.load_problem_set(path)
    assert list(back) == list(ps) and all(list(back[e]) == list(ps[e]) for e in ps)
    a = problems_to_soa([p for _, _, p in flatten_problem_set(ps)])
    b = problems_to_soa([p for _, _, p in flatten_problem_set(back)])
    for k in a:
        if isinstance(a[k], np.ndarray):
            assert np.array_equal(a[k], b[k]), k
    for k in ("target_volume", "negative_volumes"):
        for kk in a[k]:
            assert np.array_equal(a[k][kk], b[k][kk])
    assert np.array_equal(back["tabletop"]["task_oriented"][0].obstacle_point_cloud, ps["tabletop"]["task_oriented"][0].obstacle_point_cloud)
    # run_inference.py:462-467: environment / problem-type filters, "-" spelled as "_"
    only = io_.load_problem_set(path, "dresser", "neutral-start")
    assert list(only) == ["dresser"] and list(only["dresser"]) == ["neutral_start"]
    # plain pickle.load cannot read it here (the reference's packages are absent): the reader is what makes the file usable
    with pytest.raises(ModuleNotFoundError):
        pickle.loads(raw)


def test_problem_set_reader_rejects_foreign_globals():
    from mpinets_b200 import problem_io as io_
    evil = pickle.dumps({"tabletop": {"x": [subprocess.check_output]}})
    with pytest.raises(pickle.UnpicklingError):
        io_.loads_problem_set(evil)


def test_problem_set_reader_reports_unknown_layout():
    from mpinets_b200 import problem_io as io_
    mods, C = io_._ref_module_stubs()
    sys.modules.update(mods)
    try:
        o = C[("geometrout.primitive", "Cuboid")]()
        o.__dict__.update(extents=np.ones(3))
        so3 = C[("geometrout.transform", "SO3")](); so3.__dict__.update(wxyz=np.array([1.0, 0, 0, 0]))
        tgt = C[("geometrout.transform", "SE3")](); tgt.__dict__.update(_xyz=np.zeros(3), so3=so3)
        p = C[("mpinets.mpinets_types", "PlanningProblem")]()
        p.__dict__.update(target=tgt, target_volume=o, q0=np.zeros(7), obstacles=[o], obstacle_point_cloud=None, target_negative_volumes=[])
        data = pickle.dumps({"a": {"b": [p]}})
    finally:
        for m in mods:
            sys.modules.pop(m, None)
    with pytest.raises(KeyError) as e:
        io_.loads_problem_set(data)
    assert "extents" in str(e.value)          # the message lists the attribute names that WERE pickled


def test_trajectory_store_rows():
    """HDF5 layout of gen_data.py:676-700 / data_loader.py:187-235: all-zero rows are padding, their quaternions become identity"""
    from mpinets_b200.problem_io import TrajectoryStore
    rng = np.random.RandomState(0)
    n, T = 6, 50
    store = {"global_solutions": rng.uniform(-1, 1, size=(n, T, 7)).astype(np.float32),
             "cuboid_centers": rng.normal(size=(n, 12, 3)).astype(np.float32), "cuboid_dims": rng.uniform(0.1, 1, size=(n, 12, 3)).astype(np.float32),
             "cuboid_quaternions": np.tile(np.array([1, 0, 0, 0], np.float32), (n, 12, 1))}
    store["cuboid_dims"][:, 9:] = 0; store["cuboid_centers"][:, 9:] = 0; store["cuboid_quaternions"][:, 9:] = 0
    ts = TrajectoryStore(store, "global_solutions")
    assert len(ts) == n * T and ts.expert_length == T
    rows = ts.scene_rows([4, 1, 4])
    assert rows["cuboid_dims"].shape == (3, 40, 3) and rows["cylinder_radii"].shape == (3, 40, 1)
    assert np.array_equal(rows["cuboid_dims"][0, :12], store["cuboid_dims"][4]) and np.array_equal(rows["cuboid_dims"][1, :12], store["cuboid_dims"][1])
    assert np.array_equal(rows["cuboid_dims"][2], rows["cuboid_dims"][0])
    assert (rows["cuboid_quats"][:, 9:, 0] == 1).all() and (rows["cylinder_quats"][..., 0] == 1).all()     # data_loader.py:198-202
    assert (rows["cylinder_radii"] == 0).all()                                                              # no cylinder datasets: :208-214
    assert np.array_equal(ts.configurations([4, 1], [0, 49]), np.stack([store["global_solutions"][4, 0], store["global_solutions"][1, 49]]))
    with pytest.raises(ValueError):
        TrajectoryStore(store, "global_solutions", max_cuboids=8).scene_rows([0])


def test_oracle_joint_noise(oracle, tables):
    """data_loader.py:167-180: noise ~ N(0, scale^2) per joint, clamped to the limits, keyed per (sample, epoch)"""
    lim = tables.joint_limits
    mid = ((lim[:, 0] + lim[:, 1]) / 2).astype(np.float32)
    q = np.tile(mid, (20000, 1))
    out, outn = oracle.augment_joints(q, tables, 0.015, 7)
    z = (out - q) / 0.015
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02 and abs(np.mean(z ** 3)) < 0.05
    assert np.array_equal(outn, oracle.normalize(out, lim))
    again, _ = oracle.augment_joints(q[:8], tables, 0.015, 7, ids=np.arange(8))
    assert np.array_equal(again, out[:8])                                     # default ids = row index
    other, _ = oracle.augment_joints(q[:8], tables, 0.015, 7, ids=np.arange(8), epoch=1)
    assert not np.array_equal(other, out[:8])
    edge = np.tile(lim[:, 1].astype(np.float32), (1000, 1))
    oe, oen = oracle.augment_joints(edge, tables, 0.05, 7)
    assert (oe <= lim[:, 1]).all() and (oe >= lim[:, 0]).all() and (oe == lim[:, 1]).mean() > 0.3 and oen.max() <= 1.0


def test_oracle_clean_point_cloud(oracle):
    """planning_node.py:187-228: the kept set equals numpy's mask, the output is a subset without replacement of it"""
    rng = np.random.RandomState(3)
    xyz = rng.uniform([-0.6, -0.8, -0.2], [1.6, 1.9, 0.6], size=(60000, 3)).astype(np.float32)
    xyz[:7] = [[0.25, 1.0, 0.2], [1.35, 0, 0], [0.5, -0.3, 0.2], [0.5, 0, 0.35], [0.5, 0, -0.05], [-0.35, 0, 0], [0.1, 0.5, 0]]   # on the faces: strict
    rgba = rng.uniform(size=(60000, 4)).astype(np.float32)
    task = np.logical_and.reduce((xyz[:, 0] > 0.25, xyz[:, 0] < 1.35, xyz[:, 1] > -0.3, xyz[:, 1] < 1.6, xyz[:, 2] > -0.05, xyz[:, 2] < 0.35))
    mount = np.logical_and.reduce((xyz[:, 0] > -0.35, xyz[:, 0] < 0.30, xyz[:, 1] > -0.5, xyz[:, 1] < 0.5, xyz[:, 2] > -0.05, xyz[:, 2] < 0.05))
    mask = np.logical_or(task, mount)
    assert not mask[:7].any()
    kept, out, outc = oracle.clean_point_cloud(xyz, rgba, 4096, 11)
    assert kept == mask.sum() and kept > 4096
    rows = {tuple(r) for r in xyz[mask].tolist()}
    assert all(tuple(r) in rows for r in out.tolist()) and len({tuple(r) for r in out.tolist()}) == 4096
    src = {tuple(x): tuple(c) for x, c in zip(xyz.tolist(), rgba.tolist())}
    assert all(src[tuple(x)] == tuple(c) for x, c in zip(out.tolist(), outc.tolist()))
    k2, o2, _ = oracle.clean_point_cloud(xyz, None, 4096, 11, cloud_id=1)
    assert k2 == kept and not np.array_equal(o2, out)
    assert oracle.clean_point_cloud(xyz[:3000], None, 4096, 11)[1] is None         # fewer than 4096 inside: nothing written
