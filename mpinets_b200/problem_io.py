"""Problem-set and dataset ingestion.

* ``load_problem_set`` reads the reference's evaluation pickles -- ``ProblemSet = Dict[env][problem type] -> List[PlanningProblem]``
  (``/root/reference/mpinets/mpinets_types.py:34-48``, written by ``data_pipeline/gen_data.py:818-972`` and read at
  ``run_inference.py:460-468``) -- WITHOUT ``geometrout`` / ``pyquaternion`` / ``mpinets`` being importable: a restricted
  unpickler maps the pickled class paths (``mpinets.mpinets_types.PlanningProblem``, ``geometrout.primitive.{Cuboid,Cylinder,Sphere}``,
  ``geometrout.transform.{SE3,SO3}``, ``pyquaternion.quaternion.Quaternion``) to attribute bags, which are then converted to the
  light-weight types of ``mpinets_types``.  geometrout 0.0.3.4's private attribute names are not in the reference tree
  ([UNVERIFIED]), so the converter accepts the public names the reference reads (``dims``, ``radius``, ``height``, ``pose``,
  ``xyz``, ``so3``, ``wxyz``) and their underscore-prefixed forms, and fails loudly (listing the keys it found) otherwise.
  Only numpy reconstruction helpers and those class paths are allowed: anything else in the stream raises ``UnpicklingError``.
* ``dump_problem_set`` writes a pickle with the reference's class paths from this package's records (round-trip tests; exporting
  generated problems to the reference's ``run_inference.py``).
* ``TrajectoryStore`` / ``batch_inputs`` are the batched, GPU-side counterpart of ``PointCloudBase.get_inputs``
  (``data_loader.py:141-280``): rows of the HDF5 layout (``cuboid_centers / cuboid_dims / cuboid_quaternions / cylinder_* /
  <trajectory_key>``; any mapping of arrays, e.g. an open ``h5py.File`` or a dict of numpy arrays) -> the batch dict of
  ``data_loader.py:153-280,410-415`` with the cloud built on the device, including the training-time joint noise (``:167-180``).
"""
from __future__ import annotations

import io
import pickle
import sys
import types
from typing import Any, Dict, List, Mapping, Optional, Sequence

import numpy as np

from .mpinets_types import SE3, Cuboid, Cylinder, PlanningProblem, ProblemSet

_REF_CLASSES = {
    ("mpinets.mpinets_types", "PlanningProblem"), ("geometrout.primitive", "Cuboid"), ("geometrout.primitive", "Cylinder"),
    ("geometrout.primitive", "Sphere"), ("geometrout.transform", "SE3"), ("geometrout.transform", "SO3"),
    ("pyquaternion.quaternion", "Quaternion"), ("pyquaternion", "Quaternion"),
}
_SAFE_GLOBALS = {
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy", "ndarray"), ("numpy", "dtype"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"), ("numpy.core.numeric", "_frombuffer"),
    ("numpy._core.numeric", "_frombuffer"), ("builtins", "list"), ("builtins", "dict"), ("builtins", "tuple"), ("builtins", "set"),
    ("builtins", "float"), ("builtins", "int"), ("builtins", "complex"), ("collections", "OrderedDict"), ("copyreg", "_reconstructor"),
    ("builtins", "object"),
}


class _Bag:
    """stands in for a pickled reference object: keeps its attribute dict, whatever it is"""
    _ref_path = ("", "")

    def __init__(self, *a, **k):
        self.__dict__["_args"] = (a, k)

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):   # (dict, slots) form
            state = {**(state[0] or {}), **state[1]}
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state


def _bag_class(module: str, name: str):
    return type(name, (_Bag,), {"_ref_path": (module, name), "__module__": __name__})


_BAGS = {path: _bag_class(*path) for path in _REF_CLASSES}


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str):
        if (module, name) in _BAGS:
            return _BAGS[(module, name)]
        if (module, name) in _SAFE_GLOBALS:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"problem-set pickle references {module}.{name}: not part of the PlanningProblem format")


def _get(bag: Any, *names: str):
    d = bag.__dict__ if isinstance(bag, _Bag) else None
    for n in names:
        if d is not None and n in d:
            return d[n]
        if d is None and hasattr(bag, n):
            return getattr(bag, n)
    keys = sorted(k for k in (d or {}) if k != "_args")
    raise KeyError(f"{type(bag).__name__}: none of {names} among the pickled attributes {keys}")


def _has(bag: Any, *names: str) -> bool:
    try:
        _get(bag, *names)
        return True
    except KeyError:
        return False


def _wxyz(q: Any) -> np.ndarray:
    """SO3 / pyquaternion Quaternion / array -> wxyz"""
    if isinstance(q, _Bag):
        path = q._ref_path[1]
        if path == "Quaternion":
            return np.asarray(_get(q, "q", "_q"), dtype=np.float64).reshape(4)
        if _has(q, "wxyz", "_wxyz"):
            return np.asarray(_get(q, "wxyz", "_wxyz"), dtype=np.float64).reshape(4)
        return _wxyz(_get(q, "_quat", "quat", "_quaternion", "quaternion", "q"))
    if hasattr(q, "wxyz"):
        return np.asarray(q.wxyz, dtype=np.float64).reshape(4)
    return np.asarray(q, dtype=np.float64).reshape(4)


def _pose(p: Any) -> SE3:
    if isinstance(p, SE3):
        return p
    if _has(p, "_xyz", "xyz") and _has(p, "so3", "_so3"):
        return SE3(np.asarray(_get(p, "_xyz", "xyz"), dtype=np.float64).reshape(3), _wxyz(_get(p, "so3", "_so3")))
    if _has(p, "_matrix", "matrix"):
        return SE3.from_matrix(np.asarray(_get(p, "_matrix", "matrix"), dtype=np.float64))
    if _has(p, "_xyz", "xyz"):
        return SE3(np.asarray(_get(p, "_xyz", "xyz"), dtype=np.float64).reshape(3), _wxyz(_get(p, "_quat", "quat", "_quaternion", "quaternion", "wxyz")))
    raise KeyError(f"SE3: cannot find a pose among {sorted(p.__dict__)}")


def _primitive(o: Any):
    if o is None or isinstance(o, (Cuboid, Cylinder)):
        return o
    kind = o._ref_path[1] if isinstance(o, _Bag) else type(o).__name__
    if _has(o, "pose", "_pose"):
        pose = _pose(_get(o, "pose", "_pose"))
        center, quat = pose.xyz, pose.so3.wxyz
    else:
        center = np.asarray(_get(o, "center", "_center"), dtype=np.float64).reshape(3)
        quat = _wxyz(_get(o, "quaternion", "_quaternion", "_quat", "wxyz")) if _has(o, "quaternion", "_quaternion", "_quat", "wxyz") else (1, 0, 0, 0)
    if kind == "Cuboid":
        return Cuboid(center, np.asarray(_get(o, "dims", "_dims"), dtype=np.float64).reshape(3), quat)
    if kind == "Cylinder":
        return Cylinder(center, float(_get(o, "radius", "_radius")), float(_get(o, "height", "_height")), quat)
    raise TypeError(f"unsupported primitive in a PlanningProblem: {kind} (the engine takes cuboids and cylinders, gen_data.py:87-88)")


def _problem(p: Any) -> PlanningProblem:
    if isinstance(p, PlanningProblem):
        return p
    cloud = _get(p, "obstacle_point_cloud") if _has(p, "obstacle_point_cloud") else None
    return PlanningProblem(
        target=_pose(_get(p, "target")), target_volume=_primitive(_get(p, "target_volume")),
        q0=np.asarray(_get(p, "q0"), dtype=np.float64).reshape(7),
        obstacles=[_primitive(o) for o in (_get(p, "obstacles") or [])] if _has(p, "obstacles") else None,
        obstacle_point_cloud=None if cloud is None else np.asarray(cloud),
        target_negative_volumes=[_primitive(o) for o in (_get(p, "target_negative_volumes") if _has(p, "target_negative_volumes") else [])])


def loads_problem_set(data: bytes) -> ProblemSet:
    raw = _RefUnpickler(io.BytesIO(data)).load()
    if not isinstance(raw, dict):
        raise ValueError("a problem-set pickle holds Dict[env type][problem type] -> List[PlanningProblem] (mpinets_types.py:48)")
    out: ProblemSet = {}
    for env, kinds in raw.items():
        if not isinstance(kinds, dict):
            raise ValueError(f"problem set: entry {env!r} is not a dict of problem types")
        out[env] = {kind: [_problem(p) for p in plist] for kind, plist in kinds.items()}
    return out


def load_problem_set(path: str, environment_type: str = "all", problem_type: str = "all") -> ProblemSet:
    """run_inference.py:460-468: load, then optionally keep one environment class / one problem type ("-" -> "_")."""
    with open(path, "rb") as f:
        problems = loads_problem_set(f.read())
    env, kind = environment_type.replace("-", "_"), problem_type.replace("-", "_")
    if env != "all":
        problems = {env: problems[env]}
    if kind != "all":
        problems = {k: {kind: v[kind]} for k, v in problems.items()}
    return problems


# ---------------------------------------------------------------------------------------------- writer (reference class paths)
def _ref_module_stubs():
    """modules carrying plain classes at the reference's pickled paths, so pickle.dumps records exactly those paths"""
    mods: Dict[str, types.ModuleType] = {}
    classes = {}
    for module, name in sorted(_REF_CLASSES - {("pyquaternion", "Quaternion")}):
        m = mods.setdefault(module, types.ModuleType(module))
        cls = type(name, (), {"__module__": module})
        setattr(m, name, cls)
        classes[(module, name)] = cls
    for module in list(mods):   # parent packages
        parent = module.split(".")[0]
        if parent not in mods:
            mods[parent] = types.ModuleType(parent)
        if "." in module:
            setattr(mods[parent], module.split(".")[1], mods[module])
    return mods, classes


def dumps_problem_set(problem_set: ProblemSet, private_names: bool = False) -> bytes:
    """Pickle with the class paths of the reference's own pickles.  ``private_names``: store the attributes under
    underscore-prefixed names and the rotation as a pyquaternion ``Quaternion`` (the other layout the reader accepts)."""
    mods, C = _ref_module_stubs()

    def obj(path, **attrs):
        o = C[path]()
        o.__dict__.update(attrs)
        return o

    us = (lambda n: "_" + n) if private_names else (lambda n: n)

    def so3(wxyz):
        w = np.asarray(wxyz, dtype=np.float64)
        if private_names:
            return obj(("geometrout.transform", "SO3"), _quat=obj(("pyquaternion.quaternion", "Quaternion"), q=w))
        return obj(("geometrout.transform", "SO3"), wxyz=w)

    def se3(p: SE3):
        return obj(("geometrout.transform", "SE3"), **{"_xyz": np.asarray(p.xyz, dtype=np.float64), us("so3"): so3(p.so3.wxyz)})

    def prim(o):
        if o is None:
            return None
        if hasattr(o, "radius"):
            return obj(("geometrout.primitive", "Cylinder"), **{us("pose"): se3(o.pose), us("radius"): o.radius, us("height"): o.height})
        return obj(("geometrout.primitive", "Cuboid"), **{us("pose"): se3(o.pose), us("dims"): np.asarray(o.dims, dtype=np.float64)})

    def prob(p: PlanningProblem):
        return obj(("mpinets.mpinets_types", "PlanningProblem"), target=se3(p.target), target_volume=prim(p.target_volume),
                   q0=np.asarray(p.q0, dtype=np.float64), obstacles=None if p.obstacles is None else [prim(o) for o in p.obstacles],
                   obstacle_point_cloud=p.obstacle_point_cloud, target_negative_volumes=[prim(o) for o in p.target_negative_volumes])

    tree = {env: {kind: [prob(p) for p in plist] for kind, plist in kinds.items()} for env, kinds in problem_set.items()}
    saved = {name: sys.modules.get(name) for name in mods}
    sys.modules.update(mods)
    try:
        return pickle.dumps(tree, protocol=4)
    finally:
        for name, old in saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old


def dump_problem_set(problem_set: ProblemSet, path: str, private_names: bool = False) -> None:
    with open(path, "wb") as f:
        f.write(dumps_problem_set(problem_set, private_names))


# ---------------------------------------------------------------------------------------------- HDF5-layout rows -> device batch
HDF5_KEYS = {"cuboid_centers": "cuboid_centers", "cuboid_dims": "cuboid_dims", "cuboid_quats": "cuboid_quaternions",
             "cylinder_centers": "cylinder_centers", "cylinder_radii": "cylinder_radii", "cylinder_heights": "cylinder_heights",
             "cylinder_quats": "cylinder_quaternions"}   # batch key -> dataset name (data_loader.py:187-235; gen_data.py:676-700)


class TrajectoryStore:
    """The rows ``PointCloudBase`` reads from its HDF5 file (data_loader.py:82-95,153-235), from any mapping of arrays:
    ``store[trajectory_key]`` [n, T, 7] expert trajectories and the primitive datasets of ``HDF5_KEYS`` ([n, M, .]; the cylinder
    datasets may be absent, data_loader.py:208-214).  Pass an open ``h5py.File`` (when h5py is installed) or a dict / npz of arrays."""

    def __init__(self, store: Mapping[str, Any], trajectory_key: str = "global_solutions", max_cuboids: int = 40, max_cylinders: int = 40):
        self.store, self.trajectory_key = store, trajectory_key
        self.max_cuboids, self.max_cylinders = max_cuboids, max_cylinders
        shape = store[trajectory_key].shape
        self.num_trajectories, self.expert_length = int(shape[0]), int(shape[1])

    def __len__(self):   # PointCloudInstanceDataset.__len__ (data_loader.py:395-401)
        return self.num_trajectories * self.expert_length

    def _rows(self, name: str, idx: np.ndarray, width: int, rows: int) -> np.ndarray:
        out = np.zeros((len(idx), rows, width), np.float32)
        if name in self.store:
            order = np.argsort(idx, kind="stable")               # h5py fancy indexing wants increasing indices
            uniq, inv = np.unique(idx[order], return_inverse=True)
            a = np.asarray(self.store[name][uniq.tolist() if hasattr(self.store[name], "id") else uniq], dtype=np.float32)
            a = a.reshape(len(uniq), -1, width)[inv]
            if a.shape[1] > rows:
                raise ValueError(f"{name}: {a.shape[1]} primitive rows per problem, engine was built for {rows}")
            out[order, : a.shape[1]] = a
        return out

    def scene_rows(self, trajectory_idx: Sequence[int]) -> Dict[str, np.ndarray]:
        """the seven primitive arrays for these trajectories, padded to the engine's row counts; all-zero quaternions of absent
        primitives are patched to identity (data_loader.py:198-202,229-230)"""
        idx = np.asarray(trajectory_idx, dtype=np.int64)
        out = {}
        for key, name in HDF5_KEYS.items():
            width = {"centers": 3, "dims": 3, "quats": 4, "radii": 1, "heights": 1}[key.split("_")[1]]
            out[key] = self._rows(name, idx, width, self.max_cuboids if key.startswith("cuboid") else self.max_cylinders)
        for key in ("cuboid_quats", "cylinder_quats"):
            q = out[key]
            q[np.all(np.isclose(q, 0), axis=-1), 0] = 1
        return out

    def configurations(self, trajectory_idx: Sequence[int], timestep: Sequence[int]) -> np.ndarray:
        t = self.store[self.trajectory_key]
        return np.stack([np.asarray(t[int(i), int(s), :], dtype=np.float32) for i, s in zip(trajectory_idx, timestep)])


def batch_inputs(engine, store: TrajectoryStore, indices: Sequence[int], train: bool = True, random_scale: float = 0.015,
                 epoch: int = 0, trajectory_dataset: bool = False) -> Dict[str, Any]:
    """``PointCloudInstanceDataset.__getitem__`` (data_loader.py:403-417) / ``PointCloudTrajectoryDataset.__getitem__`` (:331-341)
    for a whole batch of dataset indices, with the per-item CPU work of ``get_inputs`` moved onto the device:

      configuration  = normalize(clamp(q + random_scale * N(0, 1), limits))   (train only, data_loader.py:167-180; mpn_augment_joints)
      xyz            = robot points at that configuration | 4096 obstacle surface points | 128 target gripper points (mpn_build_cloud)
      target_position = FK(trajectory[-1]).xyz,  supervision = normalize(trajectory[t + 1]) (instance dataset only)

    The noise and the sampling streams are keyed by (engine seed, dataset index, epoch): a batch is reproducible and independent of
    how the dataset is sharded over ranks or workers.  Returns torch tensors on the engine's device, batch keys of model.py:213-220."""
    import torch
    idx = np.asarray(indices, dtype=np.int64)
    if trajectory_dataset:
        traj_idx, timestep = idx, np.zeros_like(idx)
    else:
        traj_idx, timestep = np.divmod(idx, store.expert_length)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)
    q = dev(store.configurations(traj_idx, timestep))
    q_last = dev(store.configurations(traj_idx, np.full_like(traj_idx, store.expert_length - 1)))
    scene = {k: dev(v) for k, v in store.scene_rows(traj_idx).items()}
    _, eef = engine.fk(q_last)                                           # target pose = FK of the trajectory's last configuration
    if train and random_scale > 0:
        q_used, qn = engine.augment_joints(q, random_scale, sample0=dev(idx.astype(np.int64)), epoch=epoch)
    else:
        q_used, qn = q, engine.normalize(q)
    item = dict(scene)
    item["xyz"] = engine.build_cloud(scene, q_used, eef.contiguous(), problem_ids=dev(idx.astype(np.int64)), epoch=epoch if train else 0)
    item["configuration"] = qn
    item["target_position"] = eef[:, :, 3].contiguous()
    item["target_pose"] = eef
    if not trajectory_dataset:
        sup_t = np.clip(timestep + 1, 0, store.expert_length - 1)       # re-use the last point at the end (data_loader.py:405-409)
        item["supervision"] = engine.normalize(dev(store.configurations(traj_idx, sup_t)))
    return item
