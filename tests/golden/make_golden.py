"""Generates tests/golden/*.npz by running the REAL reference code in this container.

Run here only (``/root/reference`` does not exist on the GPU box):  python tests/golden/make_golden.py

* ``sdf_reference.npz``  -- ``mpinets/geometry.py`` ``TorchCuboids/TorchCylinders.sdf`` and ``.sdf_sequence`` outputs and
  the ``has_collision`` reduction of ``mpinets/model.py:301-312`` on seeded random scenes.  ``geometrout`` is not
  installable here, so ``geometrout.primitive`` is stubbed with three empty classes (geometry.py only uses them
  for type annotations and the unused ``.geometrout()`` helpers).
* ``loss_reference.npz`` -- ``mpinets/loss.py`` ``collision_loss`` and ``point_match_loss`` (the real functions, imported
  with ``robofin`` / ``mpinets.utils`` stubbed: neither is touched by these two functions) on seeded scenes and points,
  with ``torch.autograd`` gradients w.r.t. the input cloud.
* ``sparc_reference.npz`` -- ``mpinets/third_party/sparc.py`` ``sparc`` (imports as-is) on its own doctest input
  (sparc.py:87-91, -1.41403) and on seeded speed profiles shaped like rollouts (dt = 0.08, 20..150 samples).
* ``fk_reference.npz``   -- the FK known-answer pair of ``interactive_demo/mpinets_ros/nodes/interaction_node.py:54-75``
  (parsed from the file, not retyped).
"""
import importlib.util
import os
import re
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_geometry():
    prim = types.ModuleType("geometrout.primitive")
    for n in ("Sphere", "Cuboid", "Cylinder"):
        setattr(prim, n, type(n, (), {}))
    pkg = types.ModuleType("geometrout")
    pkg.primitive = prim
    sys.modules["geometrout"] = pkg
    sys.modules["geometrout.primitive"] = prim
    spec = importlib.util.spec_from_file_location("ref_geometry", os.path.join(REF, "mpinets", "geometry.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_loss(geo):
    """the real mpinets/loss.py; its module-level imports of robofin / mpinets.utils are satisfied by stubs"""
    pkg = types.ModuleType("mpinets")
    pkg.geometry = geo
    pkg.utils = types.ModuleType("mpinets.utils")
    rob = types.ModuleType("robofin"); pc = types.ModuleType("robofin.pointcloud"); pct = types.ModuleType("robofin.pointcloud.torch")
    pct.FrankaSampler = type("FrankaSampler", (), {})
    sys.modules.update({"mpinets": pkg, "mpinets.geometry": geo, "mpinets.utils": pkg.utils, "robofin": rob,
                        "robofin.pointcloud": pc, "robofin.pointcloud.torch": pct})
    spec = importlib.util.spec_from_file_location("ref_loss", os.path.join(REF, "mpinets", "loss.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def loss_fixtures(geo):
    loss = load_reference_loss(geo)
    out = {}
    for tag, yaw_only in (("yaw", True), ("free", False)):
        rng = np.random.RandomState(777 + yaw_only)
        B, M1, M2, N = 5, 6, 4, 300
        s = random_scenes(rng, B, M1, M2, yaw_only)
        # points concentrated around the primitives so that many fall inside the 3 cm margin / inside the volumes
        anchors = np.concatenate([s["cuboid_centers"], s["cylinder_centers"]], axis=1)
        pick = rng.randint(anchors.shape[1], size=(B, N))
        pts = (np.take_along_axis(anchors, pick[..., None], axis=1) + rng.normal(scale=0.25, size=(B, N, 3))).astype(np.float32)
        tp = torch.from_numpy(pts).requires_grad_(True)
        t = {k: torch.from_numpy(v) for k, v in s.items()}
        val = loss.collision_loss(tp, t["cuboid_centers"], t["cuboid_dims"], t["cuboid_quats"], t["cylinder_centers"],
                                  t["cylinder_radii"], t["cylinder_heights"], t["cylinder_quats"])
        val.backward()
        out.update({f"{tag}_{k}": v for k, v in s.items()})
        out[f"{tag}_points"] = pts
        out[f"{tag}_collision_loss"] = np.float32(val.item())
        out[f"{tag}_collision_grad"] = tp.grad.numpy().copy()
        other = (pts + rng.normal(scale=0.05, size=pts.shape)).astype(np.float32)
        ta = torch.from_numpy(pts).requires_grad_(True)
        pm = loss.point_match_loss(ta, torch.from_numpy(other))
        pm.backward()
        out[f"{tag}_other"] = other
        out[f"{tag}_point_match_loss"] = np.float32(pm.item())
        out[f"{tag}_point_match_grad"] = ta.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "loss_reference.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if "loss" in k})


def sparc_fixtures():
    spec = importlib.util.spec_from_file_location("ref_sparc", os.path.join(REF, "mpinets", "third_party", "sparc.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    t = np.arange(-1, 1, 0.01)
    move = np.exp(-5 * pow(t, 2))
    doc, _, _ = mod.sparc(move, fs=100.0)
    rng = np.random.RandomState(42)
    n_max, B = 150, 24
    profiles = np.zeros((B, n_max), np.float32); num = np.zeros(B, np.int32); sal = np.zeros(B)
    for b in range(B):
        n = int(rng.randint(20, n_max + 1)); num[b] = n
        tt = np.linspace(0, 1, n)
        v = np.sin(np.pi * tt) ** 2 * rng.uniform(0.2, 2.0) + 0.05 * rng.uniform(0, 1) * np.abs(rng.normal(size=n)) * (b % 3 != 0)
        profiles[b, :n] = v
        sal[b] = mod.sparc(profiles[b, :n].astype(np.float64), 1.0 / 0.08)[0]
    np.savez_compressed(os.path.join(HERE, "sparc_reference.npz"), doctest_move=move, doctest_sal=doc, profiles=profiles, num=num,
                        sal=sal, fs=1.0 / 0.08)
    print("sparc doctest", doc, "profiles", sal[:4])


def random_scenes(rng, B, M1, M2, yaw_only):
    def quats(n):
        if yaw_only:
            a = rng.uniform(-np.pi, np.pi, size=n)
            return np.stack([np.cos(a / 2), 0 * a, 0 * a, np.sin(a / 2)], axis=-1)
        q = rng.normal(size=(n, 4))
        return q * rng.uniform(0.5, 2.0, size=(n, 1))  # deliberately un-normalised (geometry.py:151 normalises)
    s = dict(
        cuboid_centers=rng.uniform(-1, 1, size=(B, M1, 3)),
        cuboid_dims=rng.uniform(0.02, 0.8, size=(B, M1, 3)),
        cuboid_quats=quats(B * M1).reshape(B, M1, 4),
        cylinder_centers=rng.uniform(-1, 1, size=(B, M2, 3)),
        cylinder_radii=rng.uniform(0.02, 0.3, size=(B, M2, 1)),
        cylinder_heights=rng.uniform(0.02, 0.6, size=(B, M2, 1)),
        cylinder_quats=quats(B * M2).reshape(B, M2, 4),
    )
    # zero-volume padding rows (data_loader.py:198-215): some prims masked, last scene cuboid-free, one cylinder-free
    s["cuboid_dims"][:, M1 // 2:, rng.randint(3)] = 0.0
    s["cylinder_radii"][:, M2 - 1:] = 0.0
    s["cylinder_heights"][0, :] = 0.0
    s["cuboid_dims"][B - 1] = 0.0
    s["cuboid_quats"][B - 1] = [1, 0, 0, 0]
    return {k: v.astype(np.float32) for k, v in s.items()}


def main():
    geo = load_reference_geometry()
    out = {}
    for tag, yaw_only in (("yaw", True), ("free", False)):
        rng = np.random.RandomState(20220922 + yaw_only)
        B, M1, M2, N, T, NS = 6, 7, 4, 257, 5, 9
        s = random_scenes(rng, B, M1, M2, yaw_only)
        pts = rng.uniform(-1.2, 1.2, size=(B, N, 3)).astype(np.float32)
        seq = rng.uniform(-1.2, 1.2, size=(B, T, NS, 3)).astype(np.float32)
        t = {k: torch.from_numpy(v) for k, v in s.items()}
        cub = geo.TorchCuboids(t["cuboid_centers"], t["cuboid_dims"], t["cuboid_quats"])
        cyl = geo.TorchCylinders(t["cylinder_centers"], t["cylinder_radii"], t["cylinder_heights"], t["cylinder_quats"])
        out.update({f"{tag}_{k}": v for k, v in s.items()})
        out[f"{tag}_points"] = pts
        out[f"{tag}_seq"] = seq
        out[f"{tag}_sdf_cuboids"] = cub.sdf(torch.from_numpy(pts)).numpy()
        out[f"{tag}_sdf_cylinders"] = cyl.sdf(torch.from_numpy(pts)).numpy()
        sq = torch.minimum(cub.sdf_sequence(torch.from_numpy(seq)), cyl.sdf_sequence(torch.from_numpy(seq)))
        out[f"{tag}_sdf_sequence"] = sq.numpy()
        radius = 0.06
        out[f"{tag}_has_collision_r006"] = torch.any(sq.reshape(B, -1) <= radius, dim=-1).numpy()  # model.py:309-311
    np.savez_compressed(os.path.join(HERE, "sdf_reference.npz"), **out)

    loss_fixtures(geo)
    sparc_fixtures()

    src = open(os.path.join(REF, "interactive_demo/mpinets_ros/nodes/interaction_node.py")).read()

    def arr(name):
        body = re.search(name + r"\s*=\s*(?:np\.array\(\s*)?\[(.*?)\]", src, re.S).group(1)
        return np.array([float(x) for x in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", body)])
    np.savez(os.path.join(HERE, "fk_reference.npz"), q=arr("NEUTRAL_CONFIG")[:7], xyz=arr("NEUTRAL_TARGET_XYZ"),
             xyzw=arr("NEUTRAL_TARGET_XYZW"))
    print({k: v.shape for k, v in out.items() if "sdf" in k})
    print(dict(np.load(os.path.join(HERE, "fk_reference.npz"))))


if __name__ == "__main__":
    main()
