"""CPU ORACLE (test infrastructure, NOT product code) -- Python face of ``mpn_oracle.c`` plus the
floating-point network restated with torch-CPU ops.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
import this module.  The product package ``mpinets_b200`` never does.

Restates (paths under /root/reference):
  * ``MPiNetsPointNet.forward`` / ``MotionPolicyNetwork.forward``   mpinets/model.py:75-91,360-426
  * ``PointnetSAModule`` glue (FPS -> gather -> ball query -> group -> shared MLP -> max)
                                                                  pointnet2_ops v3.2.0 (un-vendored), SURVEY App. A.1
  * ``TrainingMotionPolicyNetwork.rollout`` + validation sweep      mpinets/model.py:128-183,272-314
The integer / geometry arithmetic lives in ``mpn_oracle.c`` (bit-exact contract); the MLP arithmetic is
torch CPU fp32 (tolerance contract, 1e-5 on delta-q).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libmpn_oracle.so")
    src = os.path.join(_HERE, "mpn_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmpn_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int32))


def _scene_args(scene):
    """scene: dict with cuboid_centers[B,M1,3], cuboid_dims, cuboid_quats[B,M1,4], cylinder_centers[B,M2,3],
    cylinder_radii[B,M2,1|], cylinder_heights, cylinder_quats (keys of data_loader.py:206-235)."""
    keep = []
    cc, pcc = _f(scene["cuboid_centers"]); cd, pcd = _f(scene["cuboid_dims"]); cq, pcq = _f(scene["cuboid_quats"])
    yc, pyc = _f(scene["cylinder_centers"])
    yr, pyr = _f(np.asarray(scene["cylinder_radii"]).reshape(yc.shape[0], -1))
    yh, pyh = _f(np.asarray(scene["cylinder_heights"]).reshape(yc.shape[0], -1))
    yq, pyq = _f(scene["cylinder_quats"])
    keep += [cc, cd, cq, yc, yr, yh, yq]
    B, M1, M2 = cc.shape[0], cc.shape[1], yc.shape[1]
    return keep, B, M1, M2, (pcc, pcd, pcq, pyc, pyr, pyh, pyq)


# ----------------------------------------------------------------------------- geometry
def sincos(x):
    x, px = _f(x)
    s = np.empty_like(x); c = np.empty_like(x)
    lib().mpn_oracle_sincos(px, C.c_int(x.size), s.ctypes.data_as(C.POINTER(C.c_float)), c.ctypes.data_as(C.POINTER(C.c_float)))
    return s, c


def fk(q, prismatic=0.025):
    """q [B,7] -> (frames [B,11,3,4], right_gripper [B,3,4])"""
    q, pq = _f(q)
    B = q.shape[0]
    frames = np.empty((B, 11, 3, 4), np.float32); eef = np.empty((B, 3, 4), np.float32)
    lib().mpn_oracle_fk(pq, C.c_int(B), C.c_float(prismatic), frames.ctypes.data_as(C.POINTER(C.c_float)),
                        eef.ctypes.data_as(C.POINTER(C.c_float)))
    return frames, eef


def unnormalize(qn, limits):
    qn, pq = _f(qn); lim, pl = _f(limits)
    out = np.empty_like(qn)
    lib().mpn_oracle_unnormalize(pq, C.c_int(qn.shape[0]), pl, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def normalize(q, limits):
    q, pq = _f(q); lim, pl = _f(limits)
    out = np.empty_like(q)
    lib().mpn_oracle_normalize(pq, C.c_int(q.shape[0]), pl, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def sdf_points(scene, points, quirk=True, which=0):
    keep, B, M1, M2, ps = _scene_args(scene)
    pts, pp = _f(points)
    N = pts.shape[1]
    out = np.empty((B, N), np.float32)
    lib().mpn_oracle_sdf_points(C.c_int(B), C.c_int(N), C.c_int(M1), C.c_int(M2), *ps, C.c_int(int(quirk)), C.c_int(which),
                                pp, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def sweep_flags(scene, traj, tables, quirk=True):
    """traj [B,T,7] unnormalised -> (flags u8[B], first_step i32[B], min_margin f32[B])"""
    keep, B, M1, M2, ps = _scene_args(scene)
    traj, pt = _f(traj)
    T = traj.shape[1]
    sc, psc = _f(tables.sphere_centers); sr, psr = _f(tables.sphere_radii); sl, psl = _i(tables.sphere_links)
    flags = np.zeros(B, np.uint8); first = np.zeros(B, np.int32); mm = np.zeros(B, np.float32)
    lib().mpn_oracle_sweep_flags(C.c_int(B), C.c_int(T), C.c_int(M1), C.c_int(M2), *ps, C.c_int(int(quirk)), pt,
                                 C.c_float(tables.prismatic), C.c_int(sc.shape[0]), psc, psr, psl,
                                 flags.ctypes.data_as(C.POINTER(C.c_uint8)), first.ctypes.data_as(C.POINTER(C.c_int32)),
                                 mm.ctypes.data_as(C.POINTER(C.c_float)))
    return flags, first, mm


EVAL_COLS = 16


# ----------------------------------------------------------------------------- SPARC (third_party/sparc.py:48-140)
def sparc(movement, fs, padlevel=4, fc=10.0, amp_th=0.05):
    """spectral arc length of one speed profile, float64, line by line after sparc.py (returns only `sal`)"""
    movement = np.asarray(movement, dtype=np.float64)
    if np.allclose(movement, 0):                                       # sparc.py:95-97
        return 0.0
    nfft = int(pow(2, np.ceil(np.log2(len(movement))) + padlevel))     # :99
    f = np.arange(0, fs, fs / nfft)                                    # :102
    Mf = abs(np.fft.fft(movement, nfft))                               # :104
    Mf = Mf / max(Mf)                                                  # :105
    sel = (f <= fc).nonzero()                                          # :112
    f_sel, Mf_sel = f[sel], Mf[sel]
    inx = (Mf_sel >= amp_th).nonzero()[0]                              # :119
    rng = range(inx[0], inx[-1] + 1)
    f_sel, Mf_sel = f_sel[rng], Mf_sel[rng]
    if len(f_sel) < 2:
        return 0.0
    return float(-np.sum(np.sqrt(np.power(np.diff(f_sel) / (f_sel[-1] - f_sel[0]), 2) + np.power(np.diff(Mf_sel), 2))))   # :125-129


# ----------------------------------------------------------------------------- losses (loss.py)
def collision_loss(scene, points, margin=0.03, quirk=True, need_grad=True):
    """loss.collision_loss (loss.py:47-94): points [B,N,3] -> (loss, grad_points [B,N,3])"""
    keep, B, M1, M2, ps = _scene_args(scene)
    pts, pp = _f(points)
    N = pts.shape[1]
    loss = C.c_float(0.0)
    grad = np.zeros_like(pts) if need_grad else None
    lib().mpn_oracle_collision_loss(C.c_int(B), C.c_int(N), C.c_int(M1), C.c_int(M2), *ps, C.c_int(int(quirk)), pp, C.c_float(margin),
                                    C.byref(loss), grad.ctypes.data_as(C.POINTER(C.c_float)) if need_grad else C.POINTER(C.c_float)())
    return float(loss.value), grad


def point_match_loss(a, b, need_grad=True):
    """loss.point_match_loss (loss.py:31-44) -> (loss, grad_a)"""
    a, pa = _f(a); b, pb = _f(b)
    loss = C.c_float(0.0)
    grad = np.zeros_like(a) if need_grad else None
    lib().mpn_oracle_point_match_loss(C.c_size_t(a.size), pa, pb, C.byref(loss),
                                      grad.ctypes.data_as(C.POINTER(C.c_float)) if need_grad else C.POINTER(C.c_float)())
    return float(loss.value), grad


def fixed_robot_points(q, tables, n, seed):
    """FrankaSampler(num_fixed_points=n, with_base_link=False).sample(q) stand-in (loss.py:141-153): (points [B,n,3], table rows [n])"""
    q, pq = _f(q)
    lp, plp = _f(tables.link_points); lid, plid = _i(tables.link_ids)
    out = np.empty((q.shape[0], n, 3), np.float32); idx = np.empty(n, np.int32)
    lib().mpn_oracle_fixed_robot_points(pq, C.c_int(q.shape[0]), C.c_float(tables.prismatic), C.c_int(lp.shape[0]), plp, plid, C.c_int(n),
                                        C.c_uint32(seed & 0xFFFFFFFF), C.c_uint32(seed >> 32),
                                        out.ctypes.data_as(C.POINTER(C.c_float)), idx.ctypes.data_as(C.POINTER(C.c_int32)))
    return out, idx


def bc_collision_losses(scene, input_norm, target_norm, tables, seed, n=1024, margin=0.03, w_collision=1.0, w_bc=1.0, quirk=True):
    """CollisionAndBCLossContainer.__call__ (loss.py:111-166) -> (losses [2], grad_input [B,7])"""
    keep, B, M1, M2, ps = _scene_args(scene)
    qi, pqi = _f(input_norm); qt, pqt = _f(target_norm)
    lim, pl = _f(tables.joint_limits)
    lp, plp = _f(tables.link_points); lid, plid = _i(tables.link_ids)
    losses = np.zeros(2, np.float32); grad = np.zeros((B, 7), np.float32)
    lib().mpn_oracle_bc_collision_losses(C.c_int(B), C.c_int(M1), C.c_int(M2), *ps, C.c_int(int(quirk)), pqi, pqt, pl,
                                         C.c_float(tables.prismatic), C.c_int(lp.shape[0]), plp, plid, C.c_int(n),
                                         C.c_uint32(seed & 0xFFFFFFFF), C.c_uint32(seed >> 32), C.c_float(margin),
                                         C.c_float(w_collision), C.c_float(w_bc), losses.ctypes.data_as(C.POINTER(C.c_float)),
                                         grad.ctypes.data_as(C.POINTER(C.c_float)))
    return losses, grad


def _volume_args(vol, B):
    """optional region-test primitive lists -> (keep, n_cuboids, n_cylinders, 7 pointers)"""
    nullp = C.POINTER(C.c_float)()
    if vol is None:
        return [], 0, 0, (nullp,) * 7
    keep, n1, n2 = [], 0, 0
    ptrs = [nullp] * 7
    if vol.get("cuboid_centers") is not None:
        cc, p0 = _f(vol["cuboid_centers"]); cd, p1 = _f(vol["cuboid_dims"]); cq, p2 = _f(vol["cuboid_quats"])
        keep += [cc, cd, cq]; ptrs[0:3] = [p0, p1, p2]; n1 = cc.shape[1]
        assert cc.shape[0] == B
    if vol.get("cylinder_centers") is not None:
        yc, p3 = _f(vol["cylinder_centers"])
        yr, p4 = _f(np.asarray(vol["cylinder_radii"]).reshape(B, -1)); yh, p5 = _f(np.asarray(vol["cylinder_heights"]).reshape(B, -1))
        yq, p6 = _f(vol["cylinder_quats"])
        keep += [yc, yr, yh, yq]; ptrs[3:7] = [p3, p4, p5, p6]; n2 = yc.shape[1]
    return keep, n1, n2, tuple(ptrs)


def evaluate(scene, traj, target, tables, num_poses=None, target_volume=None, negative_volumes=None, quirk=True):
    """Evaluator.evaluate_trajectory subset (metrics.py:311-322,340-384,411-434,487-523): traj [B,T1,7] -> [B,16]"""
    keep, B, M1, M2, ps = _scene_args(scene)
    traj, pt = _f(traj); target, ptg = _f(np.asarray(target).reshape(B, 12))
    T1 = traj.shape[1]
    lim, pl = _f(tables.joint_limits)
    sc, psc = _f(tables.sphere_centers); sr, psr = _f(tables.sphere_radii); sl, psl = _i(tables.sphere_links)
    if num_poses is not None:
        npz, pn = _i(num_poses)
    else:
        pn = C.POINTER(C.c_int32)()
    k1, v1, v2, pv = _volume_args(target_volume, B)
    k2, n1, n2, pnv = _volume_args(negative_volumes, B)
    out = np.zeros((B, EVAL_COLS), np.float32)
    lib().mpn_oracle_evaluate(C.c_int(B), C.c_int(T1), C.c_int(M1), C.c_int(M2), *ps, C.c_int(int(quirk)), pt, pn, ptg, pl,
                              C.c_float(tables.prismatic), C.c_int(sc.shape[0]), psc, psr, psl,
                              C.c_int(v1), C.c_int(v2), *pv, C.c_int(n1), C.c_int(n2), *pnv,
                              out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def spheres(q, tables):
    q, pq = _f(q)
    sc, psc = _f(tables.sphere_centers); sl, psl = _i(tables.sphere_links)
    out = np.empty((q.shape[0], sc.shape[0], 3), np.float32)
    lib().mpn_oracle_spheres(pq, C.c_int(q.shape[0]), C.c_float(tables.prismatic), C.c_int(sc.shape[0]), psc, psl,
                             out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def philox(c, k):
    out = (C.c_uint32 * 4)()
    lib().mpn_oracle_philox(*[C.c_uint32(int(v)) for v in c], C.c_uint32(int(k[0])), C.c_uint32(int(k[1])), out)
    return np.array(list(out), dtype=np.uint32)


def feistel(n, key, count):
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(count, np.uint32)
    lib().mpn_oracle_feistel(C.c_uint32(n), key.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint32(count),
                             out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


def sample_obstacles(scene, n, seed, problem0=0, return_prims=False):
    keep, B, M1, M2, ps = _scene_args(scene)
    out = np.zeros((B, n, 4), np.float32)
    prim = np.zeros((B, n), np.int32)
    lib().mpn_oracle_sample_obstacles(C.c_int(B), C.c_int(n), C.c_int(M1), C.c_int(M2), *ps, C.c_uint64(seed),
                                      C.c_uint32(problem0), out.ctypes.data_as(C.POINTER(C.c_float)),
                                      prim.ctypes.data_as(C.POINTER(C.c_int32)))
    return (out, prim) if return_prims else out


def sample_robot(q, tables, n, seed, step, cloud=None):
    """Writes rows [0,n) (xyz only) of cloud [B,rows,4]; returns the cloud."""
    q, pq = _f(q)
    B = q.shape[0]
    if cloud is None:
        cloud = np.zeros((B, n, 4), np.float32)
    assert cloud.dtype == np.float32 and cloud.flags.c_contiguous
    lp, plp = _f(tables.link_points); li, pli = _i(tables.link_ids)
    lib().mpn_oracle_sample_robot(pq, C.c_int(B), C.c_float(tables.prismatic), C.c_int(lp.shape[0]), plp, pli, C.c_int(n),
                                  C.c_uint64(seed), C.c_uint32(step), cloud.ctypes.data_as(C.POINTER(C.c_float)),
                                  C.c_int(cloud.shape[1]))
    return cloud


def augment_joints(q, tables, scale, seed, ids=None, epoch=0):
    """data_loader.py:167-180 -> (clamp(q + scale * N(0,1), limits), its normalisation)"""
    q, pq = _f(q)
    lim, pl = _f(tables.joint_limits)
    B = q.shape[0]
    if ids is not None:
        ida = np.ascontiguousarray(np.asarray(ids, dtype=np.int64) & 0xFFFFFFFF, dtype=np.uint32)
        pid = ida.ctypes.data_as(C.POINTER(C.c_uint32))
    else:
        pid = C.POINTER(C.c_uint32)()
    out = np.empty_like(q); outn = np.empty_like(q)
    lib().mpn_oracle_augment_joints(pq, C.c_int(B), C.c_float(scale), pid, C.c_uint32(epoch), pl, C.c_uint64(seed),
                                    out.ctypes.data_as(C.POINTER(C.c_float)), outn.ctypes.data_as(C.POINTER(C.c_float)))
    return out, outn


def clean_point_cloud(xyz, rgba, n_out, seed, cloud_id=0):
    """planning_node.py:187-228 -> (kept count, xyz [n_out,3] | None, rgba [n_out,4] | None)"""
    xyz, px = _f(xyz)
    out = np.zeros((n_out, 3), np.float32)
    if rgba is not None:
        rgba, pr = _f(rgba)
        outc = np.zeros((n_out, 4), np.float32); pc = outc.ctypes.data_as(C.POINTER(C.c_float))
    else:
        pr = pc = C.POINTER(C.c_float)(); outc = None
    fn = lib().mpn_oracle_clean_point_cloud
    fn.restype = C.c_int
    kept = fn(px, pr, C.c_int(xyz.shape[0]), C.c_int(n_out), C.c_uint64(seed), C.c_uint32(cloud_id), out.ctypes.data_as(C.POINTER(C.c_float)), pc)
    return (kept, out, outc) if kept >= n_out else (kept, None, None)


def sample_end_effector(poses, tables, n, seed, problem0=0):
    """poses [B,3,4] right_gripper -> [B,n,3] (FrankaSampler.sample_end_effector; the keyed subset of build_cloud's target rows)"""
    B = np.asarray(poses).shape[0]
    ps, pp = _f(np.asarray(poses).reshape(B, 12))
    ee, pee = _f(tables.ee_points)
    out = np.zeros((B, n, 3), np.float32)
    lib().mpn_oracle_sample_end_effector(C.c_int(B), pp, C.c_int(ee.shape[0]), pee, C.c_int(n), C.c_uint64(seed), C.c_uint32(problem0),
                                         out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def build_cloud(q0, target, scene, tables, seed, n_robot=2048, n_obs=4096, n_tgt=128, problem0=0):
    """q0 [B,7] unnormalised, target [B,3,4] right_gripper pose -> cloud [B,N,4] (run_inference.py:93-134)."""
    keep, B, M1, M2, ps = _scene_args(scene)
    q0, pq = _f(q0); tg, ptg = _f(np.asarray(target).reshape(B, 12))
    lp, plp = _f(tables.link_points); li, pli = _i(tables.link_ids); ee, pee = _f(tables.ee_points)
    cloud = np.zeros((B, n_robot + n_obs + n_tgt, 4), np.float32)
    lib().mpn_oracle_build_cloud(C.c_int(B), pq, ptg, C.c_float(tables.prismatic), C.c_int(lp.shape[0]), plp, pli,
                                 C.c_int(ee.shape[0]), pee, C.c_int(n_robot), C.c_int(n_obs), C.c_int(n_tgt),
                                 C.c_int(M1), C.c_int(M2), *ps, C.c_uint64(seed), C.c_uint32(problem0),
                                 cloud.ctypes.data_as(C.POINTER(C.c_float)))
    return cloud


def build_cloud_from_points(q0, target, obs_points, obs_count, tables, seed, n_robot=2048, n_obs=4096, n_tgt=128, problem0=0):
    """run_inference.make_point_cloud_from_problem (run_inference.py:58-90): obstacle rows = random subset of a given cloud"""
    q0, pq = _f(q0); B = q0.shape[0]
    tg, ptg = _f(np.asarray(target).reshape(B, 12))
    op, pop = _f(obs_points); oc, poc = _i(obs_count)
    lp, plp = _f(tables.link_points); li, pli = _i(tables.link_ids); ee, pee = _f(tables.ee_points)
    cloud = np.zeros((B, n_robot + n_obs + n_tgt, 4), np.float32)
    lib().mpn_oracle_build_cloud_from_points(C.c_int(B), pq, ptg, C.c_float(tables.prismatic), C.c_int(lp.shape[0]), plp, pli,
                                             C.c_int(ee.shape[0]), pee, C.c_int(n_robot), C.c_int(n_obs), C.c_int(n_tgt),
                                             C.c_int(op.shape[1]), pop, poc, C.c_uint64(seed), C.c_uint32(problem0),
                                             cloud.ctypes.data_as(C.POINTER(C.c_float)))
    return cloud


# ----------------------------------------------------------------------------- pointnet2_ops restatement
def fps(xyz, npoint):
    """xyz [B,N,3|4] -> idx i32 [B,npoint]"""
    xyz, px = _f(xyz)
    B, N, stride = xyz.shape
    idx = np.zeros((B, npoint), np.int32)
    lib().mpn_oracle_fps(C.c_int(B), C.c_int(N), C.c_int(stride), px, C.c_int(npoint), idx.ctypes.data_as(C.POINTER(C.c_int32)))
    return idx


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, px = _f(xyz); new_xyz, pn = _f(new_xyz)
    B, N, stride = xyz.shape
    m = new_xyz.shape[1]
    idx = np.zeros((B, m, nsample), np.int32)
    lib().mpn_oracle_ball_query(C.c_int(B), C.c_int(N), C.c_int(stride), px, C.c_int(m), pn, C.c_float(radius),
                                C.c_int(nsample), idx.ctypes.data_as(C.POINTER(C.c_int32)))
    return idx


def distinct_neighbour_counts(ball_idx):
    """ball_idx [B,m,nsample] (pointnet2 first-hit padding) -> distinct neighbours per group [B,m]: the rows of a group that are not
    copies (the max-pool of PointnetSAModule, model.py:365-382, cannot see the copies)."""
    a = np.sort(np.asarray(ball_idx), axis=-1)
    return 1 + (a[..., 1:] != a[..., :-1]).sum(-1)


def packed_tile_count(counts, per_round=4, gran=32, tile_rows=128):
    """Tiles the fused SA kernels issue for groups with `counts` distinct rows (1-D, one problem): rounds of `per_round` consecutive
    centroids, each taking ceil(count / gran) units of `gran` rows, first fit in centroid order, never straddling a tile
    (mpinets_b200/csrc/tc_common.cuh: pack_round / pack_round8) -- test infrastructure, mirrors the device rule."""
    units_per_tile = tile_rows // gran
    tiles = 0
    counts = np.asarray(counts)
    for r0 in range(0, len(counts), per_round):
        fill, n = 0, 1
        for h in counts[r0:r0 + per_round]:
            q = max(1, -(-int(min(h, tile_rows)) // gran))
            if fill + q > units_per_tile:
                n, fill = n + 1, 0
            fill += q
        tiles += n
    return tiles


def gather_operation(feat, idx):
    """feat [B,C,N], idx [B,m] -> [B,C,m]"""
    B = feat.shape[0]
    return np.stack([feat[b][:, idx[b]] for b in range(B)])


def grouping_operation(feat, idx):
    """feat [B,C,N], idx [B,m,ns] -> [B,C,m,ns]"""
    B = feat.shape[0]
    return np.stack([feat[b][:, idx[b]] for b in range(B)])


# ----------------------------------------------------------------------------- network (torch CPU)
SA_SPECS = (
    dict(npoint=512, radius=0.05, nsample=128, mlp=(4, 64, 64, 64)),       # model.py:365-373 (+3 for use_xyz)
    dict(npoint=128, radius=0.3, nsample=128, mlp=(67, 128, 128, 256)),    # model.py:374-382
    dict(npoint=None, radius=None, nsample=None, mlp=(259, 512, 512, 1024)),  # model.py:383 (GroupAll)
)


def reference_state_dict(seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Default-initialised weights of the architecture at model.py:41-66,360-393 with the reference's
    state-dict key names (the Zenodo checkpoint is not available offline)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()

    def lin(name, cin, cout, conv=False):
        bound = 1.0 / np.sqrt(cin)
        w = (torch.rand((cout, cin), generator=g) * 2 - 1) * bound
        b = (torch.rand((cout,), generator=g) * 2 - 1) * bound
        sd[name + ".weight"] = w.reshape(cout, cin, 1, 1) if conv else w
        sd[name + ".bias"] = b

    for s, spec in enumerate(SA_SPECS):
        for l in range(3):
            lin(f"point_cloud_encoder.SA_modules.{s}.mlps.0.{2 * l}", spec["mlp"][l], spec["mlp"][l + 1], conv=True)
    lin("point_cloud_encoder.fc_layer.0", 1024, 4096)
    sd["point_cloud_encoder.fc_layer.1.weight"] = 1 + 0.1 * (torch.rand(4096, generator=g) - 0.5)
    sd["point_cloud_encoder.fc_layer.1.bias"] = 0.1 * (torch.rand(4096, generator=g) - 0.5)
    lin("point_cloud_encoder.fc_layer.3", 4096, 2048)
    sd["point_cloud_encoder.fc_layer.4.weight"] = 1 + 0.1 * (torch.rand(2048, generator=g) - 0.5)
    sd["point_cloud_encoder.fc_layer.4.bias"] = 0.1 * (torch.rand(2048, generator=g) - 0.5)
    lin("point_cloud_encoder.fc_layer.6", 2048, 2048)
    for i, (a, b) in zip((0, 2, 4, 6, 8), ((7, 32), (32, 64), (64, 128), (128, 128), (128, 64))):
        lin(f"feature_encoder.{i}", a, b)
    for i, (a, b) in zip((0, 2, 4, 6), ((2112, 512), (512, 256), (256, 128), (128, 7))):
        lin(f"decoder.{i}", a, b)
    return sd


def _round_bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _dense(x, w, b, emulate_bf16, dtype, round_bias=False):
    """x [..., Cin] @ w[Cout,Cin]^T + b, with optional bf16 operand rounding (fp32 products are exact,
    accumulation in `dtype`).  round_bias: the bias is itself a bf16 GEMM operand (SA1 of the tensor-core mode folds
    it into the contraction as an extra K column)."""
    if emulate_bf16:
        x = _round_bf16(x); w = _round_bf16(w)
        if round_bias:
            b = _round_bf16(b)
    return (x.to(dtype) @ w.to(dtype).t() + b.to(dtype))


def sa_module(xyz, feats, spec, weights, emulate_bf16=False, dtype=torch.float32, return_aux=False, round_bias=None, pool_idx=None):
    """xyz np[B,N,3] fp32, feats torch[B,N,C] (row-major per point) -> (new_xyz np[B,m,3] | None, feats torch[B,m,Cout])

    Shared-MLP rows are [dx,dy,dz, features...] in the Conv2d input-channel order (QueryAndGroup, use_xyz=True)."""
    B, N, _ = xyz.shape
    aux = {}
    if spec["npoint"] is not None:
        idx = fps(xyz, spec["npoint"])
        new_xyz = np.stack([xyz[b][idx[b]] for b in range(B)])
        bq = ball_query(spec["radius"], spec["nsample"], xyz, new_xyz)
        aux.update(fps_idx=idx, ball_idx=bq)
        bqt = torch.from_numpy(bq.astype(np.int64))
        xyz_t = torch.from_numpy(xyz)
        g_xyz = torch.stack([xyz_t[b][bqt[b]] for b in range(B)]) - torch.from_numpy(new_xyz)[:, :, None, :]  # [B,m,ns,3]
        g_f = torch.stack([feats[b][bqt[b]] for b in range(B)])                                            # [B,m,ns,C]
        x = torch.cat([g_xyz, g_f.to(torch.float32)], dim=-1)
    else:
        new_xyz = None
        x = torch.cat([torch.from_numpy(xyz), feats.to(torch.float32)], dim=-1)[:, None]                      # [B,1,N,3+C]
    x = x.to(dtype)
    if round_bias is None:   # biases that the bf16 mode folds into the contraction (see _dense): SA1 all layers, SA2 layer 1
        round_bias = {512: (True, True, True), 128: (True, False, False)}.get(spec["npoint"], (False,) * 3) if emulate_bf16 else (False,) * 3
    for (w, b), rb in zip(weights, round_bias):
        x = torch.relu(_dense(x, w.reshape(w.shape[0], -1), b, emulate_bf16, dtype, rb))
    if pool_idx is None:
        out = x.max(dim=2).values                                                                           # [B,m,Cout]
    else:
        # replay a given max-pool routing (row per channel): same forward value when the routing is valid, and the
        # backward sends each channel's gradient to exactly that row -- removes near-tie argmax flips from gradient parity
        pi = torch.as_tensor(np.asarray(pool_idx), dtype=torch.int64).reshape(x.shape[0], x.shape[1], 1, x.shape[3])
        out = x.gather(2, pi)[:, :, 0]
        aux.update(pool_gap=(x.max(dim=2).values - out).detach(), pool_argmax=x.argmax(dim=2).detach())
    if emulate_bf16 and spec["npoint"] is not None:
        out = _round_bf16(out.float()).to(dtype)   # hand-off tensors are stored in bf16 in the tensor-core mode
    return (new_xyz, out, aux) if return_aux else (new_xyz, out)


def encoder_forward(sd, cloud, emulate_bf16=False, dtype=torch.float32, return_aux=False, pool_idx=None):
    """cloud np[B,N,4] -> pc encoding torch[B,2048] (model.py:409-426)"""
    cloud = np.ascontiguousarray(cloud, dtype=np.float32)
    xyz = np.ascontiguousarray(cloud[..., :3])
    feats = torch.from_numpy(np.ascontiguousarray(cloud[..., 3:]))
    auxs = []
    for s, spec in enumerate(SA_SPECS):
        ws = [(sd[f"point_cloud_encoder.SA_modules.{s}.mlps.0.{2 * l}.weight"],
               sd[f"point_cloud_encoder.SA_modules.{s}.mlps.0.{2 * l}.bias"]) for l in range(3)]
        r = sa_module(xyz, feats, spec, ws, emulate_bf16, dtype, return_aux=True, pool_idx=None if pool_idx is None else pool_idx[s])
        xyz, feats = r[0], r[1]
        auxs.append(dict(r[2], new_xyz=r[0], feats=r[1]))
    x = feats[:, 0]  # [B,1024]
    p = "point_cloud_encoder.fc_layer."
    x = _dense(x, sd[p + "0.weight"], sd[p + "0.bias"], emulate_bf16, dtype)
    x = F.leaky_relu(F.group_norm(x, 16, sd[p + "1.weight"].to(dtype), sd[p + "1.bias"].to(dtype), eps=1e-5), 0.01)
    x = _dense(x, sd[p + "3.weight"], sd[p + "3.bias"], emulate_bf16, dtype)
    x = F.leaky_relu(F.group_norm(x, 16, sd[p + "4.weight"].to(dtype), sd[p + "4.bias"].to(dtype), eps=1e-5), 0.01)
    x = _dense(x, sd[p + "6.weight"], sd[p + "6.bias"], emulate_bf16, dtype)
    return (x, auxs) if return_aux else x


def policy_forward(sd, cloud, q_norm, emulate_bf16=False, dtype=torch.float32, return_aux=False, pool_idx=None):
    """MotionPolicyNetwork.forward (model.py:75-91): -> delta q (normalised) torch[B,7]"""
    enc = encoder_forward(sd, cloud, emulate_bf16, dtype, return_aux, pool_idx)
    aux = None
    if return_aux:
        enc, aux = enc
    x = torch.as_tensor(np.asarray(q_norm), dtype=dtype)
    for i in (0, 2, 4, 6, 8):
        x = _dense(x, sd[f"feature_encoder.{i}.weight"], sd[f"feature_encoder.{i}.bias"], False, dtype)
        if i != 8:
            x = F.leaky_relu(x, 0.01)
    x = torch.cat([enc, x], dim=1)
    for i in (0, 2, 4, 6):
        x = _dense(x, sd[f"decoder.{i}.weight"], sd[f"decoder.{i}.bias"], emulate_bf16 and i == 0, dtype)
        if i != 6:
            x = F.leaky_relu(x, 0.01)
    return (x, aux) if return_aux else x


def rollout(sd, cloud, q_norm, tables, steps, seed, emulate_bf16=False, n_robot=2048):
    """TrainingMotionPolicyNetwork.rollout (model.py:128-183), lock-step, unnormalize=True.
    Mutates `cloud` in place like the reference (model.py:181).  Returns traj np[B,steps+1,7] (unnormalised)."""
    q = np.ascontiguousarray(q_norm, dtype=np.float32).copy()
    traj = [unnormalize(q, tables.joint_limits)]
    for i in range(steps):
        dq = policy_forward(sd, cloud, q, emulate_bf16).to(torch.float32).numpy()
        q = np.clip(q + dq, -1.0, 1.0).astype(np.float32)
        qu = unnormalize(q, tables.joint_limits)
        traj.append(qu)
        sample_robot(qu, tables, n_robot, seed, i + 1, cloud)
    return np.stack(traj, axis=1)


# ----------------------------------------------------------------------------- training step (model.py:185-240)
def train_step_grads(sd, cloud, q_norm, supervision, scene, tables, seed, n_points=1024, margin=0.03, w_collision=5.0, w_bc=1.0,
                     dtype=torch.float32, pool_idx=None, return_aux=False):
    """TrainingMotionPolicyNetwork.training_step up to the parameter gradients.

    y_hat = clamp(q + net(xyz, q), -1, 1) (model.py:202); (collision, point match) = CollisionAndBCLossContainer(y_hat, ...,
    supervision) (loss.py:111-166, restated in C with its analytic d loss / d y_hat); the network backward is torch.autograd
    over the torch restatement above (index gather = pointnet2 grouping, amax = max_pool2d).
    pool_idx (optional, 3 arrays): replay this max-pool routing instead of torch's own argmax (see sa_module).
    Returns (losses np[2], y_hat np[B,7], grads {state-dict key: np array in the key's shape}, g_y np[B,7][, aux])."""
    p = OrderedDict((k, v.detach().clone().to(dtype).requires_grad_(True)) for k, v in sd.items())
    q = torch.as_tensor(np.asarray(q_norm), dtype=dtype)
    dq = policy_forward(p, cloud, np.asarray(q_norm, dtype=np.float32), False, dtype, return_aux, pool_idx)
    aux = None
    if return_aux:
        dq, aux = dq
    y_hat = torch.clamp(q + dq, min=-1, max=1)
    yh = y_hat.detach().to(torch.float32).numpy()
    losses, g_y = bc_collision_losses(scene, yh, np.asarray(supervision, dtype=np.float32), tables, seed, n_points, margin,
                                      w_collision, w_bc)
    y_hat.backward(torch.from_numpy(g_y).to(dtype))
    grads = OrderedDict((k, (v.grad if v.grad is not None else torch.zeros_like(v)).detach().to(torch.float64).numpy()) for k, v in p.items())
    return (losses, yh, grads, g_y, aux) if return_aux else (losses, yh, grads, g_y)


def adam_reference(params, grads, steps, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, clip_norm=1.0):
    """torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (model.py:68-73, run_training.py:112) applied `steps` times with
    the same gradients: params / grads {key: np array} -> ({key: np array}, [grad norms])"""
    # float64: torch's CPU float32 norm reduction is itself ~1.6e-4 off the exact norm on 19 M elements
    ps = [torch.nn.Parameter(torch.as_tensor(np.asarray(v), dtype=torch.float64).clone()) for v in params.values()]
    opt = torch.optim.Adam(ps, lr=lr, betas=betas, eps=eps)
    norms = []
    for _ in range(steps):
        for p_, g in zip(ps, grads.values()):
            p_.grad = torch.as_tensor(np.asarray(g), dtype=torch.float64).reshape(p_.shape).clone()
        norms.append(float(torch.nn.utils.clip_grad_norm_(ps, clip_norm)) if clip_norm and clip_norm > 0 else 0.0)
        opt.step()
    return OrderedDict((k, p_.detach().numpy()) for k, p_ in zip(params.keys(), ps)), norms


# ----------------------------------------------------------------------------- depth-camera clouds (run_inference.py:194-257)
# the evaluation cameras of run_inference.py:215-243 (world->camera = SE3(xyz, quaternion wxyz).inverse there; these are the
# un-inverted camera->world poses)
EVAL_CAMERAS = {
    "dresser": ([0.08307640315968651, 1.986952324350807, 0.9996085854670145],
                [-0.10162310189063647, -0.06726290364234049, 0.5478233048853433, 0.8276702686337273]),
    "cubby": ([0.08307640315968651, 1.986952324350807, 0.9996085854670145],
              [-0.10162310189063647, -0.06726290364234049, 0.5478233048853433, 0.8276702686337273]),
    "tabletop": ([1.5031788593125708, -1.817341016921562, 1.278088299149147],
                 [0.8687241016192855, 0.4180885960330695, 0.11516106409944685, 0.23928704613569252]),
}


def render_depth_cloud(scene, camera, width, height, fov_y_deg=60.0, near=0.01, far=10.0, quirk=True):
    """camera np[3,4] or [B,3,4] camera->world (GL frame: y up, looks along -z) -> (points [B, W*H, 3] hits first in pixel order, counts i32 [B])"""
    keep, B, M1, M2, ps = _scene_args(scene)
    cam, pc = _f(np.asarray(camera).reshape(-1, 12))
    per = int(cam.shape[0] > 1)
    ty = np.tan(np.radians(fov_y_deg) / 2.0)
    pts = np.zeros((B, width * height, 3), np.float32); cnt = np.zeros(B, np.int32)
    lib().mpn_oracle_render_depth_cloud(C.c_int(B), C.c_int(M1), C.c_int(M2), *ps, C.c_int(int(quirk)), pc, C.c_int(per),
                                        C.c_int(width), C.c_int(height), C.c_float(ty * width / height), C.c_float(ty),
                                        C.c_float(near), C.c_float(far), pts.ctypes.data_as(C.POINTER(C.c_float)),
                                        cnt.ctypes.data_as(C.POINTER(C.c_int32)))
    return pts, cnt
