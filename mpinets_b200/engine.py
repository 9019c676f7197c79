"""Host-side face of the C ABI: one :class:`Engine` per GPU.  PyTorch is used only for device memory and streams.

Every method validates its tensors the way the reference's CUDA extension does (CUDA, contiguous, float32 / int32;
violations raise ``RuntimeError``), allocates outputs with torch, and launches on the *current* torch stream without
synchronising the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .franka import RobotTables, default_tables

SCENE_KEYS = ("cuboid_centers", "cuboid_dims", "cuboid_quats", "cylinder_centers", "cylinder_radii",
              "cylinder_heights", "cylinder_quats")  # batch-dict keys of mpinets/data_loader.py:206-235


def _check(t: torch.Tensor, name: str, dtype=torch.float32, device: Optional[torch.device] = None):
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")  # same contract as pointnet2_ops ("CPU not supported")
    if device is not None and t.device != device:
        raise RuntimeError(f"{name} is on {t.device}, engine is on {device}")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    return t


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


class Engine:
    def __init__(self, device: int = 0, n_robot: int = 2048, n_obstacle: int = 4096, n_target: int = 128,
                 max_cuboids: int = 40, max_cylinders: int = 40, quirk_frames: bool = True,
                 seed: int = 0x4D50694E, tables: Optional[RobotTables] = None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.MpnError("mpinets_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        self.cfg = _lib.MpnConfig(n_robot, n_obstacle, n_target, max_cuboids, max_cylinders, int(quirk_frames), seed)
        self.n_points = n_robot + n_obstacle + n_target
        self._ctx = C.c_void_p()
        _lib.check(self.lib.mpn_ctx_create(device, C.byref(self.cfg), C.byref(self._ctx)))
        self.tables = tables if tables is not None else default_tables(max(4096, n_robot))
        self.set_tables(self.tables)
        self._weights_loaded = False
        self.weights_version = 0      # bumped by every load_state_dict: optimiser state in the library is reset with it
        self._weights_owner = None    # weakref to the nn.Module whose parameters the context currently holds (model.py)

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.mpn_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ setup
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_tables(self, t: RobotTables):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        i = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        lim, lp, li, ee = f(t.joint_limits), f(t.link_points), i(t.link_ids), f(t.ee_points)
        sc, sr, sl = f(t.sphere_centers), f(t.sphere_radii), i(t.sphere_links)
        _lib.check(self.lib.mpn_set_robot_tables(
            self._ctx, lim.ctypes.data, lp.shape[0], lp.ctypes.data, li.ctypes.data, ee.shape[0], ee.ctypes.data,
            sc.shape[0], sc.ctypes.data, sr.ctypes.data, sl.ctypes.data, float(t.prismatic)))
        self.tables = t

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        """Takes a reference ``MotionPolicyNetwork`` state dict (Lightning ``.ckpt['state_dict']`` keys)."""
        for name, v in sd.items():
            if not (name.startswith("point_cloud_encoder.") or name.startswith("feature_encoder.") or name.startswith("decoder.")):
                continue
            a = np.ascontiguousarray(v.detach().to("cpu", torch.float32).numpy())
            shape = (C.c_int64 * a.ndim)(*a.shape)
            _lib.check(self.lib.mpn_load_weight(self._ctx, name.encode(), a.ctypes.data, shape, a.ndim))
        _lib.check(self.lib.mpn_weights_finalize(self._ctx))
        self._weights_loaded = True
        self.weights_version += 1
        self._weights_owner = None

    def reserve(self, max_batch: int):
        _lib.check(self.lib.mpn_reserve(self._ctx, int(max_batch)))

    @property
    def launch_count(self) -> int:
        return int(self.lib.mpn_launch_count(self._ctx))

    def profile(self, enable: bool = True):
        _lib.check(self.lib.mpn_profile(self._ctx, int(enable)))

    def profile_read(self) -> Dict[str, Dict[str, float]]:
        """Synchronises; returns {stage: {"ms": total, "launches": n}} since the last read (CUDA events on the stream)."""
        ms = (C.c_float * len(_lib.STAGES))()
        n = (C.c_int64 * len(_lib.STAGES))()
        _lib.check(self.lib.mpn_profile_read(self._ctx, ms, n))
        return {name: {"ms": float(ms[i]), "launches": int(n[i])} for i, name in enumerate(_lib.STAGES)}

    def _scene(self, scene: Dict[str, torch.Tensor], B: int):
        s = _lib.MpnScene()
        keep = []
        for k in SCENE_KEYS:
            t = _check(scene[k], k, device=self.device)
            if t.shape[0] != B:
                raise RuntimeError(f"{k} has batch {t.shape[0]}, expected {B}")
            m = self.cfg.max_cuboids if k.startswith("cuboid") else self.cfg.max_cylinders
            if t.shape[1] != m:
                raise RuntimeError(f"{k} has {t.shape[1]} primitive rows, engine was built for {m}")
            keep.append(t)
            setattr(s, k, t.data_ptr())
        return s, keep

    def _empty(self, *shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def tc_error(self) -> bool:
        v = C.c_int(0)
        _lib.check(self.lib.mpn_tc_error(self._ctx, C.byref(v)))
        return bool(v.value)

    def sa_tile_counts(self, reset: bool = False):
        """(SA1, SA2) 128-row MMA tiles issued by the tensor-core set-abstraction kernels since the last reset (mpn_sa_tile_counts)."""
        v = (C.c_uint64 * 2)()
        _lib.check(self.lib.mpn_sa_tile_counts(self._ctx, v, 1 if reset else 0))
        return int(v[0]), int(v[1])

    def tc_selftest(self, a: torch.Tensor, b: torch.Tensor, mode: int):
        """D = A[128,K] @ B[N,K]^T on one CTA via tcgen05 (bf16 in, fp32 out); returns (D, timed_out)."""
        _check(a, "a", torch.bfloat16, self.device); _check(b, "b", torch.bfloat16, self.device)
        N, K = b.shape
        d = torch.zeros(128 * N + 16, device=self.device)
        st = torch.zeros(1, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.mpn_tc_selftest(self._ctx, self.stream, _p(a), _p(b), _p(d), N, K, mode, _p(st)))
        if mode & 0x100:
            return d[128 * N:128 * N + 8].cpu().tolist(), bool(st.item())
        return d[:128 * N].view(128, N), bool(st.item())

    def tc_gemm_selftest(self, a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, split: bool = True) -> torch.Tensor:
        """a [M,K] @ w [N,K]^T + bias through the TMA / tcgen05 row GEMM (fp32 in / out; bf16 or split-bf16 operands inside)"""
        _check(a, "a", device=self.device); _check(w, "w", device=self.device); _check(bias, "bias", device=self.device)
        M, K = a.shape
        N = w.shape[0]
        out = self._empty(M, N)
        _lib.check(self.lib.mpn_tc_gemm_selftest(self._ctx, self.stream, _p(a), _p(w), _p(bias), M, N, K, _p(out), int(split)))
        return out

    # ------------------------------------------------------------------ pointnet2_ops
    def fps(self, xyz: torch.Tensor, npoint: int, return_xyz: bool = False):
        _check(xyz, "xyz", device=self.device)
        B, N, stride = xyz.shape
        idx = self._empty(B, npoint, dtype=torch.int32)
        new_xyz = self._empty(B, npoint, 3) if return_xyz else None
        _lib.check(self.lib.mpn_fps(self._ctx, self.stream, _p(xyz), B, N, stride, npoint, _p(idx), _p(new_xyz)))
        return (idx, new_xyz) if return_xyz else idx

    def ball_query(self, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor):
        _check(xyz, "xyz", device=self.device); _check(new_xyz, "new_xyz", device=self.device)
        B, N, stride = xyz.shape
        m = new_xyz.shape[1]
        idx = self._empty(B, m, nsample, dtype=torch.int32)
        _lib.check(self.lib.mpn_ball_query(self._ctx, self.stream, float(radius), nsample, _p(xyz), B, N, stride,
                                           _p(new_xyz), m, _p(idx)))
        return idx

    def gather(self, feat: torch.Tensor, idx: torch.Tensor):
        _check(feat, "features", device=self.device); _check(idx, "idx", torch.int32, self.device)
        B, Cc, N = feat.shape
        m = idx.shape[1]
        out = self._empty(B, Cc, m)
        _lib.check(self.lib.mpn_gather_points(self._ctx, self.stream, _p(feat), B, Cc, N, _p(idx), m, _p(out)))
        return out

    def group(self, feat: torch.Tensor, idx: torch.Tensor):
        _check(feat, "features", device=self.device); _check(idx, "idx", torch.int32, self.device)
        B, Cc, N = feat.shape
        _, m, ns = idx.shape
        out = self._empty(B, Cc, m, ns)
        _lib.check(self.lib.mpn_group_points(self._ctx, self.stream, _p(feat), B, Cc, N, _p(idx), m, ns, _p(out)))
        return out

    def sa_forward(self, module: int, xyz: torch.Tensor, feats: torch.Tensor, debug: bool = False,
                   precision: int = _lib.PREC_FP32):
        """xyz [B,N,3|4], feats point-major [B,N,C] -> (new_xyz [B,m,3] | None, new_feats [B,m,Cout] [, fps_idx, ball_idx])"""
        _check(xyz, "xyz", device=self.device)
        if not (feats.is_cuda and feats.dtype == torch.float32):
            raise RuntimeError("features must be a CUDA float32 tensor")
        B, N, stride = xyz.shape
        npoint, cout = ((512, 64), (128, 256), (1, 1024))[module]
        new_xyz = self._empty(B, npoint, 3) if module < 2 else None
        out = self._empty(B, npoint, cout)
        fidx = self._empty(B, npoint, dtype=torch.int32) if debug and module < 2 else None
        bidx = self._empty(B, npoint, 128, dtype=torch.int32) if debug and module < 2 else None
        fstride = feats.stride(1) if feats.dim() == 3 else feats.shape[2]
        _lib.check(self.lib.mpn_sa_forward(self._ctx, self.stream, module, precision, _p(xyz), stride, _p(feats),
                                           fstride, B, N, _p(new_xyz), _p(out), _p(fidx), _p(bidx)))
        return (new_xyz, out, fidx, bidx) if debug else (new_xyz, out)

    # ------------------------------------------------------------------ robofin
    def fk(self, q: torch.Tensor):
        _check(q, "q", device=self.device)
        B = q.shape[0]
        frames, eef = self._empty(B, 11, 3, 4), self._empty(B, 3, 4)
        _lib.check(self.lib.mpn_fk(self._ctx, self.stream, _p(q), B, _p(frames), _p(eef)))
        return frames, eef

    def sample_robot(self, q: torch.Tensor, n: int, step: int = 0, cloud: Optional[torch.Tensor] = None):
        _check(q, "q", device=self.device)
        B = q.shape[0]
        if cloud is None:
            cloud = torch.zeros((B, n, 4), dtype=torch.float32, device=self.device)
        _check(cloud, "cloud", device=self.device)
        _lib.check(self.lib.mpn_sample_robot(self._ctx, self.stream, _p(q), B, n, step, _p(cloud), cloud.shape[1]))
        return cloud

    def sample_end_effector(self, poses: torch.Tensor, n: int, problem0: int = 0) -> torch.Tensor:
        """poses [B,3,4] (right_gripper frame) -> [B,n,3] gripper points; the keyed subset of build_cloud's target rows"""
        _check(poses, "poses", device=self.device)
        B = poses.shape[0]
        out = self._empty(B, n, 3)
        _lib.check(self.lib.mpn_sample_end_effector(self._ctx, self.stream, _p(poses), B, n, problem0, _p(out)))
        return out

    def compute_spheres(self, q: torch.Tensor):
        _check(q, "q", device=self.device)
        B = q.shape[0]
        out = self._empty(B, self.tables.sphere_centers.shape[0], 3)
        _lib.check(self.lib.mpn_compute_spheres(self._ctx, self.stream, _p(q), B, _p(out)))
        return out

    def normalize(self, q: torch.Tensor):
        _check(q, "q", device=self.device)
        out = torch.empty_like(q)
        _lib.check(self.lib.mpn_normalize_joints(self._ctx, self.stream, _p(q), q.numel() // 7, _p(out)))
        return out

    def unnormalize(self, qn: torch.Tensor):
        _check(qn, "q", device=self.device)
        out = torch.empty_like(qn)
        _lib.check(self.lib.mpn_unnormalize_joints(self._ctx, self.stream, _p(qn), qn.numel() // 7, _p(out)))
        return out

    # ------------------------------------------------------------------ geometry
    def sdf_points(self, scene, points: torch.Tensor, which: int = 0):
        _check(points, "points", device=self.device)
        B, N, _ = points.shape
        s, keep = self._scene(scene, B)
        out = self._empty(B, N)
        _lib.check(self.lib.mpn_sdf_points(self._ctx, self.stream, C.byref(s), B, _p(points), N, which, _p(out)))
        return out

    def build_cloud(self, scene, q0: torch.Tensor, target: torch.Tensor, problem0: int = 0, problem_ids: Optional[torch.Tensor] = None,
                    epoch: int = 0):
        """problem b's sampling streams are keyed by problem0 + b, or by problem_ids[b] (+ epoch * 2^20, wrapping) when given"""
        _check(q0, "q0", device=self.device); _check(target, "target", device=self.device)
        B = q0.shape[0]
        s, keep = self._scene(scene, B)
        cloud = self._empty(B, self.n_points, 4)
        if problem_ids is None:
            _lib.check(self.lib.mpn_build_cloud(self._ctx, self.stream, C.byref(s), B, _p(q0), _p(target), problem0, _p(cloud)))
        else:
            ids = self._ids(problem_ids, epoch)
            _lib.check(self.lib.mpn_build_cloud_ids(self._ctx, self.stream, C.byref(s), B, _p(q0), _p(target), _p(ids), _p(cloud)))
        return cloud

    def _ids(self, ids: torch.Tensor, epoch: int = 0) -> torch.Tensor:
        """int64 / int32 indices -> the u32 counters (stored in an int32 tensor) of the per-sample RNG streams"""
        if not ids.is_cuda:
            raise RuntimeError("ids must be a CUDA tensor")
        v = (ids.to(torch.int64) + (int(epoch) << 20)) & 0xFFFFFFFF
        return torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32).contiguous()

    def augment_joints(self, q: torch.Tensor, random_scale: float, sample0: Optional[torch.Tensor] = None, epoch: int = 0):
        """data_loader.py:167-180 on the device: (clamp(q + random_scale * N(0,1), limits), its normalisation), both [B,7]"""
        _check(q, "q", device=self.device)
        B = q.shape[0]
        ids = None if sample0 is None else self._ids(sample0)
        out, outn = torch.empty_like(q), torch.empty_like(q)
        _lib.check(self.lib.mpn_augment_joints(self._ctx, self.stream, _p(q), B, float(random_scale), _p(ids), int(epoch) & 0xFFFFFFFF,
                                               _p(out), _p(outn)))
        return out, outn

    def clean_point_cloud(self, xyz: torch.Tensor, rgba: Optional[torch.Tensor] = None, n_out: int = 4096, cloud_id: int = 0):
        """planning_node.py:187-228: workspace crop + random subset without replacement -> (xyz [n_out,3], rgba [n_out,4] | None).
        Raises ValueError when fewer than n_out points lie inside the workspace (np.random.choice does the same)."""
        _check(xyz, "xyz", device=self.device)
        if rgba is not None:
            _check(rgba, "rgba", device=self.device)
        N = xyz.shape[0]
        out = self._empty(n_out, 3)
        out_c = self._empty(n_out, 4) if rgba is not None else None
        kept = self._empty(1, dtype=torch.int32)
        scratch = self._empty(N, dtype=torch.int32)
        _lib.check(self.lib.mpn_clean_point_cloud(self._ctx, self.stream, _p(xyz), _p(rgba), N, n_out, cloud_id, _p(out), _p(out_c),
                                                  _p(kept), _p(scratch)))
        k = int(kept.item())
        if k < n_out:
            raise ValueError(f"Cannot take a larger sample than population when 'replace=False' ({k} points in the workspace, {n_out} wanted)")
        return out, out_c

    def build_cloud_from_points(self, q0: torch.Tensor, target: torch.Tensor, obstacle_points: torch.Tensor,
                                obstacle_counts: torch.Tensor, problem0: int = 0):
        """run_inference.make_point_cloud_from_problem (run_inference.py:58-90): obstacle_points [B,P,3], counts i32 [B]"""
        _check(q0, "q0", device=self.device); _check(target, "target", device=self.device)
        _check(obstacle_points, "obstacle_points", device=self.device)
        _check(obstacle_counts, "obstacle_counts", torch.int32, self.device)
        B = q0.shape[0]
        if obstacle_points.shape[0] != B or obstacle_points.shape[2] != 3:
            raise RuntimeError(f"obstacle_points must be [B, P, 3], got {tuple(obstacle_points.shape)}")
        cloud = self._empty(B, self.n_points, 4)
        _lib.check(self.lib.mpn_build_cloud_from_points(self._ctx, self.stream, B, _p(q0), _p(target), _p(obstacle_points),
                                                        _p(obstacle_counts), obstacle_points.shape[1], problem0, _p(cloud)))
        return cloud

    def render_depth_cloud(self, scene, camera: torch.Tensor, width: int = 640, height: int = 480, fov_y_deg: float = 60.0,
                           near: float = 0.01, far: float = 10.0):
        """Partial-view obstacle clouds (run_inference.convert_primitive_problems_to_depth, run_inference.py:194-257):
        camera [3,4] or [B,3,4] camera->world -> (points [B, W*H, 3] with the hits first, counts i32 [B])."""
        import math
        _check(camera, "camera", device=self.device)
        B = scene["cuboid_centers"].shape[0]
        s, keep = self._scene(scene, B)
        per = camera.dim() == 3
        if per and camera.shape[0] != B:
            raise RuntimeError(f"camera has batch {camera.shape[0]}, expected {B}")
        ty = math.tan(math.radians(fov_y_deg) / 2.0)
        pts = self._empty(B, width * height, 3)
        cnt = self._empty(B, dtype=torch.int32)
        _lib.check(self.lib.mpn_render_depth_cloud(self._ctx, self.stream, C.byref(s), B, _p(camera), int(per), width, height,
                                                   ty * width / height, ty, near, far, _p(pts), _p(cnt)))
        return pts, cnt

    def sweep_flags(self, scene, traj: torch.Tensor):
        _check(traj, "traj", device=self.device)
        B, T, _ = traj.shape
        s, keep = self._scene(scene, B)
        flags = self._empty(B, dtype=torch.uint8)
        first = self._empty(B, dtype=torch.int32)
        _lib.check(self.lib.mpn_sweep_flags(self._ctx, self.stream, C.byref(s), B, _p(traj), T, 0, 0, _p(flags), _p(first)))
        return flags, first

    def _volumes(self, vol: Optional[Dict[str, torch.Tensor]], B: int):
        """optional primitive lists of the region test -> (MpnScene or None, n_cuboids, n_cylinders, keep-alive)"""
        if vol is None:
            return None, 0, 0, []
        s = _lib.MpnScene()
        keep, n1, n2 = [], 0, 0
        for k in SCENE_KEYS:
            if k not in vol or vol[k] is None:
                continue
            t = _check(vol[k], k, device=self.device)
            if t.shape[0] != B:
                raise RuntimeError(f"{k} has batch {t.shape[0]}, expected {B}")
            keep.append(t)
            setattr(s, k, t.data_ptr())
        if "cuboid_centers" in vol and vol["cuboid_centers"] is not None:
            n1 = vol["cuboid_centers"].shape[1]
        if "cylinder_centers" in vol and vol["cylinder_centers"] is not None:
            n2 = vol["cylinder_centers"].shape[1]
        return s, n1, n2, keep

    def evaluate(self, scene, traj: torch.Tensor, target: torch.Tensor, num_poses: Optional[torch.Tensor] = None,
                 target_volume: Optional[Dict[str, torch.Tensor]] = None,
                 negative_volumes: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        """Evaluator.evaluate_trajectory for B trajectories (metrics.py:447-523): traj [B,T+1,7] -> eval [B,16]
        (columns: mpinets_b200._lib.EVAL_COLUMNS)."""
        _check(traj, "traj", device=self.device); _check(target, "target", device=self.device)
        B, T1, _ = traj.shape
        s, keep = self._scene(scene, B)
        if num_poses is not None:
            _check(num_poses, "num_poses", torch.int32, self.device)
        tv, v1, v2, k1 = self._volumes(target_volume, B)
        nv, n1, n2, k2 = self._volumes(negative_volumes, B)
        out = self._empty(B, _lib.EVAL_COLS)
        _lib.check(self.lib.mpn_evaluate(self._ctx, self.stream, C.byref(s), B, _p(traj), T1,
                                         _p(num_poses) if num_poses is not None else None, _p(target),
                                         C.byref(tv) if tv is not None else None, v1, v2,
                                         C.byref(nv) if nv is not None else None, n1, n2, _p(out)))
        return out

    def sparc(self, movement: torch.Tensor, fs: float, num_samples: Optional[torch.Tensor] = None, padlevel: int = 4, fc: float = 10.0,
              amp_th: float = 0.05) -> torch.Tensor:
        """third_party/sparc.py for B speed profiles [B, n] -> spectral arc length [B]"""
        _check(movement, "movement", device=self.device)
        B, n = movement.shape
        if num_samples is not None:
            _check(num_samples, "num_samples", torch.int32, self.device)
        out = self._empty(B)
        _lib.check(self.lib.mpn_sparc(self._ctx, self.stream, B, n, _p(movement), _p(num_samples), fs, padlevel, fc, amp_th, _p(out)))
        return out

    # ------------------------------------------------------------------ losses (loss.py)
    def collision_loss(self, scene, points: torch.Tensor, margin: float = 0.03, need_grad: bool = False):
        """loss.collision_loss (loss.py:47-94): points [B,N,3] -> (loss [1], grad_points [B,N,3] or None)"""
        _check(points, "input_pc", device=self.device)
        B, N, _ = points.shape
        s, keep = self._scene(scene, B)
        loss = self._empty(1)
        grad = self._empty(B, N, 3) if need_grad else None
        _lib.check(self.lib.mpn_collision_loss(self._ctx, self.stream, C.byref(s), B, N, _p(points), margin, _p(loss), _p(grad)))
        return loss, grad

    def point_match_loss(self, a: torch.Tensor, b: torch.Tensor, need_grad: bool = False):
        """loss.point_match_loss (loss.py:31-44) -> (loss [1], grad_a or None)"""
        _check(a, "input_pc", device=self.device); _check(b, "target_pc", device=self.device)
        if a.shape != b.shape:
            raise RuntimeError(f"point clouds differ in shape: {tuple(a.shape)} vs {tuple(b.shape)}")
        loss = self._empty(1)
        grad = torch.empty_like(a) if need_grad else None
        _lib.check(self.lib.mpn_point_match_loss(self._ctx, self.stream, a.numel(), _p(a), _p(b), _p(loss), _p(grad)))
        return loss, grad

    def bc_collision_losses(self, scene, input_normalized: torch.Tensor, target_normalized: torch.Tensor, n_points: int = 1024,
                            margin: float = 0.03, w_collision: float = 1.0, w_bc: float = 1.0, need_grad: bool = False):
        """CollisionAndBCLossContainer.__call__ (loss.py:111-166) -> (losses [2] = (collision, point match), grad_input)"""
        _check(input_normalized, "input_normalized", device=self.device)
        _check(target_normalized, "target_normalized", device=self.device)
        B = input_normalized.shape[0]
        s, keep = self._scene(scene, B)
        losses = self._empty(2)
        grad = self._empty(B, 7) if need_grad else None
        _lib.check(self.lib.mpn_bc_collision_losses(self._ctx, self.stream, C.byref(s), B, _p(input_normalized), _p(target_normalized),
                                                    n_points, margin, w_collision, w_bc, _p(losses), _p(grad)))
        return losses, grad

    # ------------------------------------------------------------------ model
    def encoder_forward(self, cloud: torch.Tensor, precision: int = _lib.PREC_FP32):
        _check(cloud, "point_cloud", device=self.device)
        assert cloud.size(2) == 4  # model.py:420
        B, N, _ = cloud.shape
        out = self._empty(B, 2048)
        _lib.check(self.lib.mpn_encoder_forward(self._ctx, self.stream, precision, _p(cloud), B, N, _p(out)))
        return out

    def policy_forward(self, cloud: torch.Tensor, q_norm: torch.Tensor, precision: int = _lib.PREC_FP32):
        _check(cloud, "xyz", device=self.device); _check(q_norm, "q", device=self.device)
        assert cloud.size(2) == 4
        B, N, _ = cloud.shape
        dq = self._empty(B, 7)
        _lib.check(self.lib.mpn_policy_forward(self._ctx, self.stream, precision, _p(cloud), _p(q_norm), B, N, _p(dq)))
        return dq

    def rollout(self, scene, cloud: torch.Tensor, q0: torch.Tensor, target: torch.Tensor, steps: int,
                early_exit: bool = False, check_every_step: bool = False, precision: int = _lib.PREC_FP32,
                traj: Optional[torch.Tensor] = None, metrics: Optional[torch.Tensor] = None):
        """Lock-step rollout of `steps` policy steps; `cloud` is updated in place (model.py:181).
        Returns (traj [B,steps+1,7] unnormalised, metrics [B,8])."""
        _check(cloud, "xyz", device=self.device); _check(q0, "q0", device=self.device); _check(target, "target", device=self.device)
        B, N, _ = cloud.shape
        s, keep = self._scene(scene, B)
        if traj is None:
            traj = self._empty(B, steps + 1, 7)
        if metrics is None:
            metrics = self._empty(B, _lib.METRICS_COLS)
        _lib.check(self.lib.mpn_rollout(self._ctx, self.stream, precision, C.byref(s), B, N, _p(cloud), _p(q0), _p(target),
                                        steps, int(early_exit), int(check_every_step), _p(traj), _p(metrics)))
        return traj, metrics

    # ------------------------------------------------------------------ training (model.py:185-240, 68-73)
    @property
    def param_count(self) -> int:
        return int(self.lib.mpn_param_count(self._ctx))

    def param_layout(self):
        """[(state-dict key, offset, numel)] of the flat parameter / gradient vector"""
        out, i = [], 0
        name = C.create_string_buffer(160)
        off, n = C.c_int64(0), C.c_int64(0)
        while self.lib.mpn_param_info(self._ctx, i, name, 160, C.byref(off), C.byref(n)) == 0:
            out.append((name.value.decode(), int(off.value), int(n.value)))
            i += 1
        if not out:
            raise _lib.MpnError("no parameters: load a state dict first")
        return out

    def get_params(self) -> torch.Tensor:
        p = self._empty(self.param_count)
        _lib.check(self.lib.mpn_get_params(self._ctx, self.stream, _p(p)))
        return p

    def set_params(self, flat: torch.Tensor):
        _check(flat, "params", device=self.device)
        if flat.numel() != self.param_count:
            raise RuntimeError(f"params has {flat.numel()} elements, expected {self.param_count}")
        _lib.check(self.lib.mpn_set_params(self._ctx, self.stream, _p(flat)))

    def weights_sync(self):
        """rebuild the packed bf16 tensor-core weights from the fp32 parameters (after optimisation steps)"""
        _lib.check(self.lib.mpn_weights_sync(self._ctx))

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """The current parameters under the reference's state-dict keys (Conv2d weights as [Cout, Cin, 1, 1])."""
        flat = self.get_params()
        sd = {}
        lay = self.param_layout()
        sizes = {k: m for (k, o, m) in lay}
        for name, off, n in lay:
            t = flat[off:off + n].clone()
            if name.endswith(".weight") and ".fc_layer.1." not in name and ".fc_layer.4." not in name:
                out_f = sizes[name[:-len("weight")] + "bias"]
                t = t.view(out_f, n // out_f)
                if ".SA_modules." in name:
                    t = t.view(out_f, n // out_f, 1, 1)
            sd[name] = t
        return sd

    def unflatten(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        """views of a flat parameter / gradient vector by state-dict key (2-D weights)"""
        lay = self.param_layout()
        sizes = {k: m for (k, o, m) in lay}
        out = {}
        for name, off, n in lay:
            t = flat[off:off + n]
            if name.endswith(".weight") and ".fc_layer.1." not in name and ".fc_layer.4." not in name:
                t = t.view(sizes[name[:-len("weight")] + "bias"], -1)
            out[name] = t
        return out

    def train_step_grads(self, scene, cloud: torch.Tensor, q_norm: torch.Tensor, supervision: torch.Tensor,
                         n_loss_points: int = 1024, margin: float = 0.03, w_collision: float = 5.0, w_bc: float = 1.0,
                         grads: Optional[torch.Tensor] = None, need_grad: bool = True, precision: int = _lib.PREC_FP32):
        """training_step (model.py:185-240) up to the gradients: returns (losses [2] = (collision, point match),
        y_hat [B,7], grads [param_count] or None).  precision: PREC_FP32 (parity mode) or PREC_BF16 (SA1 / SA2 backward GEMMs
        on tcgen05, bf16 operands)."""
        _check(cloud, "xyz", device=self.device); _check(q_norm, "configuration", device=self.device)
        _check(supervision, "supervision", device=self.device)
        assert cloud.size(2) == 4
        B, N, _ = cloud.shape
        s, keep = self._scene(scene, B)
        losses, y_hat = self._empty(2), self._empty(B, 7)
        if need_grad and grads is None:
            grads = self._empty(self.param_count)
        if grads is not None:
            _check(grads, "grads", device=self.device)
        _lib.check(self.lib.mpn_train_step_grads(self._ctx, self.stream, C.byref(s), B, N, _p(cloud), _p(q_norm), _p(supervision),
                                                 n_loss_points, margin, w_collision, w_bc, _p(losses), _p(y_hat),
                                                 _p(grads) if need_grad else None, precision))
        return losses, y_hat, (grads if need_grad else None)

    def train_tc_gemm(self, A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor] = None, epi: int = 0,
                      mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """C [M,N] bf16 = epi(A [M,128] @ W [N,128]^T + bias) on tcgen05 (train_tc.cu); epi 0 relu, 1 none, 2 relu'(mask)"""
        _check(A, "A", torch.bfloat16, self.device); _check(W, "W", torch.bfloat16, self.device)
        M, N = A.shape[0], W.shape[0]
        out = self._empty(M, N, dtype=torch.bfloat16)
        _lib.check(self.lib.mpn_train_tc_gemm(self._ctx, self.stream, epi, _p(A), _p(W), _p(bias), _p(mask), M, N, _p(out)))
        return out

    def train_tc_wgrad(self, dY: torch.Tensor, X: torch.Tensor, variant: int = 0) -> torch.Tensor:
        """dY^T X for [R,128] bf16 operands on tcgen05 (MN-major operands) -> [128,128] fp32"""
        _check(dY, "dY", torch.bfloat16, self.device); _check(X, "X", torch.bfloat16, self.device)
        part = self._empty(512, 128, 128)
        n = C.c_int(0)
        _lib.check(self.lib.mpn_train_tc_wgrad(self._ctx, self.stream, _p(dY), _p(X), dY.shape[0], _p(part), part.numel(), C.byref(n), variant))
        return part[: n.value].sum(dim=0)

    def train_pooled_rows(self, B: int):
        """max-pool routing of the last training step: (u8 [B,512,64], u8 [B,128,256], u8 [B,1024])"""
        outs = []
        for m, shape in enumerate(((B, 512, 64), (B, 128, 256), (B, 1024))):
            t = self._empty(*shape, dtype=torch.uint8)
            _lib.check(self.lib.mpn_train_pooled_rows(self._ctx, self.stream, m, B, _p(t)))
            outs.append(t)
        return tuple(outs)

    def adam_step(self, grads: torch.Tensor, step: int, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                  clip_norm: float = 1.0) -> torch.Tensor:
        """clip_grad_norm_(clip_norm) + torch.optim.Adam on the flat vector; returns the pre-clip gradient norm [1]"""
        _check(grads, "grads", device=self.device)
        norm = self._empty(1)
        _lib.check(self.lib.mpn_adam_step(self._ctx, self.stream, _p(grads), lr, betas[0], betas[1], eps, clip_norm, int(step), _p(norm)))
        return norm
