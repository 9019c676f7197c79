"""``mpinets.model`` surface (``/root/reference/mpinets/model.py``): PyTorch holds only the weights (same module tree and
state-dict keys as the reference, so a Lightning checkpoint's ``state_dict`` loads unchanged); ``forward`` / ``rollout``
run in ``libmpinets_b200.so``."""
from __future__ import annotations

import weakref
from typing import Callable, Dict, List, Optional

import numpy as np
import torch
from torch import nn

from . import _lib
from .pointnet2_modules import PointnetSAModule
from .runtime import get_engine

END_EFFECTOR_FRAME = "right_gripper"   # run_inference.py:51-55
NUM_ROBOT_POINTS, NUM_OBSTACLE_POINTS, NUM_TARGET_POINTS, MAX_ROLLOUT_LENGTH = 2048, 4096, 128, 150
PRECISIONS = _lib.PRECISIONS   # "fp32" (SIMT parity mode), "bf16x3" (tensor cores, parity grade), "bf16" (tensor cores, throughput)


class MPiNetsPointNet(nn.Module):
    """model.py:355-426"""

    def __init__(self):
        super().__init__()
        self.SA_modules = nn.ModuleList([
            PointnetSAModule(npoint=512, radius=0.05, nsample=128, mlp=[1, 64, 64, 64], bn=False),
            PointnetSAModule(npoint=128, radius=0.3, nsample=128, mlp=[64, 128, 128, 256], bn=False),
            PointnetSAModule(mlp=[256, 512, 512, 1024], bn=False),
        ])
        self.fc_layer = nn.Sequential(nn.Linear(1024, 4096), nn.GroupNorm(16, 4096), nn.LeakyReLU(inplace=True),
                                      nn.Linear(4096, 2048), nn.GroupNorm(16, 2048), nn.LeakyReLU(inplace=True),
                                      nn.Linear(2048, 2048))


class MotionPolicyNetwork(nn.Module):
    """model.py:35-91.  ``precision``: "bf16x3" (default: split-bf16 operands on tcgen05, delta-q within 1e-5 of the reference's fp32
    forward), "bf16" (tcgen05 throughput mode, ~2e-4) or "fp32" (SIMT FMA parity mode)."""

    def __init__(self, precision: str = "bf16x3"):
        super().__init__()
        self.point_cloud_encoder = MPiNetsPointNet()
        self.feature_encoder = nn.Sequential(nn.Linear(7, 32), nn.LeakyReLU(), nn.Linear(32, 64), nn.LeakyReLU(),
                                             nn.Linear(64, 128), nn.LeakyReLU(), nn.Linear(128, 128), nn.LeakyReLU(),
                                             nn.Linear(128, 64))
        self.decoder = nn.Sequential(nn.Linear(2048 + 64, 512), nn.LeakyReLU(), nn.Linear(512, 256), nn.LeakyReLU(),
                                     nn.Linear(256, 128), nn.LeakyReLU(), nn.Linear(128, 7))
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
        self.precision = precision
        self._synced_device = None
        self._dirty = False

    # -- weights -> engine
    def sync_engine(self, device: Optional[torch.device] = None):
        """Loads this module's weights into the device's engine context.  The context holds ONE parameter set: if another module
        owns it and has optimiser updates that exist only in the context, those are pulled back into that module first."""
        eng = get_engine(device)
        prev = eng._weights_owner() if eng._weights_owner is not None else None
        if prev is not None and prev is not self and getattr(prev, "_dirty", False):
            prev.pull_weights()
        eng.load_state_dict(self.state_dict())
        eng._weights_owner = weakref.ref(self)
        self._synced_device = eng.device
        return eng

    def _engine(self, like: torch.Tensor):
        eng = get_engine(like.device)
        owner = eng._weights_owner() if eng._weights_owner is not None else None
        if self._synced_device != like.device or owner is not self:   # never run on another module's weights
            return self.sync_engine(like.device)
        return eng

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._synced_device = None
        return r

    def forward(self, xyz: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
        """xyz [B,N,4], q [B,7] normalised -> delta q [B,7] (normalised space)"""
        return self._engine(xyz).policy_forward(xyz.contiguous(), q.contiguous(), PRECISIONS[self.precision])

    def encode(self, xyz: torch.Tensor) -> torch.Tensor:
        return self._engine(xyz).encoder_forward(xyz.contiguous(), PRECISIONS[self.precision])

    def rollout(self, batch: Dict[str, torch.Tensor], rollout_length: int, sampler: Optional[Callable] = None,
                unnormalize: bool = False, scene: Optional[Dict[str, torch.Tensor]] = None) -> List[torch.Tensor]:
        """TrainingMotionPolicyNetwork.rollout (model.py:128-183): lock-step, whole loop in one library call;
        ``batch["xyz"]`` is updated in place.  ``sampler`` is accepted for signature compatibility (the engine resamples
        the robot surface itself)."""
        xyz, q = batch["xyz"], batch["configuration"]
        if q.ndim == 1:
            xyz, q = xyz.unsqueeze(0), q.unsqueeze(0)
        eng = self._engine(xyz)
        B = q.shape[0]
        if scene is None:
            keys = ("cuboid_centers", "cuboid_dims", "cuboid_quats", "cylinder_centers", "cylinder_radii", "cylinder_heights", "cylinder_quats")
            scene = {k: batch[k].contiguous().float() for k in keys}
        q0 = eng.unnormalize(q.contiguous())
        target = batch.get("target_pose")
        if target is None:
            target = torch.eye(4, device=xyz.device)[:3].expand(B, 3, 4).contiguous()
        traj, metrics = eng.rollout(scene, xyz, q0, target.contiguous(), rollout_length, precision=PRECISIONS[self.precision])
        self.last_metrics = metrics
        if unnormalize:
            return [traj[:, t] for t in range(rollout_length + 1)]
        return [eng.normalize(traj[:, t].contiguous()) for t in range(rollout_length + 1)]


def rollout_until_success(mdl: MotionPolicyNetwork, q0: np.ndarray, target_pose: np.ndarray, point_cloud: torch.Tensor,
                          scene: Dict[str, torch.Tensor], max_steps: int = MAX_ROLLOUT_LENGTH) -> np.ndarray:
    """run_inference.rollout_until_success (run_inference.py:137-191) for B = 1 .. many problems in lock-step with a done
    mask: returns the trajectories [B, T'+1, 7] truncated at each problem's stopping step (list when B > 1)."""
    eng = mdl._engine(point_cloud)
    q = torch.as_tensor(np.atleast_2d(q0), dtype=torch.float32, device=point_cloud.device)
    tgt = torch.as_tensor(np.asarray(target_pose, dtype=np.float32).reshape(-1, 4, 4)[:, :3], device=point_cloud.device).contiguous()
    traj, metrics = eng.rollout(scene, point_cloud, q, tgt, max_steps, early_exit=True, precision=PRECISIONS[mdl.precision])
    steps = metrics[:, _lib_steps_col()].long().cpu().numpy()
    out = [traj[b, : steps[b] + 1].cpu().numpy() for b in range(q.shape[0])]
    return out[0] if len(out) == 1 else out


def _lib_steps_col() -> int:
    return 2   # MPN_M_STEPS


class FlatAdam:
    """``configure_optimizers`` (model.py:68-73: ``torch.optim.Adam(self.parameters(), lr=1e-4)``) plus the Trainer's
    ``gradient_clip_val=1.0`` (run_training.py:112) and the DDP gradient averaging (run_training.py:71-77), on the engine's
    flat parameter vector: ``step()`` = all-reduce-mean (when torch.distributed is initialised) -> clip -> Adam."""

    def __init__(self, module: "TrainingMotionPolicyNetwork", lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 clip_norm: float = 1.0):
        self.module, self.lr, self.betas, self.eps, self.clip_norm = module, lr, betas, eps, clip_norm
        self.steps = 0
        self.last_grad_norm = None
        self._weights_version = None

    def zero_grad(self, set_to_none: bool = True):
        pass   # the training step overwrites the gradient vector

    def step(self):
        from .parallel import allreduce_mean_
        m = self.module
        if m.grads is None:
            raise RuntimeError("FlatAdam.step() before training_step()")
        allreduce_mean_(m.grads)
        eng = m._engine(m.grads)
        if self._weights_version != eng.weights_version:   # the context's parameters were (re)loaded: its Adam moments start from zero
            if self._weights_version is None:
                from .parallel import broadcast_params_
                broadcast_params_(eng)                      # DDP wraps the module with rank 0's parameters (run_training.py:71-77)
            self._weights_version = eng.weights_version
            self.steps = 0
        self.steps += 1
        self.last_grad_norm = eng.adam_step(m.grads, self.steps, self.lr, self.betas, self.eps, self.clip_norm)
        m._dirty = True


class TrainingMotionPolicyNetwork(MotionPolicyNetwork):
    """model.py:94-352, the training side: ``training_step`` runs forward + losses + backward in the library (fp32) and
    leaves the gradient of ``point_match_loss_weight * point_match + collision_loss_weight * collision`` in ``self.grads``
    (flat, ``Engine.unflatten`` gives per-key views); ``configure_optimizers().step()`` applies it.  The nn.Module copies of
    the weights are refreshed lazily (``pull_weights``)."""

    def __init__(self, num_robot_points: int = NUM_ROBOT_POINTS, point_match_loss_weight: float = 1.0,
                 collision_loss_weight: float = 5.0, precision: str = "fp32"):
        super().__init__(precision)
        self.num_robot_points = num_robot_points
        self.point_match_loss_weight = point_match_loss_weight
        self.collision_loss_weight = collision_loss_weight
        self.grads: Optional[torch.Tensor] = None
        self.logged: Dict[str, torch.Tensor] = {}
        self._dirty = False

    def configure_optimizers(self) -> FlatAdam:
        return FlatAdam(self, lr=1e-4)

    def log(self, name: str, value: torch.Tensor):
        self.logged[name] = value

    def training_step(self, batch: Dict[str, torch.Tensor], batch_idx: int = 0) -> torch.Tensor:
        """batch keys of data_loader.py:153-280: xyz [B,N,4], configuration [B,7], supervision [B,7], the seven primitive
        tensors.  Returns the weighted loss (model.py:235-238)."""
        xyz, q = batch["xyz"].contiguous(), batch["configuration"].contiguous()
        eng = self._engine(xyz)
        keys = ("cuboid_centers", "cuboid_dims", "cuboid_quats", "cylinder_centers", "cylinder_radii", "cylinder_heights", "cylinder_quats")
        scene = {k: batch[k].contiguous().float() for k in keys}
        losses, y_hat, self.grads = eng.train_step_grads(scene, xyz, q, batch["supervision"].contiguous(),
                                                         w_collision=self.collision_loss_weight, w_bc=self.point_match_loss_weight,
                                                         grads=self.grads, precision=PRECISIONS[self.precision])
        self.log("point_match_loss", losses[1])
        self.log("collision_loss", losses[0])
        val_loss = self.point_match_loss_weight * losses[1] + self.collision_loss_weight * losses[0]
        self.log("val_loss", val_loss)
        return val_loss

    def pull_weights(self):
        """copy the engine's (optimised) parameters back into this module and refresh the bf16 tensor-core copies"""
        dev = self._synced_device
        if dev is None:
            return
        eng = get_engine(dev)
        owner = eng._weights_owner() if eng._weights_owner is not None else None
        if owner is not self:
            raise RuntimeError("the engine context no longer holds this module's parameters")
        nn.Module.load_state_dict(self, {k: v for k, v in eng.state_dict().items()}, strict=False)
        eng.weights_sync()
        self._dirty = False

    # the optimiser updates the engine's flat parameter vector; the nn.Module copies are refreshed whenever they are read
    def state_dict(self, *args, **kwargs):
        if self._dirty:
            self.pull_weights()
        return super().state_dict(*args, **kwargs)

    def named_parameters(self, *args, **kwargs):
        if self._dirty:
            self.pull_weights()
        return super().named_parameters(*args, **kwargs)

    # -- validation (model.py:252-352)
    VALIDATION_ROLLOUT_LENGTH = 69   # model.py:272

    def validation_step(self, batch: Dict[str, torch.Tensor], batch_idx: int = 0, rollout_length: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """model.py:252-318 in one library call: 69-step lock-step rollout, final end-effector error against
        batch["target_position"] [B,3], link-sphere SDF sweep over the 70 configurations -> avg_target_error,
        avg_collision_rate.  batch["xyz"] is updated in place like the reference's rollout."""
        T = self.VALIDATION_ROLLOUT_LENGTH if rollout_length is None else rollout_length
        xyz, q = batch["xyz"], batch["configuration"].contiguous()
        eng = self._engine(xyz)
        B = q.shape[0]
        keys = ("cuboid_centers", "cuboid_dims", "cuboid_quats", "cylinder_centers", "cylinder_radii", "cylinder_heights", "cylinder_quats")
        scene = {k: batch[k].contiguous().float() for k in keys}
        target = torch.eye(4, device=xyz.device)[:3].repeat(B, 1, 1)
        target[:, :, 3] = batch["target_position"]
        traj, metrics = eng.rollout(scene, xyz, eng.unnormalize(q), target.contiguous(), T, precision=PRECISIONS[self.precision])
        position_error = metrics[:, 3]       # MPN_M_POS_ERR: ||eff(rollout[-1]) - target_position||
        has_collision = metrics[:, 0] > 0    # MPN_M_COLLISION: any sphere, any of the T + 1 configurations, sdf <= radius
        self.last_rollout = traj
        return {"avg_target_error": position_error.mean(), "avg_collision_rate": torch.count_nonzero(has_collision) / B}

    def validation_step_end(self, batch_parts: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return {"avg_target_error": torch.mean(batch_parts["avg_target_error"]),
                "avg_collision_rate": torch.mean(batch_parts["avg_collision_rate"])}

    def validation_epoch_end(self, validation_step_outputs):
        self.log("avg_target_error", torch.mean(torch.stack([x["avg_target_error"] for x in validation_step_outputs])))
        self.log("avg_collision_rate", torch.mean(torch.stack([x["avg_collision_rate"] for x in validation_step_outputs])))
