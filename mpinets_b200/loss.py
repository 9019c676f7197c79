"""mpinets/loss.py over the C ABI: same names, argument order and reductions, differentiable through torch.autograd
(each Function's backward is the analytic gradient the CUDA kernel produced in the same launch as the value).

    point_match_loss(input_pc, target_pc)                      loss.py:31-44
    collision_loss(input_pc, cuboid_centers, ..., cylinder_quaternions)   loss.py:47-94
    CollisionAndBCLossContainer()(input_normalized, <7 scene tensors>, target_normalized)   loss.py:97-166
"""
from __future__ import annotations

from typing import Tuple

import torch

from .runtime import get_engine


def _scene(cuboid_centers, cuboid_dims, cuboid_quaternions, cylinder_centers, cylinder_radii, cylinder_heights, cylinder_quaternions):
    c = lambda t: t.detach().to(torch.float32).contiguous()  # noqa: E731
    return dict(cuboid_centers=c(cuboid_centers), cuboid_dims=c(cuboid_dims), cuboid_quats=c(cuboid_quaternions),
                cylinder_centers=c(cylinder_centers), cylinder_radii=c(cylinder_radii), cylinder_heights=c(cylinder_heights),
                cylinder_quats=c(cylinder_quaternions))


class _PointMatch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input_pc, target_pc):
        eng = get_engine(input_pc.device)
        loss, grad = eng.point_match_loss(input_pc.detach().contiguous(), target_pc.detach().contiguous(), need_grad=True)
        ctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return g * grad, -g * grad   # d/d target = -d/d input for both the mse and the l1 term


def point_match_loss(input_pc: torch.Tensor, target_pc: torch.Tensor) -> torch.Tensor:
    """mse(mean) + l1(mean) between two [B, N, 3] clouds (loss.py:31-44)"""
    return _PointMatch.apply(input_pc, target_pc)


class _Collision(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input_pc, scene):
        eng = get_engine(input_pc.device)
        loss, grad = eng.collision_loss(scene, input_pc.detach().contiguous(), margin=0.03, need_grad=True)
        ctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return g * grad, None


def collision_loss(input_pc, cuboid_centers, cuboid_dims, cuboid_quaternions, cylinder_centers, cylinder_radii, cylinder_heights,
                   cylinder_quaternions) -> torch.Tensor:
    """hinge loss on the scene sdf with a 3 cm margin (loss.py:47-94); zero-volume primitives are ignored"""
    return _Collision.apply(input_pc, _scene(cuboid_centers, cuboid_dims, cuboid_quaternions, cylinder_centers, cylinder_radii,
                                             cylinder_heights, cylinder_quaternions))


class _BCAndCollision(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input_normalized, target_normalized, scene, num_points):
        eng = get_engine(input_normalized.device)
        # the two per-loss gradients are needed separately (the caller weights the losses after the fact, model.py:232-236)
        l_c, g_c = eng.bc_collision_losses(scene, input_normalized.detach().contiguous(), target_normalized.detach().contiguous(),
                                           n_points=num_points, w_collision=1.0, w_bc=0.0, need_grad=True)
        l_p, g_p = eng.bc_collision_losses(scene, input_normalized.detach().contiguous(), target_normalized.detach().contiguous(),
                                           n_points=num_points, w_collision=0.0, w_bc=1.0, need_grad=True)
        ctx.save_for_backward(g_c, g_p)
        return l_c[0], l_p[1]

    @staticmethod
    def backward(ctx, g_collision, g_point_match):
        g_c, g_p = ctx.saved_tensors
        return g_collision * g_c + g_point_match * g_p, None, None, None


class CollisionAndBCLossContainer:
    """loss.py:97-166.  The reference caches a FrankaSampler with a fixed 1024-point robot cloud; here the fixed subset is the
    engine's seeded permutation of the non-base link points (include/mpinets_b200.h: mpn_bc_collision_losses)."""

    def __init__(self):
        self.fk_sampler = None      # kept for attribute parity; the engine owns the tables
        self.num_points = 1024

    def __call__(self, input_normalized, cuboid_centers, cuboid_dims, cuboid_quaternions, cylinder_centers, cylinder_radii,
                 cylinder_heights, cylinder_quaternions, target_normalized) -> Tuple[torch.Tensor, torch.Tensor]:
        scene = _scene(cuboid_centers, cuboid_dims, cuboid_quaternions, cylinder_centers, cylinder_radii, cylinder_heights,
                       cylinder_quaternions)
        return _BCAndCollision.apply(input_normalized, target_normalized, scene, self.num_points)
