"""``pointnet2_ops.pointnet2_utils`` names (fishbotics/pointnet2_ops v3.2.0, ``/root/reference/docker/Dockerfile:152``)
over the C ABI.  Contracts: CUDA, contiguous, float32 / int32 tensors; index outputs are int32 and non-differentiable."""
from __future__ import annotations

import torch
from torch import nn

from .runtime import get_engine


def furthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """xyz [B,N,3] -> idx int32 [B,npoint]"""
    return get_engine(xyz.device).fps(xyz, npoint)


def gather_operation(features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """features [B,C,N], idx [B,npoint] -> [B,C,npoint]"""
    return get_engine(features.device).gather(features, idx)


def ball_query(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    """-> idx int32 [B,npoint,nsample]"""
    return get_engine(xyz.device).ball_query(radius, nsample, xyz, new_xyz)


def grouping_operation(features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """features [B,C,N], idx [B,npoint,nsample] -> [B,C,npoint,nsample]"""
    return get_engine(features.device).group(features, idx)


class QueryAndGroup(nn.Module):
    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx) - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            return grouped_xyz
        grouped = grouping_operation(features, idx)
        return torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped


class GroupAll(nn.Module):
    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz=None, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped = features.unsqueeze(2)
        return torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped
