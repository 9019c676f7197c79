"""``pointnet2_ops.pointnet2_modules.PointnetSAModule`` as used at ``/root/reference/mpinets/model.py:365-383``:
holds the shared-MLP weights under the reference's parameter names (``mlps.0.{0,2,4}.{weight,bias}``, Conv2d 1x1 with
bias, ``bn=False``); ``forward`` runs the fused CUDA set-abstraction kernel (FPS -> ball query -> group -> MLP -> max)."""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from . import _lib
from .runtime import get_engine

_MODULE_OF = {(512, 0.05, 128, (4, 64, 64, 64)): 0, (128, 0.3, 128, (67, 128, 128, 256)): 1, (None, None, None, (259, 512, 512, 1024)): 2}


def build_shared_mlp(mlp_spec: List[int], bn: bool = True) -> nn.Sequential:
    if bn:
        raise NotImplementedError("MPiNets builds its SA modules with bn=False (model.py:371,380,383)")
    layers: List[nn.Module] = []
    for i in range(1, len(mlp_spec)):
        layers += [nn.Conv2d(mlp_spec[i - 1], mlp_spec[i], kernel_size=1, bias=True), nn.ReLU(True)]
    return nn.Sequential(*layers)


class PointnetSAModule(nn.Module):
    def __init__(self, mlp: List[int], npoint: Optional[int] = None, radius: Optional[float] = None,
                 nsample: Optional[int] = None, bn: bool = True, use_xyz: bool = True):
        super().__init__()
        spec = list(mlp)
        if use_xyz:
            spec[0] += 3
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.mlps = nn.ModuleList([build_shared_mlp(spec, bn)])
        key = (npoint, radius, nsample, tuple(spec))
        if key not in _MODULE_OF:
            raise NotImplementedError(f"only the three MPiNets set-abstraction configurations are built in: {key}")
        self.module_index = _MODULE_OF[key]

    def forward(self, xyz: torch.Tensor, features: torch.Tensor, precision: int = _lib.PREC_FP32):
        """xyz [B,N,3], features [B,C,N] -> (new_xyz [B,npoint,3] | None, new_features [B,C_out,npoint])"""
        eng = get_engine(xyz.device)
        if not eng._weights_loaded:
            raise RuntimeError("load the network weights into the engine first (MotionPolicyNetwork.sync_engine())")
        if precision != _lib.PREC_FP32 and self.module_index == 2:
            raise NotImplementedError("the group-all module has a per-module entry in the fp32 mode only (the tensor-core modes run it "
                                      "inside MotionPolicyNetwork.forward / MPiNetsPointNet)")
        if precision != _lib.PREC_FP32 and self.module_index == 0:
            # the tensor-core SA1 kernels read [B,N,4] cloud rows (xyz + the mask feature) as 16-byte rows: re-join _break_up_pc's halves
            cloud = torch.cat([xyz[..., :3], features.transpose(1, 2)], dim=-1).contiguous()
            new_xyz, out = eng.sa_forward(0, cloud, cloud[..., 3:], precision=precision)
            return new_xyz, out.transpose(1, 2).contiguous()
        new_xyz, out = eng.sa_forward(self.module_index, xyz.contiguous(), features.transpose(1, 2).contiguous(), precision=precision)
        return new_xyz, out.transpose(1, 2).contiguous()
