"""mpinets_b200 -- B200 (sm_100a) rollout engine for Motion Policy Networks behind the reference's Python surface.

Public API mirrors the reference (`/root/reference/mpinets`): ``model.MotionPolicyNetwork``, ``model.MPiNetsPointNet``,
``geometry.TorchCuboids / TorchCylinders / construct_mixed_point_cloud``, ``utils.(un)normalize_franka_joints``,
``pointnet2_utils`` / ``pointnet2_modules`` (pointnet2_ops names) and ``robofin_shim.FrankaSampler /
FrankaCollisionSampler``.  All arithmetic runs in ``libmpinets_b200.so`` (``include/mpinets_b200.h``); there is no CPU path.
"""
__version__ = "0.1.0"
