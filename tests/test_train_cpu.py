"""CPU tests of the training-step oracle and of the data-parallel host logic (no GPU).

* the oracle's parameter gradients (torch.autograd over the torch restatement + the C loss gradient) are checked against
  central finite differences of the oracle's own weighted loss -- this is what pins the checker the GPU parity test uses;
* DDP gradient averaging over world_size-2 gloo: the mean of the two ranks' half-batch gradients equals the full-batch
  gradient (losses are batch means, model.py:235-238 + DDP averaging)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 0x4D50694E


def _batch(B, seed=0):
    sys.path.insert(0, ROOT)
    from mpinets_b200 import scenes, franka
    from oracle import oracle as O
    tables = franka.default_tables()
    p = scenes.config_problems(4, B, SEED, 0)
    rng = np.random.default_rng(seed)
    cloud = O.build_cloud(p["q0"], p["target"], p, tables, SEED)
    qn = O.normalize(p["q0"], tables.joint_limits)
    qn[B // 2:] = rng.uniform(-0.9, 0.9, (B - B // 2, 7)).astype(np.float32)
    O.sample_robot(O.unnormalize(qn, tables.joint_limits), tables, 2048, SEED, 0, cloud)
    sup = np.clip(qn + rng.normal(scale=0.05, size=qn.shape), -1, 1).astype(np.float32)
    return O, tables, p, cloud, qn, sup


def test_oracle_gradients_match_finite_differences():
    O, tables, p, cloud, qn, sup = _batch(2)
    sd = O.reference_state_dict(0)
    wc, wb = 5.0, 1.0
    losses, yh, grads, gy = O.train_step_grads(sd, cloud, qn, sup, p, tables, SEED, w_collision=wc, w_bc=wb, dtype=torch.float64)
    assert losses[1] > 0

    def total(sd_):
        dq = O.policy_forward(sd_, cloud, qn, False, torch.float64).numpy()
        y = np.clip(qn.astype(np.float64) + dq, -1, 1).astype(np.float32)
        l, _ = O.bc_collision_losses(p, y, sup, tables, SEED, 1024, 0.03, wc, wb)
        return wc * float(l[0]) + wb * float(l[1])

    # the loss itself is evaluated in fp32 (C restatement): use steps large enough to clear its rounding, on entries with
    # the largest gradients of a late, a middle and an early tensor
    for key, eps in (("decoder.6.bias", 2e-3), ("point_cloud_encoder.fc_layer.6.bias", 2e-2), ("feature_encoder.8.bias", 2e-2)):
        g = grads[key].reshape(-1)
        i = int(np.abs(g).argmax())
        sp = {k: v.clone().double() for k, v in sd.items()}
        sm = {k: v.clone().double() for k, v in sd.items()}
        sp[key].view(-1)[i] += eps
        sm[key].view(-1)[i] -= eps
        fd = (total(sp) - total(sm)) / (2 * eps)
        assert abs(fd - g[i]) < 0.05 * abs(g[i]) + 1e-6, (key, fd, g[i])


def test_param_plan_matches_reference_state_dict():
    """the flat layout documented in include/mpinets_b200.h is the reference state-dict order with 4-float alignment"""
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    sd = O.reference_state_dict(0)
    off = 0
    for k, v in sd.items():
        off += (v.numel() + 3) // 4 * 4
    assert sum(v.numel() for v in sd.values()) == 19068103
    assert off - 19068103 < 4 * len(sd)


def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from mpinets_b200.parallel import allreduce_mean_, shard_range
    O, tables, p, cloud, qn, sup = _batch(4)
    lo, hi = shard_range(rank, world, 4)
    sl = slice(lo, hi)
    ps = {k: (v[sl] if isinstance(v, np.ndarray) and v.shape[:1] == (4,) else v) for k, v in p.items()}
    sd = O.reference_state_dict(0)
    _, _, grads, _ = O.train_step_grads(sd, cloud[sl], qn[sl], sup[sl], ps, tables, SEED, dtype=torch.float64)
    flat = torch.cat([torch.from_numpy(g.reshape(-1)) for g in grads.values()])
    allreduce_mean_(flat)
    if rank == 0:
        np.save(out, flat.numpy())
    dist.destroy_process_group()


def test_two_rank_gradient_average_equals_full_batch(tmp_path):
    out = str(tmp_path / "g.npy")
    port = 31500 + os.getpid() % 2000
    mp.spawn(_ddp_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    O, tables, p, cloud, qn, sup = _batch(4)
    sd = O.reference_state_dict(0)
    _, _, grads, _ = O.train_step_grads(sd, cloud, qn, sup, p, tables, SEED, dtype=torch.float64)
    exp = np.concatenate([g.reshape(-1) for g in grads.values()])
    assert np.abs(got - exp).max() < 1e-6 * np.abs(exp).max()
