"""GPU tests of the packed row tiles of the tensor-core set-abstraction kernels (tc_common.cuh: TilePack).

pointnet2's ball query pads a group of H < 128 hits with copies of its first hit (ball_query_gpu.cu semantics) and
the max-pool of `PointnetSAModule` (model.py:365-382) ignores duplicates, so the kernels push only the distinct rows of a group through
the shared MLP, several groups per 128-row MMA tile.  A row's MLP output does not depend on which tile it sits in, hence the bar here is
BIT equality with the unpacked launch (MPN_SA_NOPACK=1: one tile per group, the reference formulation) -- in both tensor-core modes,
for both modules, with and without the index-ordered neighbour lists, on regular scenes and on adversarial clouds (> 128 hits,
the linear-scan fallback, groups of a single point)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene_cloud(engine_w, oracle, tables, config, B):
    from mpinets_b200 import scenes
    p = scenes.config_problems(config, B)
    return oracle.build_cloud(p["q0"], p["target"], p, tables, engine_w.cfg.seed)


def _adversarial_cloud():
    rng = np.random.RandomState(11)
    N = 6272
    cloud = np.zeros((4, N, 4), np.float32)
    cloud[..., :3] = rng.uniform(-1.0, 1.5, size=(4, N, 3))
    cloud[0, 0:300, :3] = 0.5 + rng.uniform(-0.02, 0.02, size=(300, 3))         # > 128 hits: the 128 smallest indices survive
    cloud[1, 0:1500, :3] = -0.3 + rng.uniform(-0.015, 0.015, size=(1500, 3))     # > 256 candidates: linear-scan fallback
    cloud[2, :, :3] = np.round(cloud[2, :, :3] / 0.0501) * 0.0501                # points on grid-cell corners
    cloud[3, :, :3] = rng.uniform(-4.0, 4.0, size=(N, 3))                        # sparse: almost every group is a single point
    cloud[..., 3] = rng.randint(0, 3, size=(4, N))
    return cloud


def _run(engine_w, module, xyz, feats, precision, debug, nopack):
    if nopack:
        os.environ["MPN_SA_NOPACK"] = "1"
    try:
        engine_w.sa_tile_counts(reset=True)
        out = engine_w.sa_forward(module, xyz, feats, precision=precision, debug=debug)
        tiles = engine_w.sa_tile_counts(reset=True)
    finally:
        os.environ.pop("MPN_SA_NOPACK", None)
    assert not engine_w.tc_error()
    return out, tiles


@pytest.mark.parametrize("mode", ["bf16", "bf16x3"])
@pytest.mark.parametrize("case", ["tabletop", "mixed", "adversarial"])
def test_packed_tiles_equal_unpacked(engine_w, oracle, tables, mode, case):
    from mpinets_b200 import _lib
    prec = _lib.PRECISIONS[mode]
    cloud = _adversarial_cloud() if case == "adversarial" else _scene_cloud(engine_w, oracle, tables, 2 if case == "tabletop" else 4, 5)
    B = cloud.shape[0]
    d = torch.from_numpy(cloud).cuda()
    # ---- SA1
    (nx, f_ref, _, bi_ref), t_ref = _run(engine_w, 0, d, d[..., 3:], prec, True, True)      # unpacked, ordered lists
    (_, f_pk, _, bi_pk), t_pk = _run(engine_w, 0, d, d[..., 3:], prec, True, False)         # packed, ordered lists
    (_, f_set), t_set = _run(engine_w, 0, d, d[..., 3:], prec, False, False)                # packed, hits left in bucket order
    assert t_ref[0] == B * 512 and t_ref[1] == 0                                             # one tile per group = the reference formulation
    assert torch.equal(bi_ref, bi_pk)
    assert torch.equal(f_ref, f_pk) and torch.equal(f_ref, f_set)
    assert t_pk == t_set and B * 512 / 4 <= t_pk[0] <= B * 512
    # the packing is the first fit the header describes: recompute the tile count from the neighbour lists
    counts = oracle.distinct_neighbour_counts(bi_ref.cpu().numpy())
    assert t_pk[0] == sum(oracle.packed_tile_count(counts[b], per_round=4, gran=32) for b in range(B))
    print(f"SA1 {mode} {case}: {t_pk[0] / (B * 512):.3f} tiles per group (distinct neighbours: mean {counts.mean():.1f}, max {counts.max()})")
    # ---- SA2 on SA1's output
    xyz1 = nx.contiguous()
    (_, g_ref, _, bj_ref), u_ref = _run(engine_w, 1, xyz1, f_ref.contiguous(), prec, True, True)
    (_, g_pk, _, bj_pk), u_pk = _run(engine_w, 1, xyz1, f_ref.contiguous(), prec, True, False)
    (_, g_set), _ = _run(engine_w, 1, xyz1, f_ref.contiguous(), prec, False, False)
    assert u_ref[1] == B * 128 and torch.equal(bj_ref, bj_pk)
    assert torch.equal(g_ref, g_pk) and torch.equal(g_ref, g_set)
    counts2 = oracle.distinct_neighbour_counts(bj_ref.cpu().numpy())
    rule = dict(per_round=8, gran=16) if mode == "bf16x3" else dict(per_round=4, gran=32)   # sa2x3h: eighths of 8 centroids; sa2w3: quarters of 4
    assert u_pk[1] == sum(oracle.packed_tile_count(counts2[b], **rule) for b in range(B))
    print(f"SA2 {mode} {case}: {u_pk[1] / (B * 128):.3f} tiles per group")


def test_sa2x3_chain_variants_agree(engine_w, oracle, tables):
    """the 8-warps-per-chain SA2 kernel of the parity-grade mode (default) against the 4-warp one (MPN_SA2X3_V1=1): same bits"""
    from mpinets_b200 import _lib
    cloud = _scene_cloud(engine_w, oracle, tables, 4, 6)
    d = torch.from_numpy(cloud).cuda()
    nx, f1 = engine_w.sa_forward(0, d, d[..., 3:], precision=_lib.PREC_BF16X3)
    _, g8 = engine_w.sa_forward(1, nx.contiguous(), f1.contiguous(), precision=_lib.PREC_BF16X3)
    os.environ["MPN_SA2X3_V1"] = "1"
    try:
        _, g4 = engine_w.sa_forward(1, nx.contiguous(), f1.contiguous(), precision=_lib.PREC_BF16X3)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("MPN_SA2X3_V1", None)
    assert not engine_w.tc_error()
    assert torch.equal(g8, g4)
