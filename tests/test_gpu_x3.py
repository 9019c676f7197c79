"""GPU parity tests of the parity-grade tensor-core mode MPN_PREC_BF16X3 (split-bf16 operands on tcgen05, sa_x3.cu / gemm_tc.cu).

The bar is north_star's: FPS / ball-query indices and collision flags bit-exact, delta-q within 1e-5 of the **fp32** oracle
(the reference runs model.py:75-91 in plain fp32) -- no rounding-emulating oracle is involved in this mode."""
import os

import numpy as np
import pytest
import torch

from conftest import to_dev

pytestmark = pytest.mark.gpu

DQ_TOL = 1e-5      # north_star: "delta-q within 1e-5 fp32"
FEAT_RTOL = 2e-5   # per-module features, relative to the module's largest activation (measured ~2e-6)


def _problems(config, B, seed=0x4D50694E, problem0=0):
    from mpinets_b200 import scenes
    return scenes.config_problems(config, B, seed, problem0)


def _sa_weights(sd, m):
    return [(sd[f"point_cloud_encoder.SA_modules.{m}.mlps.0.{2 * l}.weight"], sd[f"point_cloud_encoder.SA_modules.{m}.mlps.0.{2 * l}.bias"])
            for l in range(3)]


@pytest.mark.parametrize("M,N,K", [(300, 128, 80), (512, 512, 272), (130, 1024, 512), (7, 4096, 1024), (256, 2048, 4096)])
def test_split_gemm_matches_fp64(engine, M, N, K):
    """the three-pass split-bf16 TMA GEMM behind the per-point layer of SA2, SA3 and the FC head, vs a float64 product"""
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    ref = (a.double() @ w.double().t() + b.double())
    got = engine.tc_gemm_selftest(a.cuda(), w.cuda(), b.cuda(), split=True).cpu().double()
    assert not engine.tc_error()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    got1 = engine.tc_gemm_selftest(a.cuda(), w.cuda(), b.cuda(), split=False).cpu().double()
    err1 = (got1 - ref).abs().max().item() / ref.abs().max().item()
    f32 = (a @ w.t() + b).double()
    print(f"split GEMM M={M} N={N} K={K}: rel err split {err:.2e}, plain bf16 {err1:.2e}, torch fp32 {(f32 - ref).abs().max().item() / ref.abs().max().item():.2e}")
    assert err < 2e-5 and err1 < 2e-2


def test_sa_modules_x3(engine_w, oracle, tables, state_dict):
    from mpinets_b200 import _lib
    p = _problems(4, 6)
    cloud = oracle.build_cloud(p["q0"], p["target"], p, tables, engine_w.cfg.seed)
    d_cloud = torch.from_numpy(cloud).cuda()
    xyz = np.ascontiguousarray(cloud[..., :3])
    feats = torch.from_numpy(np.ascontiguousarray(cloud[..., 3:]))
    o_xyz1, o_f1, aux1 = oracle.sa_module(xyz, feats, oracle.SA_SPECS[0], _sa_weights(state_dict, 0), return_aux=True)
    nx, f1, fi1, bi1 = engine_w.sa_forward(0, d_cloud, d_cloud[..., 3:], precision=_lib.PREC_BF16X3, debug=True)
    assert not engine_w.tc_error()
    assert np.array_equal(nx.cpu().numpy(), o_xyz1)
    assert np.array_equal(fi1.cpu().numpy(), aux1["fps_idx"])
    assert np.array_equal(bi1.cpu().numpy(), aux1["ball_idx"])
    o64 = oracle.sa_module(xyz, feats, oracle.SA_SPECS[0], _sa_weights(state_dict, 0), dtype=torch.float64)[1]
    e1 = (f1.cpu().double() - o64).abs().max().item() / o64.abs().max().item()
    e1_32 = (o_f1.double() - o64).abs().max().item() / o64.abs().max().item()
    print(f"SA1 bf16x3 vs fp64 oracle: rel err {e1:.2e} (fp32 oracle itself: {e1_32:.2e})")
    assert e1 < FEAT_RTOL
    # SA2 fed with the oracle's SA1 output
    o_xyz2, o_f2, aux2 = oracle.sa_module(o_xyz1, o_f1, oracle.SA_SPECS[1], _sa_weights(state_dict, 1), return_aux=True)
    nx2, f2, fi2, bi2 = engine_w.sa_forward(1, torch.from_numpy(o_xyz1).cuda(), o_f1.contiguous().cuda(), precision=_lib.PREC_BF16X3, debug=True)
    assert not engine_w.tc_error()
    assert np.array_equal(nx2.cpu().numpy(), o_xyz2)
    assert np.array_equal(bi2.cpu().numpy(), aux2["ball_idx"])
    o64_2 = oracle.sa_module(o_xyz1, o_f1, oracle.SA_SPECS[1], _sa_weights(state_dict, 1), dtype=torch.float64)[1]
    e2 = (f2.cpu().double() - o64_2).abs().max().item() / o64_2.abs().max().item()
    e2_32 = (o_f2.double() - o64_2).abs().max().item() / o64_2.abs().max().item()
    print(f"SA2 bf16x3 vs fp64 oracle: rel err {e2:.2e} (fp32 oracle itself: {e2_32:.2e})")
    assert e2 < FEAT_RTOL


@pytest.mark.parametrize("config", [2, 3, 4])
def test_policy_forward_x3_within_1e5(engine_w, oracle, tables, state_dict, config):
    """delta-q of the tensor-core parity mode vs the fp32 oracle at B = 64 on the scene mixes of configs[1..3]"""
    from mpinets_b200 import _lib
    B = 64
    p = _problems(config, B)
    sc = to_dev(p)
    q0, tg = torch.from_numpy(p["q0"]).cuda(), torch.from_numpy(p["target"]).cuda()
    cloud = engine_w.build_cloud(sc, q0, tg)
    qn = oracle.normalize(p["q0"], tables.joint_limits)
    dq = engine_w.policy_forward(cloud, torch.from_numpy(qn).cuda(), _lib.PREC_BF16X3).cpu()
    assert not engine_w.tc_error()
    dq32 = engine_w.policy_forward(cloud, torch.from_numpy(qn).cuda(), _lib.PREC_FP32).cpu()
    ch = cloud.cpu().numpy()
    exp = oracle.policy_forward(state_dict, ch, qn)
    err = (dq - exp).abs().max().item()
    print(f"config {config}: bf16x3 delta-q max-abs-err vs fp32 oracle {err:.3e}; fp32 SIMT mode vs oracle {(dq32 - exp).abs().max().item():.3e}; "
          f"bf16x3 vs fp32 mode {(dq - dq32).abs().max().item():.3e}; |dq| max {exp.abs().max().item():.3f}")
    assert err <= DQ_TOL
    enc = engine_w.encoder_forward(cloud, _lib.PREC_BF16X3).cpu()
    enc32 = engine_w.encoder_forward(cloud, _lib.PREC_FP32).cpu()
    assert (enc - enc32).abs().max().item() <= 2e-5 * max(1.0, enc32.abs().max().item())


def test_rollout_x3_flags_and_drift(engine_w, oracle, tables, state_dict):
    """20 lock-step steps in the bf16x3 mode vs the fp32 SIMT mode of the same library on 32 mixed problems: step 1 within 1e-5
    (identical inputs), identical collision flags and first-collision steps, bounded drift afterwards; flags bit-exact against
    the oracle sweep of the produced trajectory."""
    from mpinets_b200 import _lib
    B, T = 32, 20
    p = _problems(4, B)
    sc = to_dev(p)
    q0, tg = torch.from_numpy(p["q0"]).cuda(), torch.from_numpy(p["target"]).cuda()
    c3 = engine_w.build_cloud(sc, q0, tg)
    c32 = c3.clone()
    traj3, m3 = engine_w.rollout(sc, c3, q0, tg, T, check_every_step=True, precision=_lib.PREC_BF16X3)
    assert not engine_w.tc_error()
    traj32, m32 = engine_w.rollout(sc, c32, q0, tg, T, check_every_step=True, precision=_lib.PREC_FP32)
    rng_ = torch.from_numpy((tables.joint_limits[:, 1] - tables.joint_limits[:, 0]) / 2).cuda()
    d = ((traj3 - traj32).abs() / rng_).amax(dim=(0, 2)).cpu().numpy()
    print("bf16x3 vs fp32 mode, max normalised |dq| drift per step:", np.array2string(d, precision=2))
    assert d[1] <= DQ_TOL
    assert d.max() <= 1e-3
    assert torch.equal(m3[:, :2], m32[:, :2])
    th = traj3.cpu().numpy()
    oflags, ofirst, _ = oracle.sweep_flags(p, th, tables)
    assert np.array_equal(m3[:, 0].cpu().numpy().astype(np.uint8), oflags)
    assert np.array_equal(m3[:, 1].cpu().numpy().astype(np.int32), ofirst)
    otraj = oracle.rollout(state_dict, oracle.build_cloud(p["q0"][:4], p["target"][:4], {k: v[:4] for k, v in p.items()}, tables, engine_w.cfg.seed),
                           oracle.normalize(p["q0"][:4], tables.joint_limits), tables, 2, engine_w.cfg.seed)
    assert (np.abs(th[:4, 1] - otraj[:, 1]) / rng_.cpu().numpy()).max() <= DQ_TOL


def test_x3_full_batch_strided_subset(engine_w, oracle, tables, state_dict):
    """the benchmarked size: 4096 problems per launch; every 257th problem (CTA indices 0 .. 3855) against the fp32 oracle"""
    from mpinets_b200 import _lib
    B = 4096
    p = _problems(2, B)
    sc = to_dev(p)
    q0, tg = torch.from_numpy(p["q0"]).cuda(), torch.from_numpy(p["target"]).cuda()
    cloud = engine_w.build_cloud(sc, q0, tg)
    qn = oracle.normalize(p["q0"], tables.joint_limits)
    dq = engine_w.policy_forward(cloud, torch.from_numpy(qn).cuda(), _lib.PREC_BF16X3)
    dqb = engine_w.policy_forward(cloud, torch.from_numpy(qn).cuda(), _lib.PREC_BF16)
    assert not engine_w.tc_error()
    sub = np.arange(0, B, 257)
    ch = cloud[torch.from_numpy(sub).cuda()].cpu().numpy()
    exp = oracle.policy_forward(state_dict, ch, qn[sub])
    err = (dq[torch.from_numpy(sub).cuda()].cpu() - exp).abs().max().item()
    errb = (dqb[torch.from_numpy(sub).cuda()].cpu() - exp).abs().max().item()
    print(f"B=4096 strided subset ({len(sub)} problems): bf16x3 delta-q err {err:.3e}, bf16 mode {errb:.3e}")
    assert err <= DQ_TOL
    assert errb <= 3e-4
    # all problems: the two tensor-core modes agree to the bf16 mode's error everywhere (no CTA-index-dependent fault)
    assert (dq - dqb).abs().max().item() <= 5e-4
