"""CPU test of the bench contract's reference arm: `bench.py --impl reference` runs without a GPU, prints exactly one JSON
line on stdout and carries the keys the driver reads (the GPU arm of bench.py is exercised on the B200 box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_roofline_accounting_of_packed_tiles():
    """bench.py's roofline of a set-abstraction stage: `achieved` is the contract's figure (SURVEY 8(d) flops of the reference formulation
    per launch / launch time), the executed and issued-MMA rates follow from the device's tile count (mpn_sa_tile_counts)."""
    sys.path.insert(0, ROOT)
    import bench
    B = 4096
    stages = {"fps1": {"ms": 5.0, "launches": 1}, "sa1": {"ms": 5.6, "launches": 1}, "sa2": {"ms": 7.0, "launches": 1},
              "sa3": {"ms": 2.4, "launches": 1}}
    per_stage = {k: v["ms"] for k, v in stages.items()}
    peaks = dict(hbm_gbs=6552.0, bf16_tflops=1636.2, bf16_sustained=1407.1, source="test")
    tiles = {"sa1": 0.28 * 512 * B, "sa2": 0.33 * 128 * B}
    r = bench.roofline_of(stages, per_stage, B, peaks, "bf16x3", {}, tiles)
    assert r["kernel"] == "sa2" and r["bound"] == "tensor"
    ref = 2 * 945_815_552 * B
    assert abs(r["achieved"] - ref / 7.0e-3 / 1e12) < 1e-6 and abs(r["frac"] - r["achieved"] / 1407.1) < 1e-12
    assert abs(r["tiles_per_group"] - 0.33) < 1e-9
    assert abs(r["executed_flop_per_launch"] - ref * 0.33) < 1e3                       # the distinct rows only
    issued = tiles["sa2"] * 72 * 2 * 128 * 128 * 16 + 3 * 2 * 512 * 80 * 128 * B      # 72 MMAs per tile + the per-point layer-1 GEMM
    assert abs(r["issued_mma_flop_per_launch"] - issued) < 1e3
    assert r["issued_mma_frac_of_peak"] < 1.0 < r["achieved"] / r["executed_tflops"]
    # one tile per group (MPN_SA_NOPACK) reproduces the reference formulation: executed == algorithmic, issued == 3 passes of layers 2-3
    full = bench.sa_work("sa2", B, "bf16x3", 128 * B)
    assert abs(full["executed_flop"] - ref) < 1e3 and abs(full["tiles_per_group"] - 1.0) < 1e-12
    k = bench.tensor_kernels_of(stages, tiles, B, "bf16x3", peaks)
    assert set(k) == {"sa1", "sa2", "sa3"} and abs(k["sa1"]["tiles_per_group"] - 0.28) < 1e-9
    assert abs(k["sa3"]["executed_tflops"] - k["sa3"]["TFLOPs"]) < 1e-9                 # no packing in the group-all level
