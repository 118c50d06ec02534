"""Host-side logic of the product (no GPU): cutout sampler, resize tables, sharding, shapes, error behaviour."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import sampler as sampler_oracle
from perceptor_b200 import cutouts, losses, models
from perceptor_b200.resize_tables import CUBIC, LANCZOS3, ResizeTableCache, build_dim_table, choose_method
from perceptor_b200.vit import SHAPES, random_state_dict, required_keys, resolve_shape

GOLDEN = Path(__file__).parent / "golden"


def test_cutout_rows_bit_exact_against_oracle_and_golden():
    z = np.load(GOLDEN / "sampler_rows.npz")
    for j in range(4):
        seed, b, h, w, n, lo, hi = (int(v) for v in z[f"args{j}"])
        pw = float(z[f"pow{j}"])
        rows = cutouts.sample_cutouts(torch.Generator().manual_seed(seed), b, h, w, n, pw, lo, hi)
        assert rows.dtype == np.int32 and rows.shape == (b * n, 4)
        assert np.array_equal(rows, z[f"rows{j}"])
        ref = sampler_oracle.sample_cutouts(torch.Generator().manual_seed(seed), b, h, w, n, pw, lo, hi)
        assert rows.tolist() == [list(r) for r in ref]
        assert (rows[:, 3] >= lo).all() and (rows[:, 3] <= hi).all()
        assert (rows[:, 1] >= 0).all() and (rows[:, 1] + rows[:, 3] <= h).all()
        assert (rows[:, 2] >= 0).all() and (rows[:, 2] + rows[:, 3] <= w).all()
        assert np.array_equal(rows[:, 0], np.repeat(np.arange(b), n))


def test_cutout_sampler_edges():
    g = torch.Generator().manual_seed(0)
    rows = cutouts.sample_cutouts(g, 1, 64, 64, 3, 1.0, 64, 64)  # min == max == side: only one legal box
    assert rows.tolist() == [[0, 0, 0, 64]] * 3
    with pytest.raises(ValueError):
        cutouts.sample_cutouts(g, 1, 64, 64, 3, 1.0, 65, 80)
    with pytest.raises(ValueError):
        cutouts.sample_cutouts(g, 1, 64, 64, 0)
    whole = cutouts.whole_image_cutouts(2, 30, 50)
    assert whole.tolist() == [[0, 0, 0, 30, 50], [1, 0, 0, 30, 50]]


def test_generator_stream_is_consumed_in_order():
    g = torch.Generator().manual_seed(5)
    a = cutouts.sample_cutouts(g, 1, 128, 128, 4, 1.0, 32, 128)
    b = cutouts.sample_cutouts(g, 1, 128, 128, 4, 1.0, 32, 128)
    g2 = torch.Generator().manual_seed(5)
    a2 = cutouts.sample_cutouts(g2, 1, 128, 128, 4, 1.0, 32, 128)
    assert np.array_equal(a, a2) and not np.array_equal(a, b)


def test_shard_rows_partitions_exactly():
    for n in (0, 1, 7, 128, 256, 257):
        for world in (1, 2, 3, 4, 8):
            parts = [cutouts.shard_rows(n, r, world) for r in range(world)]
            covered = [i for s in parts for i in range(s.start, s.stop)]
            assert covered == list(range(n))
            sizes = [s.stop - s.start for s in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        cutouts.shard_rows(4, 2, 2)


def test_resize_tables_match_reference_golden():
    z = np.load(GOLDEN / "resize_tables.npz")
    pairs = sorted({tuple(int(v) for v in k.split("_")[1:]) for k in z.files})
    for in_sz, out_sz in pairs:
        t = build_dim_table(in_sz, out_sz, LANCZOS3 if in_sz >= out_sz else CUBIC)
        assert np.array_equal(t.left, z[f"left_{in_sz}_{out_sz}"]), "tap indices must be bit-exact"
        assert t.weights.shape == z[f"w_{in_sz}_{out_sz}"].shape
        assert float(np.abs(t.weights - z[f"w_{in_sz}_{out_sz}"]).max()) <= 1e-6
        assert np.allclose(t.weights.sum(1), 1.0, atol=1e-5)


def test_resize_table_structure():
    t = build_dim_table(224, 224, LANCZOS3)  # identity: the reference skips dims with scale 1.0
    assert t.taps == 1 and np.array_equal(t.left, np.arange(224)) and (t.weights == 1).all()
    for in_sz, out_sz, m in [(512, 224, LANCZOS3), (100, 224, CUBIC), (33, 32, LANCZOS3), (300, 336, CUBIC)]:
        t = build_dim_table(in_sz, out_sz, m)
        assert (np.diff(t.left) >= 0).all(), "left boundaries must be monotone (the backward relies on it)"
        # inverse ranges: output o reads input i  <=>  inv[i,0] <= o < inv[i,1] (within tap range)
        for i in range(in_sz):
            reads = [o for o in range(out_sz) if t.left[o] <= i < t.left[o] + t.taps]
            lo, hi = t.inv[i]
            assert reads == list(range(lo, hi)), (in_sz, out_sz, i)
    assert choose_method(300, 300, 224, 224) == LANCZOS3
    assert choose_method(300, 200, 224, 224) == CUBIC and choose_method(100, 100, 224, 224) == CUBIC
    with pytest.raises(ValueError):
        build_dim_table(0, 224, LANCZOS3)


def test_resize_table_cache_layout():
    c = ResizeTableCache(32)
    a = c.table_id(64, LANCZOS3)
    b = c.table_id(20, CUBIC)
    assert c.table_id(64, LANCZOS3) == a and a != b
    desc, left, weight, inv = c.flat_arrays()
    assert desc.shape == (2, 8) and desc.dtype == np.int32
    for tid in (a, b):
        t = c.table(tid)
        taps, lo, wo, io, in_size = desc[tid, :5]
        assert taps == t.taps and in_size == t.in_size
        assert np.array_equal(left[lo:lo + 32], t.left)
        assert np.array_equal(weight[wo:wo + 32 * taps].reshape(32, taps), t.weights)
        assert np.array_equal(inv[io:io + 2 * in_size].reshape(in_size, 2), t.inv)


def test_shapes_and_flop_model_match_baseline_table():
    # total GF per cutout
    want = {"ViT-B-32": 17.73, "ViT-B-16": 71.68, "ViT-L-14": 330.54, "ViT-L-14-336": 796.57}
    for name, gf in want.items():
        assert abs(SHAPES[name].flops_per_cutout() / 1e9 - gf) <= 0.02, name
    assert SHAPES["ViT-L-14"].tokens == 257 and SHAPES["ViT-L-14"].kpad == 640 and SHAPES["ViT-B-32"].kpad == 3072
    assert resolve_shape("ViT-B-32-quickgelu")[0] == "ViT-B-32" and resolve_shape("ViT-L-14-336px")[0] == "ViT-L-14-336"
    with pytest.raises(ValueError):
        resolve_shape("ViT-bigG-14")
    # wide heads (the reference's OpenCLIP default is ViT-H-14): head dim 80 / 88, stored padded to 128 columns
    h, g14 = SHAPES["ViT-H-14"], SHAPES["ViT-g-14"]
    assert (h.head_dim, h.head_stride, h.mlp, h.tokens) == (80, 128, 5120, 257)
    assert (g14.head_dim, g14.head_stride, g14.mlp) == (88, 128, 6144) and SHAPES["ViT-L-14"].head_stride == 64
    assert abs(h.flops_per_cutout() / 1e9 - 680.0) <= 3.0  # "ViT-H/14: 680 GF"
    sd = random_state_dict(SHAPES["ViT-B-32"], 0)
    assert set(required_keys(12)) == set(sd)
    assert sd["conv1.weight"].shape == (768, 3, 32, 32) and sd["proj"].shape == (768, 512)


def test_executed_flop_model_of_the_class_token_last_block():
    """flops_per_cutout(pooled_last_block=True) counts what the default sequencer executes: the last block keeps only its
    K / V projections (and the full dqkv dgrad GEMM) on every token; counted by hand for ViT-L/14."""
    s = SHAPES["ViT-L-14"]
    t, d, mlp = s.tokens, s.width, s.mlp
    per_token = 8 * d * d + 4 * d * mlp
    dense, pooled = s.flops_per_cutout(), s.flops_per_cutout(pooled_last_block=True)
    saved_fwd = t * per_token - (t * 4 * d * d + (4 * d * d + 4 * d * mlp)) + 4 * t * t * d - 4 * t * d
    saved_bwd = t * per_token - (t * 6 * d * d + (2 * d * d + 4 * d * mlp)) + 8 * t * t * d - 8 * t * d
    assert abs((dense - pooled) - (saved_fwd + saved_bwd)) <= 1e-6 * dense
    assert 0.96 < pooled / dense < 0.97  # 3.3 % of the step for ViT-L/14
    for name in SHAPES:
        assert SHAPES[name].flops_per_cutout(True) < SHAPES[name].flops_per_cutout()


def test_resize_table_ids_from_the_lookup_array():
    """GuidanceEngine._table_ids_for_sizes (the per-step host path) against ResizeTableCache.table_id, on a stub engine:
    first use fills the lookup array, sizes beyond it grow it, ids stay those of the cache."""
    from types import SimpleNamespace

    from perceptor_b200.guidance import GuidanceEngine
    from perceptor_b200.resize_tables import ResizeTableCache

    stub = SimpleNamespace(tables=ResizeTableCache(32), shape=SimpleNamespace(image_size=32))
    stub._method_for_size = lambda size: GuidanceEngine._method_for_size(stub, size)
    rng = np.random.default_rng(0)
    for top in (40, 40, 1500):
        sizes = rng.integers(8, top, size=64).astype(np.int32)
        got = GuidanceEngine._table_ids_for_sizes(stub, sizes)
        want = np.array([stub.tables.table_id(int(v), stub._method_for_size(int(v))) for v in sizes], dtype=np.int32)
        assert got.dtype == np.int32 and np.array_equal(got, want)


def test_module_surface_and_error_behaviour_without_gpu(tmp_path):
    with pytest.raises(ValueError):
        models.OpenCLIP("ViT-B-32", "not-a-weight-name")
    with pytest.raises(ValueError):
        losses.CLIP("RN50")
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    loss = losses.CLIP("ViT-B-32", n_cutouts=4)
    assert loss.name == "ViT-B-32" and loss.multiplier == 1.0 and loss.encodings is None and loss.weights is None
    assert loss.mul_(0.5) is loss and loss.multiplier == 0.5
    enc = torch.randn(2, 512)
    assert loss.add_encodings_(enc, [1.0, 2.0]) is loss
    assert isinstance(loss.encodings, torch.nn.Parameter) and not loss.encodings.requires_grad
    assert torch.allclose(loss.encodings.norm(dim=1), torch.ones(2), atol=1e-6)  # CLIP re-normalises targets
    loss.add_encodings_(torch.randn(1, 512))
    assert loss.encodings.shape == (3, 512) and loss.weights.tolist() == [1.0, 2.0, 1.0]
    assert "encodings" in loss.state_dict() and "weights" in loss.state_dict()
    assert loss.model.image_size == (224, 224) and loss.device.type == "cpu"
    assert models.CLIP("ViT-B-32", "fp32") is loss.model  # weak-valued constructor cache shares the encoder
    # no CPU fallback: the image path must fail loudly off-GPU
    with pytest.raises(RuntimeError):
        loss(torch.rand(1, 3, 64, 64))
    # add_texts_ works out of the box with the packaged CLIP merge table (text tower in plain torch, CPU is fine)
    before = loss.encodings.shape[0]
    loss.add_texts_(["hello world"], [0.5])
    assert loss.encodings.shape == (before + 1, 512) and abs(float(loss.encodings[-1].norm()) - 1.0) < 1e-5
    with pytest.raises(FileNotFoundError, match="merge table"):
        losses.CLIP("ViT-B-32", bpe_path=str(tmp_path / "missing.gz")).add_texts_(["hello"])
    stub = tmp_path / "textoff.json"
    stub.write_text('{"ViT-B-32": [[0.1, 0.2]]}')
    with pytest.raises(ValueError, match="There is no textoff"):  # perceptor/losses/clip/clip.py:57-58
        losses.CLIP("ViT-B-16").add_text_off_(path=str(stub))
    oc = losses.OpenCLIP("ViT-B-32", "laion2b_s34b_b79k")
    raw = torch.randn(2, 512) * 3
    oc.add_encodings_(raw)
    assert torch.allclose(oc.encodings, raw)  # OpenCLIP stores targets as given


def test_schedule_ts_and_glue_host_logic_match_reference_golden():
    """schedule_ts is host arithmetic: bit-equal to the reference's VelocityDiffusion.schedule_ts output."""
    from pathlib import Path

    from perceptor_b200 import velocity_diffusion as vd

    z = np.load(Path(__file__).parent / "golden" / "diffusion_glue.npz")
    assert np.array_equal(vd.schedule_ts(50).numpy(), z["schedule_50"])
    assert np.array_equal(vd.schedule_ts(7, 0.9, 0.05, 5.0).numpy(), z["schedule_7"])
    rows = vd.schedule_ts(n_steps=50)
    assert rows.shape == (50, 2) and bool((rows[:, 0] > rows[:, 1]).all()) and torch.equal(rows[1:, 0], rows[:-1, 1])
    # no CPU fallback for the fused maps
    p = vd.Predictions(torch.rand(1, 3, 4, 4), torch.tensor([0.5]), torch.randn(1, 3, 4, 4))
    with pytest.raises(RuntimeError):
        _ = p.denoised_images
    with pytest.raises(ValueError):
        p.alphas(torch.zeros(2, 2))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the native one) on a one-cutout sample: one JSON
    line with the contract's keys, no GPU needed."""
    import json
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "1", "--workload", "vit_b32_224_16cut_256px"], capture_output=True, text=True,
                         timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cutouts/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["gpu_launches"] == 0
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 == d["e2e"]["d2h_bytes_per_step"]
    assert d["config"]["workload"] == "vit_b32_224_16cut_256px" and d["metric"].startswith("CLIP-guidance cutouts/sec")


def test_validate_rows_rejects_out_of_bounds():
    """Caller-supplied cutout rows index the image (and its gradient) on the device: anything outside must raise."""
    from perceptor_b200 import cutouts
    ok = np.array([[0, 0, 0, 32], [1, 8, 16, 24]], dtype=np.int32)
    assert cutouts.validate_rows(ok, 2, 32, 40).dtype == np.int32
    assert cutouts.validate_rows(np.array([[0, 0, 0, 32, 40]]), 1, 32, 40).shape == (1, 5)
    assert cutouts.validate_rows(np.zeros((0, 4), dtype=np.int64), 1, 8, 8).shape == (0, 4)
    for bad in ([[2, 0, 0, 8]], [[-1, 0, 0, 8]], [[0, -1, 0, 8]], [[0, 0, -3, 8]], [[0, 0, 0, 0]], [[0, 25, 0, 8]],
                [[0, 0, 33, 8]], [[0, 0, 0, 33]], [[0, 0, 0, 8, 41]], [[0, 30, 0, 3, 8]]):
        with pytest.raises(ValueError):
            cutouts.validate_rows(np.array(bad), 2, 32, 40)
    with pytest.raises(ValueError):
        cutouts.validate_rows(np.zeros((2, 3), dtype=np.int32), 1, 8, 8)
    with pytest.raises(ValueError):
        cutouts.validate_rows(np.zeros((2, 4), dtype=np.float32), 1, 8, 8)


def test_local_rows_rebases_image_index():
    from perceptor_b200 import cutouts
    rows = cutouts.sample_cutouts(torch.Generator().manual_seed(3), 4, 64, 64, 5, 1.0, 16, 64)
    for rank in range(4):
        loc = cutouts.local_rows(rows, rank, 4, b_offset=rank)
        assert loc.shape == (5, 4) and (loc[:, 0] == 0).all()
        assert np.array_equal(loc[:, 1:], rows[rank * 5:(rank + 1) * 5, 1:])
    assert np.array_equal(cutouts.local_rows(rows, 1, 2), rows[10:])
