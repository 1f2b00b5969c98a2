"""Dedup mode (mode B): every position of the global patch lattice generated once, blended into
canvas-shaped accumulators by msr_blend_accumulate / msr_blend_finalize.

Bars: for a model whose output does not depend on batch composition the result must equal the REFERENCE's tile-by-tile
result bit for bit (every pixel receives the same patches in the same order) -- for one rank and for any number of
ranks (the seam rows continue from the previous rank's accumulator strip, so the order is kept).  For the SPADE
generators (batch statistics) parity is against the oracle run with the same batch plan, within the bf16 / fp32
tolerance of BASELINE.json."""
import numpy as np
import pytest

import golden_inputs
import toy_models
from oracle import tiling as OT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def msr(torch):
    import moonsuperresolution_b200 as m
    return m


def engine_for(msr, case, model, **kw):
    cfg = msr.DSRConfig(image_size=case["I"], stride=case["S"], batch_size=case["B"], tile_size=case["T"],
                        no_value=case["NV"], mode="dedup")
    return msr.DEMSuperResolution(cfg, model=model, **kw)


def f32_identity(x, training=False):
    return np.asarray(x, dtype=np.float32)


CASES = {
    "wobble_200x260": golden_inputs.CASES["wobble_200x260"],
    "wobble_1100x1300": golden_inputs.CASES["wobble_1100x1300"],
    "identity_700x900": golden_inputs.CASES["identity_700x900"],
    "no_overlap": dict(H=200, W=200, I=32, S=32, B=3, T=128, NV=-32768.0, seed=11, holes=False),
    "smallest_I": dict(H=90, W=300, I=16, S=4, B=16, T=64, NV=-32768.0, seed=12, holes=True),
    "tiny": dict(H=40, W=50, I=64, S=8, B=4, T=256, NV=-32768.0, seed=13, holes=False),
}


@pytest.mark.parametrize("name", list(CASES))
def test_dedup_equals_reference_order_for_per_sample_models(msr, name):
    """Host plug-in model (per-sample): dedup result == oracle of the reference's tile-by-tile path, bit for bit."""
    case = CASES[name]
    dem, img = golden_inputs.make_rasters(case)
    eng = engine_for(msr, case, toy_models.ripple)
    mean, std, good = eng.run(dem, img)
    with np.errstate(all="ignore"):
        ref = OT.process_map(dem, img, case["I"], case["S"], case["B"], case["T"], case["NV"], toy_models.ripple)
    np.testing.assert_array_equal(good, ref[2])
    np.testing.assert_array_equal(mean, ref[0])
    np.testing.assert_array_equal(std, ref[1])


@pytest.mark.parametrize("name", ["wobble_200x260", "identity_700x900"])
def test_dedup_device_identity_model(msr, name):
    """Device model path (float32 predictions stay on the GPU, `+ 0.5` inside the kernel)."""
    case = CASES[name]
    dem, img = golden_inputs.make_rasters(case)
    eng = engine_for(msr, case, msr.IdentityModel(case["I"], case["B"]))
    mean, std, good = eng.run(dem, img)
    with np.errstate(all="ignore"):
        ref = OT.process_map(dem, img, case["I"], case["S"], case["B"], case["T"], case["NV"], f32_identity)
    np.testing.assert_array_equal(good, ref[2])
    np.testing.assert_array_equal(mean, ref[0])
    np.testing.assert_array_equal(std, ref[1])


def run_ranks(msr, case, model_fn, world, dem, img, sharded=False):
    """Ranks emulated one after the other on one GPU; the seam strips are handed over by hand (what
    DEMSuperResolution._exchange_seams does with send / recv)."""
    from moonsuperresolution_b200.distributed import assemble_bands
    engines = []
    for r in range(world):
        eng = engine_for(msr, case, model_fn(), rank=r, world_size=world)
        if sharded:
            n0, n1 = eng.rowsNeeded(case["H"], case["W"])
            eng.setRasters(dem[n0:n1], img[n0:n1], row_offset=n0, full_height=case["H"])
        else:
            eng.setRasters(dem, img)
        eng.padInputs()
        eng.processBandMain()
        engines.append(eng)
    for r in range(1, world):
        strip = engines[r - 1].seamOut()
        if strip is not None:
            engines[r].seamIn(strip.clone())
    parts = []
    for eng in engines:
        eng.processBandFinish()
        parts.append(eng.results())
        parts[-1] = tuple(np.array(a) if isinstance(a, np.ndarray) else a for a in parts[-1])
    return [assemble_bands([(p[3], p[k]) for p in parts], case["H"], case["W"], parts[0][k].dtype) for k in range(3)]


@pytest.mark.parametrize("name,world", [("wobble_200x260", 2), ("wobble_200x260", 3), ("wobble_1100x1300", 4),
                                        ("identity_700x900", 3)])
def test_dedup_is_bit_identical_for_any_rank_count(msr, name, world):
    case = CASES[name]
    dem, img = golden_inputs.make_rasters(case)
    with np.errstate(all="ignore"):
        ref = OT.process_map(dem, img, case["I"], case["S"], case["B"], case["T"], case["NV"], toy_models.ripple)
    got = run_ranks(msr, case, lambda: toy_models.ripple, world, dem, img, sharded=(world == 3))
    np.testing.assert_array_equal(got[2], ref[2])
    np.testing.assert_array_equal(got[0], ref[0])
    np.testing.assert_array_equal(got[1], ref[1])


def test_dedup_too_many_ranks_is_rejected(msr):
    case = CASES["wobble_200x260"]
    dem, img = golden_inputs.make_rasters(case)
    eng = engine_for(msr, case, toy_models.ripple, rank=0, world_size=16)
    eng.setRasters(dem, img)
    with pytest.raises(ValueError):
        eng.padInputs()
    cfg = msr.DSRConfig(image_size=24, stride=16, batch_size=2, tile_size=120, mode="dedup")   # S | T + I, not T
    eng = msr.DEMSuperResolution(cfg, model=toy_models.ripple)
    eng.setRasters(dem, img)
    with pytest.raises(ValueError):
        eng.padInputs()
    with pytest.raises(ValueError):
        msr.DEMSuperResolution(msr.DSRConfig(mode="dedup", save_tiles=True), model=toy_models.ripple)


def test_dedup_spade_generator_matches_oracle_with_same_batch_plan(msr):
    """CNN-SPADE (batch statistics): the oracle generator driven through the oracle's dedup pipeline -- same lattice
    order, same batches -- against the engine, fp32 and bf16 modes."""
    from moonsuperresolution_b200 import weights as W
    from oracle import generator as OG
    i, s, b, t = 64, 32, 4, 128
    rng = np.random.default_rng(0)
    h, w_ = 150, 300
    dem = np.cumsum(np.cumsum(rng.standard_normal((h, w_)), 0), 1).astype(np.float32)
    img = rng.uniform(1, 255, (h, w_)).astype(np.float32)
    dem[60:63, 80:90] = -32768.0
    weights = W.random_init("cnn", i, seed=1, perturb_affine=True)
    ref = OT.process_map_dedup(dem, img, i, s, b, t, -32768.0, OG.OracleModel("cnn", weights))
    scale = float(dem[dem > -32768].max() - dem[dem > -32768].min())
    for precision, tol in (("fp32", 1e-4), ("bf16", 1e-2)):
        cfg = msr.DSRConfig(image_size=i, stride=s, batch_size=b, tile_size=t, mode="dedup")
        eng = msr.DEMSuperResolution(cfg, model=msr.CNNSpade(i, b, precision=precision, weights=weights, max_groups=3))
        mean, std, good = eng.run(dem, img)
        np.testing.assert_array_equal(good, ref[2])
        g = good.astype(bool)
        assert g.any()
        assert np.abs(mean[g] - ref[0][g]).max() / scale <= tol, (precision, np.abs(mean[g] - ref[0][g]).max() / scale)
        assert np.abs(std[g] - ref[1][g]).max() / scale <= tol, (precision, np.abs(std[g] - ref[1][g]).max() / scale)
        assert (mean[~g] == cfg.no_value).all()


def test_dedup_pix2pix_equals_faithful_mode(msr):
    """pix2pix at inference is per-sample (BatchNorm folded into scale / shift), so dedup mode must reproduce the
    tile-by-tile result of the same CUDA generator: same `good`, mean / std equal to float32 rounding of the network
    (the generator's own output for a patch may depend on its slot only through the order of identical operations)."""
    from moonsuperresolution_b200 import weights as W
    i, s, b = 256, 64, 4
    rng = np.random.default_rng(4)
    h, w_ = 1200, 700
    dem = np.cumsum(np.cumsum(rng.standard_normal((h, w_)), 0), 1).astype(np.float32)
    img = rng.uniform(1, 255, (h, w_)).astype(np.float32)
    weights = W.random_init("pix2pix", i, seed=5, perturb_affine=True)
    model = msr.Pix2Pix(batch_size=b, weights=weights, max_groups=4)
    outs = {}
    for mode in ("faithful", "dedup"):
        cfg = msr.DSRConfig(image_size=i, stride=s, batch_size=b, tile_size=1024, mode=mode)
        eng = msr.DEMSuperResolution(cfg, model=model)
        outs[mode] = eng.run(dem, img)
        outs[mode] = tuple(np.array(a) for a in outs[mode])
        outs[mode + "_slots"] = eng.slots_executed
    assert outs["dedup_slots"] < outs["faithful_slots"]
    np.testing.assert_array_equal(outs["dedup"][2], outs["faithful"][2])
    np.testing.assert_array_equal(outs["dedup"][0], outs["faithful"][0])
    np.testing.assert_array_equal(outs["dedup"][1], outs["faithful"][1])


def test_dedup_full_size_identity_round_trip(msr, torch):
    """BASELINE.json configs[2] geometry in dedup mode: 3721 patches (3728 slots) instead of 6724 (7168), same
    size-independent properties as the tile-by-tile test, and bit-identical rasters to the tile-by-tile engine."""
    h = w = 8192
    i, s_, b, t = 512, 128, 16, 1024
    gen = torch.Generator(device="cuda").manual_seed(3)
    dem = (torch.cumsum(torch.randn((h, w), generator=gen, device="cuda"), 1) * 3.0 + 1500.0).contiguous()
    img = (torch.rand((h, w), generator=gen, device="cuda") * 254.0 + 1.0).contiguous()
    res = {}
    for mode in ("dedup", "faithful"):
        cfg = msr.DSRConfig(image_size=i, stride=s_, batch_size=b, tile_size=t, mode=mode)
        eng = msr.DEMSuperResolution(cfg, model=msr.IdentityModel(i, b))
        eng.setRasters(dem, img)
        eng.padInputs()
        eng.processTiles()
        res[mode] = (eng.mean_out.clone(), eng.std_out.clone(), eng.good_out.clone(), eng.slots_executed)
    assert res["dedup"][3] == 3728 and res["faithful"][3] == 7168
    for k in range(3):
        assert torch.equal(res["dedup"][k], res["faithful"][k])
    good = res["dedup"][2].bool()
    err = (res["dedup"][0] - dem).abs()[good].max().item()
    assert err <= 2e-3, err
