"""Repeated-sample mode (DSRConfig.samples_per_patch = R > 1; SURVEY.md section 8f row 4, beyond the reference): every batch
is generated R times and blended batch by batch, repetition by repetition, patch by patch through msr_blend_accumulate /
msr_blend_finalize.  Parity is against the oracle run with the same plan (oracle/tiling.py, ``repeats``): bit-exact for
host plug-in models, within the generator tolerances for the CUDA generators."""
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


def cfg_for(msr, case, **kw):
    return msr.DSRConfig(image_size=case["I"], stride=case["S"], batch_size=case["B"], tile_size=case["T"],
                         no_value=case["NV"], **kw)


@pytest.mark.parametrize("name,repeats", [("wobble_200x260", 3), ("wobble_1100x1300", 2)])
@pytest.mark.parametrize("mode", ["faithful", "dedup"])
def test_repeats_bit_exact_against_the_oracle(msr, name, repeats, mode):
    """Stateful host plug-in (every call differs, like fresh sampler noise): same call sequence, same blend order."""
    case = golden_inputs.CASES[name]
    dem, img = golden_inputs.make_rasters(case)
    eng = msr.DEMSuperResolution(cfg_for(msr, case, mode=mode, samples_per_patch=repeats), model=toy_models.Flicker())
    mean, std, good = eng.run(dem, img)
    fn = OT.process_map if mode == "faithful" else OT.process_map_dedup
    with np.errstate(all="ignore"):
        ref = fn(dem, img, case["I"], case["S"], case["B"], case["T"], case["NV"], toy_models.Flicker(), repeats=repeats)
    np.testing.assert_array_equal(good, ref[2])
    np.testing.assert_array_equal(mean, ref[0])
    np.testing.assert_array_equal(std, ref[1])
    assert std[good.astype(bool)].max() > 0


def test_accumulate_path_with_one_repeat_equals_the_tile_kernel(msr):
    """The repeats path with R = 1 (accumulate + finalize per tile) and the gather-form msr_blend_tile implement the
    same loop: bit-identical rasters, and equal to the reference's golden output for a per-sample model."""
    case = golden_inputs.CASES["wobble_200x260"]
    dem, img = golden_inputs.make_rasters(case)
    normal = msr.DEMSuperResolution(cfg_for(msr, case), model=toy_models.ripple).run(dem, img)
    normal = tuple(np.array(a) for a in normal)
    eng = msr.DEMSuperResolution(cfg_for(msr, case), model=toy_models.ripple)
    eng.setRasters(dem, img)
    eng.padInputs()
    for tile in eng.my_tiles:
        eng._process_tile_repeats(*tile)
    got = eng.results()[:3]
    for a, b in zip(got, normal):
        np.testing.assert_array_equal(a, b)


def test_repeats_with_cuda_generators(msr):
    """CNN-SPADE (deterministic: the R generations coincide) and GauGAN with explicit sampler noise (R, slots, 256) per
    tile against the oracle generator driven through the oracle pipeline with the same repeats."""
    from moonsuperresolution_b200 import weights as W
    from oracle import generator as OG
    i, s, b, t, repeats = 64, 32, 4, 128, 2
    rng = np.random.default_rng(0)
    h, w_ = 150, 170
    dem = np.cumsum(np.cumsum(rng.standard_normal((h, w_)), 0), 1).astype(np.float32)
    img = rng.uniform(1, 255, (h, w_)).astype(np.float32)
    scale = float(dem.max() - dem.min())
    cfg = msr.DSRConfig(image_size=i, stride=s, batch_size=b, tile_size=t, samples_per_patch=repeats)
    # deterministic generator
    weights = W.random_init("cnn", i, seed=1, perturb_affine=True)
    ref = OT.process_map(dem, img, i, s, b, t, cfg.no_value, OG.OracleModel("cnn", weights), repeats=repeats)
    eng = msr.DEMSuperResolution(cfg, model=msr.CNNSpade(i, b, precision="fp32", weights=weights, max_groups=3))
    mean, std, good = eng.run(dem, img)
    np.testing.assert_array_equal(good, ref[2])
    g = good.astype(bool)
    assert np.abs(mean[g] - ref[0][g]).max() / scale <= 4e-4 and np.abs(std[g] - ref[1][g]).max() / scale <= 4e-4
    # stochastic generator: one noise tensor per tile, consumed batch by batch and repetition by repetition
    weights = W.random_init("spade", i, seed=2, perturb_affine=True)
    eng = msr.DEMSuperResolution(cfg, model=msr.GauGAN(i, b, precision="fp32", weights=weights, max_groups=3))
    eng.setRasters(dem, img)
    eng.padInputs()
    calls = []                                                      # noise of every oracle call, in call order
    for (px, py) in eng.my_tiles:
        slots = eng.plan.batch_slots(eng._tile_plan[(px, py)]["n_valid"])
        eps = rng.standard_normal((repeats, slots, 256)).astype(np.float32)
        eng.processTile(px, py, eps=eps if slots else None)
        for j in range(0, slots, b):
            for r in range(repeats):
                calls.append(eps[r, j:j + b])
    mean, std, good = eng.results()[:3]
    oracle_model = OG.OracleModel("spade", weights, eps_fn=lambda c, n: calls[c])
    ref = OT.process_map(dem, img, i, s, b, t, cfg.no_value, oracle_model, repeats=repeats)
    assert oracle_model.calls == len(calls)
    np.testing.assert_array_equal(good, ref[2])
    g = good.astype(bool)
    assert np.abs(mean[g] - ref[0][g]).max() / scale <= 4e-4 and np.abs(std[g] - ref[1][g]).max() / scale <= 4e-4
    assert std[g].max() / scale > 1e-4                              # the repetitions do disagree


@pytest.mark.parametrize("world", [2, 3])
def test_dedup_repeats_across_ranks(msr, world):
    """Seam rows with repeats: the kept (batch, repetition) predictions are replayed in order on the neighbour's strip.
    With R > 1 the blend order follows the batches, and batches never span ranks, so the yardstick is the oracle run
    with the same bands of lattice rows (bit-exact), not the single-rank rasters (which differ at rounding level)."""
    from moonsuperresolution_b200.distributed import assemble_bands
    case = golden_inputs.CASES["wobble_200x260"]
    dem, img = golden_inputs.make_rasters(case)
    cfg = cfg_for(msr, case, mode="dedup", samples_per_patch=2)
    engines = []
    for r in range(world):
        eng = msr.DEMSuperResolution(cfg, model=toy_models.ripple, rank=r, world_size=world)
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
        parts.append(tuple(np.array(a) if isinstance(a, np.ndarray) else a for a in eng.results()))
    bands = [(e._dband.j0, e._dband.j1) for e in engines]
    with np.errstate(all="ignore"):
        ref = OT.process_map_dedup(dem, img, case["I"], case["S"], case["B"], case["T"], case["NV"], toy_models.ripple,
                                   row_bands=bands, repeats=2)
    single = msr.DEMSuperResolution(cfg, model=toy_models.ripple).run(dem, img)
    for k in range(3):
        got = assemble_bands([(p[3], p[k]) for p in parts], case["H"], case["W"], ref[k].dtype)
        np.testing.assert_array_equal(got, ref[k])
    g = ref[2].astype(bool)
    np.testing.assert_array_equal(single[2], ref[2])
    assert np.abs(single[0][g] - ref[0][g]).max() <= 1e-5 * np.abs(ref[0][g]).max()


def test_spade_reuse_across_generations(msr, torch):
    """Repeated-sample mode with the bf16 GauGAN: the first generation of a batch stores the encoder's mean | variance
    and gamma | beta of all 15 SPADE layers (spade.py:18-20 depend only on the patch), the next ones reuse them
    (msr_generator_forward_repeat).  FIRST and NEXT are the same arithmetic, so identical noise gives identical
    output; against the plain forward the only difference is the bf16 rounding of the stored gamma | beta, which must
    stay far inside north_star's bf16 bar when measured against the oracle."""
    from moonsuperresolution_b200 import _lib
    from moonsuperresolution_b200 import weights as W
    from oracle import generator as OG
    i, b, groups = 128, 4, 2
    w = W.random_init("spade", i, seed=31, perturb_affine=True)
    rng = np.random.default_rng(2)
    x = rng.uniform(-0.5, 0.5, (groups * b, i, i, 2)).astype(np.float32)
    x[-1] = 0.0
    eps = rng.standard_normal((3, groups * b, 256)).astype(np.float32)
    model = msr.GauGAN(i, b, precision="bf16", weights=w, max_groups=groups)
    src = torch.from_numpy(x).cuda()
    d_eps = torch.from_numpy(eps).cuda()
    out = {k: torch.empty((groups * b, i, i), device="cuda") for k in ("plain0", "first0", "next0", "next1", "plain1")}
    model.forward_device(src, out["plain0"], d_eps[0], groups)
    model.forward_device(src, out["first0"], d_eps[0], groups, repeat_phase=_lib.REPEAT_FIRST)
    n_first = model.last_launch_count
    model.forward_device(src, out["next1"], d_eps[1], groups, repeat_phase=_lib.REPEAT_NEXT)
    n_next = model.last_launch_count
    model.forward_device(src, out["next0"], d_eps[0], groups, repeat_phase=_lib.REPEAT_NEXT)
    model.forward_device(src, out["plain1"], d_eps[1], groups)
    torch.cuda.synchronize()
    assert torch.equal(out["first0"], out["next0"])                    # reuse changes nothing
    assert n_next < n_first - 30                                        # encoder + 15 x (mask conv, gamma | beta conv) gone
    for g in range(groups):
        sl = slice(g * b, (g + 1) * b)
        for k, e in (("first0", 0), ("next1", 1)):
            want = OG.gaugan_call(x[sl], w, eps[e, sl], "spade")[..., 0]
            err = np.abs(out[k][sl].cpu().numpy() - want).max() / max(1.0, np.abs(want).max())
            assert err <= 1e-2, (k, g, err)
    assert (out["plain0"] - out["first0"]).abs().max().item() <= 1e-2
    assert (out["plain1"] - out["next1"]).abs().max().item() <= 1e-2
    # a NEXT call without its FIRST is refused
    model.forward_device(src, out["plain0"], d_eps[0], groups)
    with pytest.raises(_lib.MoonSRError):
        model.forward_device(src, out["next0"], d_eps[0], groups, repeat_phase=_lib.REPEAT_NEXT)


def test_repeated_sample_engine_with_and_without_reuse(msr):
    """DSRConfig(samples_per_patch=3) with the bf16 GauGAN, reuse on (default) and off: same `good`, mean / std within
    the bf16 budget of each other (the stored gamma | beta are bf16), and fewer generator launches with reuse."""
    from moonsuperresolution_b200 import weights as W
    i, s, b, t = 64, 32, 4, 128
    rng = np.random.default_rng(5)
    h, w_ = 150, 170
    dem = np.cumsum(np.cumsum(rng.standard_normal((h, w_)), 0), 1).astype(np.float32)
    img = rng.uniform(1, 255, (h, w_)).astype(np.float32)
    weights = W.random_init("spade", i, seed=3, perturb_affine=True)
    model = msr.GauGAN(i, b, precision="bf16", weights=weights, max_groups=2)
    res, launches = {}, {}
    for reuse in (True, False):
        cfg = msr.DSRConfig(image_size=i, stride=s, batch_size=b, tile_size=t, samples_per_patch=3, seed=9,
                            reuse_spade=reuse)
        eng = msr.DEMSuperResolution(cfg, model=model)
        res[reuse] = eng.run(dem, img)
        launches[reuse] = eng.model_launches
    np.testing.assert_array_equal(res[True][2], res[False][2])
    g = res[True][2].astype(bool)
    scale = float(dem.max() - dem.min())
    assert np.abs(res[True][0][g] - res[False][0][g]).max() / scale <= 1e-2
    assert np.abs(res[True][1][g] - res[False][1][g]).max() / scale <= 1e-2
    assert res[True][1][g].max() > 0 and launches[True] < launches[False]
