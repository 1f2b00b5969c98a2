"""Parity of the tiling / blending kernels (through the C ABI and the reference-shaped engine) with the CPU oracle and
the golden fixtures produced by the reference's own class.  Bar: bit-exact."""
import hashlib

import numpy as np
import pytest

import golden_inputs
import toy_models
from oracle import tiling as OT

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def msr(torch):
    import moonsuperresolution_b200 as m
    from moonsuperresolution_b200 import _lib
    m._lib = _lib
    return m


def engine_for(msr, case, model, **kw):
    cfg = msr.DSRConfig(image_size=case["I"], stride=case["S"], batch_size=case["B"], tile_size=case["T"],
                        no_value=case["NV"])
    return msr.DEMSuperResolution(cfg, model=model, **kw)


@pytest.mark.parametrize("name", ["wobble_200x260", "wobble_I48_S16"])
def test_pad_validity_normalize_bit_exact(msr, torch, name):
    case = golden_inputs.CASES[name]
    dem, img = golden_inputs.make_rasters(case)
    geo = OT.Geometry(case["H"], case["W"], case["I"], case["S"], case["T"])
    eng = engine_for(msr, case, toy_models.identity)
    eng.setRasters(dem, img)
    eng.padInputs()
    dem_c, img_c = OT.pad_inputs(dem, img, geo, case["NV"])
    # the engine keeps only the canvas rows its tiles read (all of them end at last tile row + T + 2 off)
    ch = eng.dem_padded.shape[0]
    assert ch == min(geo.canvas_h, OT.tile_list(geo)[-1][1] + case["T"] + 2 * geo.off)
    np.testing.assert_array_equal(eng.dem_padded.cpu().numpy(), dem_c[:ch])
    np.testing.assert_array_equal(eng.img_padded.cpu().numpy(), img_c[:ch])
    # summed-area table of the invalid mask
    inv = ((img_c[:ch] <= case["NV"]) | (dem_c[:ch] <= case["NV"])).astype(np.int64)
    sat = np.zeros((ch + 1, geo.canvas_w + 1), np.int64)
    sat[1:, 1:] = inv.cumsum(0).cumsum(1)
    np.testing.assert_array_equal(eng._sat.cpu().numpy().astype(np.int64), sat)
    # validity + normalisation of every patch of every tile
    i = case["I"]
    for (px, py) in OT.tile_list(geo):
        tp = eng._tile_plan[(px, py)]
        want = [(x, y) for (x, y) in OT.patch_origins(geo, px, py) if OT.patch_is_valid(dem_c, img_c, x, y, i, case["NV"])]
        assert [tuple(r) for r in tp["xy"]] == want
    px, py = OT.tile_list(geo)[0]
    tp = eng._tile_plan[(px, py)]
    n = min(7, tp["n_valid"])
    assert n > 0
    xy = np.full((n + 1, 2), -1, np.int32)
    xy[:n] = tp["xy"][:n]
    d_xy = torch.from_numpy(xy).cuda()
    out = torch.empty((n + 1, i, i, 2), dtype=torch.float32, device="cuda")
    mm = torch.empty((n + 1, 4), dtype=torch.float32, device="cuda")
    part = torch.empty(((n + 1) * 128,), dtype=torch.float32, device="cuda")
    eng._gather(d_xy, n + 1, out, mm, part)
    out, mm = out.cpu().numpy(), mm.cpu().numpy()
    for k in range(n):
        x, y = xy[k]
        want, (lo, hi) = OT.normalize_patch(img_c[y:y + i, x:x + i], dem_c[y:y + i, x:x + i])
        np.testing.assert_array_equal(out[k], want)
        assert mm[k, 2] == lo and mm[k, 3] == hi
    assert (out[n] == 0).all()                      # padding slot: all-zero input (process_full_tiles.py:472)
    # reference-shaped per-patch helpers
    x, y = (int(v) for v in xy[0])
    valid, ip, dp = eng.getPatch(x, y)
    assert valid and np.array_equal(dp, dem_c[y:y + i, x:x + i])
    norm, lohi = eng.normalize(ip, dp)
    want, wl = OT.normalize_patch(ip, dp)
    np.testing.assert_array_equal(norm, want)
    assert tuple(lohi) == tuple(wl)


@pytest.mark.parametrize("name", list(golden_inputs.CASES))
def test_engine_bit_exact_vs_reference_golden(msr, golden, name):
    """Whole path with the toy plug-in models: pad -> validity -> gather/normalise -> model -> blend -> assemble must
    reproduce the rasters the reference's own DEMSuperResolution produced (tests/golden/make_golden.py)."""
    case = golden_inputs.CASES[name]
    dem, img = golden_inputs.make_rasters(case)
    eng = engine_for(msr, case, getattr(toy_models, case["model"]))
    mean, std, good = eng.run(dem, img)
    assert mean.shape == (case["H"], case["W"])
    assert int(good.sum()) == int(golden[f"{name}/good_count"])
    assert sha(good) == str(golden[f"{name}/sha_good"])
    if case.get("store_full"):
        np.testing.assert_array_equal(mean, golden[f"{name}/mean"])
        np.testing.assert_array_equal(std, golden[f"{name}/std"])
    assert sha(mean) == str(golden[f"{name}/sha_mean"])
    assert sha(std) == str(golden[f"{name}/sha_std"])


def test_rebuild_tile_api_matches_oracle(msr):
    """rebuildTile(generated_dems, generated_minmax) with arbitrary keys / order and mixed float32 / float64
    predictions (the scan form of msr_blend_tile)."""
    case = dict(H=300, W=300, I=32, S=8, B=4, T=128, NV=-32768.0)
    eng = engine_for(msr, case, toy_models.identity)
    geo = OT.Geometry(300, 300, 32, 8, 128)
    rng = np.random.default_rng(5)
    keys = [(int(x) * 8, int(y) * 8) for x, y in rng.integers(0, 19, (40, 2))]
    keys = list(dict.fromkeys(keys))
    gen, mm = {}, {}
    for n, k in enumerate(keys):
        a = rng.uniform(0, 1, (32, 32))
        gen[k] = a if n % 3 == 0 else a.astype(np.float32)
        lo = np.float32(rng.uniform(-50, 50))
        mm[k] = (lo, np.float32(lo + rng.uniform(1, 30)))
    mean, std, good = eng.rebuildTile(gen, mm)
    with np.errstate(all="ignore"):
        wm, ws, wg = OT.rebuild_tile(gen, mm, geo, case["NV"])
    np.testing.assert_array_equal(good, wg)
    np.testing.assert_array_equal(mean, wm)
    np.testing.assert_array_equal(std, ws)
    # empty input: nothing reconstructed
    mean, std, good = eng.rebuildTile({}, {})
    assert (good == 0).all() and (mean == case["NV"]).all() and (std == case["NV"]).all()


@pytest.mark.parametrize("world", [2, 3])
def test_band_sharding_is_bit_identical(msr, world):
    """Mode A of SURVEY.md 8e: ranks own bands of tile rows and hold only their canvas rows; stitched bands equal the
    single-process result bit for bit (ranks emulated one after the other on one GPU)."""
    case = golden_inputs.CASES["wobble_1100x1300"]
    dem, img = golden_inputs.make_rasters(case)
    model = getattr(toy_models, case["model"])
    full = engine_for(msr, case, model).run(dem, img)
    from moonsuperresolution_b200.distributed import assemble_bands
    parts = []
    for r in range(world):
        eng = engine_for(msr, case, model, rank=r, world_size=world)
        eng.setRasters(dem, img)
        eng.padInputs()
        assert eng.dem_padded.shape[0] < eng.plan.canvas_h
        eng.processTiles()
        parts.append(eng.results())
    for k in range(3):
        got = assemble_bands([(p[3], p[k]) for p in parts], case["H"], case["W"], full[k].dtype)
        np.testing.assert_array_equal(got, full[k])


def test_large_tile_size_matches_oracle(msr):
    """tile_size is a free parameter of the reference (process_full_tiles.py:58,104-106): one 2048-pixel tile covers the
    whole raster here, so no halo patch is generated twice; still bit-exact against the oracle with the same T."""
    case = dict(H=1100, W=1500, I=64, S=16, B=16, T=2048, NV=-32768.0, model="wobble", seed=7, holes=True)
    dem, img = golden_inputs.make_rasters(case)
    eng = engine_for(msr, case, toy_models.wobble)
    mean, std, good = eng.run(dem, img)
    with np.errstate(all="ignore"):
        ref = OT.process_map(dem, img, case["I"], case["S"], case["B"], case["T"], case["NV"], toy_models.wobble)
    np.testing.assert_array_equal(good, ref[2])
    np.testing.assert_array_equal(mean, ref[0])
    np.testing.assert_array_equal(std, ref[1])


@pytest.mark.parametrize("case", [
    dict(H=40, W=50, I=64, S=8, B=4, T=256, holes=False),        # raster smaller than a patch: nothing is reconstructed
    dict(H=130, W=70, I=64, S=16, B=1, T=128, holes=True),       # batch of one, narrow raster, NV stripes
    dict(H=200, W=200, I=32, S=32, B=3, T=128, holes=False),     # stride == image_size: no overlap at all
    dict(H=90, W=300, I=16, S=4, B=16, T=64, holes=True),        # smallest legal image_size (purge = 1)
])
def test_edge_geometries_match_oracle(msr, case):
    """Ragged / degenerate inputs the reference handles implicitly: empty patch sets, B = 1, S = I, I = 16, NV stripes."""
    c = dict(case, NV=-32768.0, model="wobble", seed=11)
    dem, img = golden_inputs.make_rasters(c)
    eng = engine_for(msr, c, toy_models.wobble)
    mean, std, good = eng.run(dem, img)
    with np.errstate(all="ignore"):
        ref = OT.process_map(dem, img, c["I"], c["S"], c["B"], c["T"], c["NV"], toy_models.wobble)
    np.testing.assert_array_equal(good, ref[2])
    np.testing.assert_array_equal(mean, ref[0])
    np.testing.assert_array_equal(std, ref[1])
    if c["H"] < c["I"]:
        assert good.sum() == 0 and (mean == c["NV"]).all()


def test_all_no_value_raster_and_bad_parameters(msr):
    case = dict(H=100, W=120, I=32, S=8, B=4, T=128, NV=-32768.0)
    dem = np.full((100, 120), -32768.0, np.float32)
    eng = engine_for(msr, case, toy_models.identity)
    mean, std, good = eng.run(dem, dem.copy())
    assert good.sum() == 0 and (mean == -32768.0).all() and (std == -32768.0).all()
    # parameter combinations that crash the reference half-way are rejected up front
    bad = msr.DEMSuperResolution(msr.DSRConfig(image_size=256, stride=224, tile_size=1024, batch_size=2), model=toy_models.identity)
    bad.setRasters(np.ones((300, 300), np.float32), np.ones((300, 300), np.float32))
    with pytest.raises(ValueError):
        bad.padInputs()
    with pytest.raises(ValueError):
        engine_for(msr, case, toy_models.identity).padInputs()          # no rasters loaded
    cfg = msr.DSRConfig(source_folder_path="/nonexistent", map_name="m", save_path="/tmp")
    with pytest.raises(ValueError):
        msr.DEMSuperResolution(cfg).loadImages()                         # process_full_tiles.py:167-170
    with pytest.raises(AssertionError):
        msr.load_GAN_model("/nonexistent/", 64, 2)                       # process_full_tiles.py:27


@pytest.mark.parametrize("name", ["wobble_200x260", "wobble_1100x1300", "identity_700x900"])
def test_device_model_blend_path_is_bit_exact(msr, name):
    """Device models keep float32 predictions on the GPU (no host round trip, `+ 0.5` applied inside msr_blend_tile);
    with the device identity model the result must equal the oracle driven by a float32 identity, bit for bit."""
    case = golden_inputs.CASES[name]
    dem, img = golden_inputs.make_rasters(case)
    eng = engine_for(msr, case, msr.IdentityModel(case["I"], case["B"]))
    mean, std, good = eng.run(dem, img)
    f32_identity = lambda x, training=False: np.asarray(x, dtype=np.float32)
    with np.errstate(all="ignore"):
        ref = OT.process_map(dem, img, case["I"], case["S"], case["B"], case["T"], case["NV"], f32_identity)
    np.testing.assert_array_equal(good, ref[2])
    np.testing.assert_array_equal(mean, ref[0])
    np.testing.assert_array_equal(std, ref[1])


def test_full_size_identity_round_trip(msr, torch):
    """BASELINE.json configs[2] geometry (8192 x 8192, I = 512, S = 128, B = 16, T = 1024) with the reference's identity
    model: size-independent properties instead of an element-wise oracle (which would take hours on the CPU):
    the blended mean reproduces the DEM wherever good == 1, std ~ 0, `good` is exactly the rectangle
    [purge, last patch end - purge), everything else is no_value."""
    h = w = 8192
    i, s_, b, t = 512, 128, 16, 1024
    gen = torch.Generator(device="cuda").manual_seed(3)
    dem = (torch.cumsum(torch.randn((h, w), generator=gen, device="cuda"), 1) * 3.0 + 1500.0).contiguous()
    img = (torch.rand((h, w), generator=gen, device="cuda") * 254.0 + 1.0).contiguous()
    cfg = msr.DSRConfig(image_size=i, stride=s_, batch_size=b, tile_size=t)
    eng = msr.DEMSuperResolution(cfg, model=msr.IdentityModel(i, b))
    eng.setRasters(dem, img)
    eng.padInputs()
    eng.processTiles()
    assert eng.slots_executed == 7168                       # 6724 valid patches, 7168 slots
    good = eng.good_out.bool()
    p = i // 16
    last = ((h - i) // s_) * s_ + i - p
    want = torch.zeros((h, w), dtype=torch.bool, device="cuda")
    want[p:last, p:last] = True
    assert torch.equal(good, want)
    err = (eng.mean_out - dem).abs()[good].max().item()
    assert err <= 2e-3, err                                  # float32 rounding of (v - lo) / (hi - lo) * (hi - lo) + lo
    assert eng.std_out[good].max().item() <= 2e-3 and eng.std_out[good].min().item() >= 0.0
    assert (eng.mean_out[~good] == cfg.no_value).all() and (eng.std_out[~good] == cfg.no_value).all()


def test_save_tiles_and_geotiff_layout(msr, tmp_path):
    """Output layout of saveTile / saveGTiff (process_full_tiles.py:416-429, 481-531): names, dtypes, NoData."""
    from moonsuperresolution_b200 import geotiff
    case = golden_inputs.CASES["wobble_200x260"]
    dem, img = golden_inputs.make_rasters(case)
    geotiff.write(str(tmp_path / "run-DEM.tif"), dem)
    geotiff.write(str(tmp_path / "run-DRG.tif"), img)
    cfg = msr.DSRConfig(image_size=case["I"], stride=case["S"], batch_size=case["B"], tile_size=case["T"],
                        no_value=case["NV"], map_name="m", save_path=str(tmp_path / "out"),
                        source_folder_path=str(tmp_path), save_tiles=True, preprocess=False)
    (tmp_path / "out").mkdir()
    eng = msr.DEMSuperResolution(cfg, model=toy_models.wobble)
    eng.processMap()
    mean, _ = geotiff.read(str(tmp_path / "out" / "m_mean.tiff"))
    good, _ = geotiff.read(str(tmp_path / "out" / "m_good.tiff"))
    assert mean.dtype == np.float32 and good.dtype == np.uint16 and mean.shape == dem.shape
    ref = OT.process_map(dem, img, case["I"], case["S"], case["B"], case["T"], case["NV"], toy_models.wobble)
    np.testing.assert_array_equal(mean, ref[0])
    np.testing.assert_array_equal(good, ref[2].astype(np.uint16))
    tile, _ = geotiff.read(str(tmp_path / "out" / "tile_128_0" / "tile_128_0_mean.tif"))
    assert tile.shape == (case["T"], case["T"])
    np.testing.assert_array_equal(tile, ref[0][:case["T"], 128:128 + case["T"]])


@pytest.mark.parametrize("mode", ["faithful", "dedup"])
def test_fast_blend_matches_exact_blend(msr, mode):
    """DSRConfig(blend="fast"): float32 update, four pixels per thread, 128-bit accesses (msr_blend_tile_fast /
    msr_blend_accumulate_fast).  Same predictions (the generator is deterministic), same placement: `good` must be
    identical and mean / std within float32 rounding of the bit-exact kernels (process_full_tiles.py:395-413)."""
    from moonsuperresolution_b200 import weights as W
    i, s, b, t = 64, 16, 4, 256
    rng = np.random.default_rng(3)
    h, w_ = 300, 404
    dem = np.cumsum(np.cumsum(rng.standard_normal((h, w_)), 0), 1).astype(np.float32)
    img = rng.uniform(1, 255, (h, w_)).astype(np.float32)
    dem[100:104, 200:207] = -32768.0
    weights = W.random_init("cnn", i, seed=2, perturb_affine=True)
    model = msr.CNNSpade(i, b, precision="fp32", weights=weights, max_groups=4)
    out = {}
    for blend in ("exact", "fast"):
        cfg = msr.DSRConfig(image_size=i, stride=s, batch_size=b, tile_size=t, mode=mode, blend=blend)
        eng = msr.DEMSuperResolution(cfg, model=model)
        eng.setRasters(dem, img)
        eng.padInputs()
        assert eng._fast_blend_ok() == (blend == "fast")
        eng.processTiles()
        out[blend] = tuple(np.array(a) for a in eng.results()[:3])
    (me, se, ge), (mf, sf, gf) = out["exact"], out["fast"]
    np.testing.assert_array_equal(ge, gf)
    g = ge.astype(bool)
    assert g.sum() > 10000 and se[g].max() > 0
    scale = float(dem[dem > -32768].max() - dem[dem > -32768].min())
    e_mean = np.abs(mf[g] - me[g]).max() / scale
    e_std = np.abs(sf[g] - se[g]).max() / scale
    print(f"fast blend vs exact ({mode}): mean {e_mean:.3g}, std {e_std:.3g} of the DEM range")
    assert e_mean <= 2e-6 and e_std <= 2e-6
    assert (mf[~g] == -32768.0).all() and (sf[~g] == -32768.0).all()


def test_bands_taller_than_65535_rows(msr, torch):
    """A single-GPU dedup run of the reference's largest raster (70000 rows) finalises its whole band in one launch:
    the blend kernels index rows through a flattened 1-D grid (gridDim.y stops at 65535)."""
    i, s, b, t = 64, 32, 16, 1024
    h, w_ = 66000, 96
    y = np.arange(h, dtype=np.float32)[:, None]
    x = np.arange(w_, dtype=np.float32)[None, :]
    dem = (np.sin(y / 37.0) * 50.0 + x * 0.25 + (y % 13) * 0.5).astype(np.float32)
    img = (1.0 + (x * 7 + y * 3) % 250).astype(np.float32)
    res = {}
    for mode in ("faithful", "dedup"):
        cfg = msr.DSRConfig(image_size=i, stride=s, batch_size=b, tile_size=t, mode=mode)
        eng = msr.DEMSuperResolution(cfg, model=msr.IdentityModel(i, b))
        res[mode] = eng.run(dem, img)
    for a, r in zip(res["dedup"], res["faithful"]):
        np.testing.assert_array_equal(a, r)
    good = res["dedup"][2].astype(bool)
    assert good[65900, 40] and not good[65990, 40] and good.sum() > 60000 * 80   # last patch ends at row 65984, purge 4
    # the finalize entry point on its own, rows > 65535
    rows, cols = 70000, 8
    acc = torch.rand((3, rows, cols), device="cuda") + 0.5
    mean = torch.empty((rows, cols), device="cuda")
    std = torch.empty_like(mean)
    gd = torch.empty((rows, cols), dtype=torch.uint8, device="cuda")
    msr._lib.check(msr._lib.lib().msr_blend_finalize(acc[0].data_ptr(), acc[1].data_ptr(), acc[2].data_ptr(), cols, rows,
                                                     cols, -32768.0, mean.data_ptr(), std.data_ptr(), gd.data_ptr(), cols,
                                                     msr._lib.stream_ptr()), "msr_blend_finalize")
    torch.cuda.synchronize()
    assert torch.equal(mean, acc[1]) and bool(gd.all())
    torch.testing.assert_close(std, torch.sqrt(acc[2] / acc[0]), rtol=1e-6, atol=0)


def test_maximum_size_identity_round_trip(msr, torch):
    """BASELINE.json configs[4] geometry -- the largest raster the reference is said to handle, 15000 x 70000 (README.md:13)
    with I = 512, S = 128, B = 16, T = 1024 -- on ONE GPU with the identity model, tile by tile (1035 tiles, 123 792
    slots) and in dedup mode (61 902 positions): the two rasters must be bit-identical, `good` is exactly
    the rectangle the patch lattice covers, the mean reproduces the DEM there and everything else is no_value."""
    h, w = 70000, 15000
    i, s_, b, t = 512, 128, 16, 1024
    gen = torch.Generator(device="cuda").manual_seed(5)
    dem = (torch.cumsum(torch.randn((h, w), generator=gen, device="cuda"), 1) * 2.0 + 1200.0).contiguous()
    img = (torch.rand((h, w), generator=gen, device="cuda") * 254.0 + 1.0).contiguous()
    res = {}
    for mode in ("faithful", "dedup"):
        cfg = msr.DSRConfig(image_size=i, stride=s_, batch_size=b, tile_size=t, mode=mode)
        eng = msr.DEMSuperResolution(cfg, model=msr.IdentityModel(i, b))
        eng.setRasters(dem, img)
        eng.padInputs()
        eng.processTiles()
        res[mode] = (eng.mean_out, eng.std_out, eng.good_out, eng.slots_executed)
        del eng
    assert res["faithful"][3] == 123792 and res["dedup"][3] == 61904        # 61 902 positions in whole batches of 16
    for k in range(3):
        assert torch.equal(res["faithful"][k], res["dedup"][k])
    mean, std, good = res["dedup"][0], res["dedup"][1], res["dedup"][2].bool()
    p = i // 16
    want = torch.zeros((h, w), dtype=torch.bool, device="cuda")
    want[p:((h - i) // s_) * s_ + i - p, p:((w - i) // s_) * s_ + i - p] = True
    assert torch.equal(good, want)
    assert (mean - dem).abs()[good].max().item() <= 2e-3
    assert 0.0 <= std[good].min().item() and std[good].max().item() <= 2e-3
    assert bool((mean[~good] == -32768.0).all()) and bool((std[~good] == -32768.0).all())
