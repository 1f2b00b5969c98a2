"""Self-checks of the torch-CPU generator oracle (parity vs TensorFlow is UNPINNED, see oracle/generator.py)."""
import numpy as np
import pytest
import torch

from moonsuperresolution_b200 import weights as W
from oracle import generator as G


def test_param_counts_match_survey_appendix_a():
    assert round(W.param_count(W.spade_generator_spec(256)) / 1e6, 2) == 100.86
    assert round(W.param_count(W.spade_generator_spec(512)) / 1e6, 2) == 113.50
    assert round(W.param_count(W.encoder_spec(256)) / 1e6, 2) == 20.69
    assert round(W.param_count(W.encoder_spec(512)) / 1e6, 2) == 71.02
    # 54.4 M (App. A.3): the 3-in/3-out tutorial U-Net has 54,425,859 parameters incl. BatchNorm moving statistics;
    # 2-in / 1-out removes 4*4*64 + 4*4*128*2 + 2 of them.
    assert W.param_count(W.pix2pix_spec()) == 54425859 - 1024 - 4096 - 2


def test_same_padding_rules():
    """App. B.2: k=4,s=1 -> (1,2); k=3,s=2 even -> (0,1); k=4,s=2 -> (1,1)."""
    x = torch.arange(36, dtype=torch.float32).reshape(1, 1, 6, 6)
    assert G._same_pad(x, 4, 1).shape[-1] == 9 and G._same_pad(x, 4, 1)[0, 0, 1, 1] == 0 and G._same_pad(x, 4, 1)[0, 0, 0, 0] == 0
    p = G._same_pad(x, 3, 2)
    assert p.shape[-1] == 7 and p[0, 0, 0, 0] == 0 and p[0, 0, 0, 1] == 1 and p[0, 0, 6, 6] == 0
    assert G._same_pad(x, 4, 2).shape[-1] == 8


def test_resize_nearest_half_pixel():
    """App. B.3: power-of-two reduction 2^k picks src = dst * 2^k + 2^(k-1)."""
    m = torch.arange(64, dtype=torch.float32).reshape(1, 1, 8, 8)
    r = G.resize_nearest_tf(m, 2)
    assert r[0, 0].tolist() == [[m[0, 0, 2, 2].item(), m[0, 0, 2, 6].item()], [m[0, 0, 6, 2].item(), m[0, 0, 6, 6].item()]]
    assert torch.equal(G.resize_nearest_tf(m, 8), m)
    assert torch.equal(G.resize_nearest_tf(m, 4), torch.nn.functional.interpolate(m, size=4, mode="nearest-exact"))


@pytest.mark.parametrize("arch", ["spade", "cnn"])
def test_spade_shapes_and_fp32_fp64_agree(arch):
    i, b = 64, 3
    w = W.random_init(arch, i, seed=3, perturb_affine=True)
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.5, 0.5, (b, i, i, 2)).astype(np.float32)
    eps = rng.standard_normal((b, 256)).astype(np.float32)
    y32 = G.gaugan_call(x, w, eps, arch, torch.float32)
    y64 = G.gaugan_call(x, w, eps, arch, torch.float64)
    assert y32.shape == (b, i, i, 1)
    assert np.abs(y32 - y64).max() < 2e-4 * max(1.0, np.abs(y64).max())


def test_spade_batch_statistics_couple_samples():
    """spade.py:21 -- moments over (N, H, W): a sample's output depends on its batch-mates."""
    i = 64
    w = W.random_init("cnn", i, seed=1)
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.5, 0.5, (2, i, i, 2)).astype(np.float32)
    both = G.gaugan_call(x, w, None, "cnn")
    alone = G.gaugan_call(x[:1], w, None, "cnn")
    assert np.abs(both[0] - alone[0]).max() > 1e-3


def test_pix2pix_shape_range_and_fp64():
    w = W.random_init("pix2pix", 256, seed=2, perturb_affine=True)
    x = np.random.default_rng(0).uniform(-0.5, 0.5, (1, 256, 256, 2)).astype(np.float32)
    y32 = G.pix2pix_call(x, w, torch.float32)
    y64 = G.pix2pix_call(x, w, torch.float64)
    assert y32.shape == (1, 256, 256, 1) and np.abs(y32).max() <= 1.0
    assert np.abs(y32 - y64).max() < 1e-4
