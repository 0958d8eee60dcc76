"""Self-checks of the tokenizer restatement (oracle/vae_oracle.py).  PARITY UNPINNED: diffusers' AutoencoderKLCosmos is
absent and the reference holds no fixture for it, so these are the invariants SURVEY.md §8(c) lists: parameter count,
shape contract, causality, Haar round trip, image (T = 1) path, and an independent conv-based Haar (the upstream
formulation with grouped strided convolutions) against the butterfly used by the oracle."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import vae_oracle as vo


def test_parameter_count_matches_published_size():
    n = sum(math.prod(s) for _, s in vo.vae_param_shapes(vo.FULL_VAE))
    enc = sum(math.prod(s) for k, s in vo.vae_param_shapes(vo.FULL_VAE) if k.startswith(("encoder.", "quant_conv")))
    assert 105.0e6 < n < 106.2e6, n            # ~42.9 M + ~62.7 M = ~105.6 M (NVIDIA: ~105 M)
    assert 42.5e6 < enc < 43.3e6, enc


def test_block_plan():
    e, d = vo.encoder_plan(vo.FULL_VAE), vo.decoder_plan(vo.FULL_VAE)
    assert [(b["cin"], b["cout"], b["spatial"], b["temporal"]) for b in e] == [(128, 256, True, True), (256, 512, False, False), (512, 512, False, False)]
    assert [(b["cin"], b["cout"], b["spatial"], b["temporal"]) for b in d] == [(512, 512, False, False), (512, 512, True, True), (512, 256, False, False)]


def _dwt_conv(x):
    """upstream formulation: grouped conv3d with 2-tap filters, stride 2 along T, H, W (autoencoder_kl_cosmos._dwt)"""
    g = x.shape[1]
    s = 0.7071067811865476
    hl = torch.tensor([s, s]).reshape(1, 1, 2).repeat(g, 1, 1)
    hh = torch.tensor([s, -s]).reshape(1, 1, 2).repeat(g, 1, 1)
    xl = F.conv3d(x, hl.unsqueeze(3).unsqueeze(4), groups=g, stride=(2, 1, 1))
    xh = F.conv3d(x, hh.unsqueeze(3).unsqueeze(4), groups=g, stride=(2, 1, 1))
    out = []
    for xt in (xl, xh):
        for fh in (hl, hh):
            xs = F.conv3d(xt, fh.unsqueeze(2).unsqueeze(4), groups=g, stride=(1, 2, 1))
            for fw in (hl, hh):
                out.append(F.conv3d(xs, fw.unsqueeze(2).unsqueeze(3), groups=g, stride=(1, 1, 2)))
    return torch.cat(out, dim=1) / math.sqrt(8.0)


def test_haar_matches_conv_formulation_and_round_trips():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 3, 9, 16, 24, generator=g)
    y = vo.haar_patch(x, 4)
    assert y.shape == (1, 192, 3, 4, 6)
    xp = torch.cat([x[:, :, :1].repeat_interleave(4, dim=2), x[:, :, 1:]], dim=2)
    assert torch.allclose(y, _dwt_conv(_dwt_conv(xp)), atol=1e-6)
    assert torch.allclose(vo.haar_unpatch(y, 4), x, atol=1e-5)


@pytest.mark.parametrize("frames,lat", [(1, 1), (9, 2), (17, 3)])
def test_shape_contract_and_causality(frames, lat):
    d = vo.SMALL_VAE
    sd = vo.make_vae_state_dict(d, seed=1)
    g = torch.Generator().manual_seed(frames)
    x = torch.rand(1, 3, frames, 32, 48, generator=g) * 2 - 1
    with torch.no_grad():
        z = vo.encode(sd, d, x)
        assert z.shape == (1, 16, lat, 4, 6)
        y = vo.decode(sd, d, z)
        assert y.shape == x.shape
        if frames > 1:
            # latent frame k depends only on pixel frames <= 8k: perturb the frames after 8 and compare latent frames 0..1
            x2 = x.clone()
            x2[:, :, 9:] += 1.0
            z2 = vo.encode(sd, d, x2)
            assert torch.allclose(z2[:, :, :2], z[:, :, :2], atol=1e-5)
            if frames > 9:
                assert not torch.allclose(z2[:, :, 2:], z[:, :, 2:], atol=1e-3)
            # decoder: pixel frames <= 8k depend only on latent frames <= k
            zz = z.clone()
            zz[:, :, -1] += 1.0
            y2 = vo.decode(sd, d, zz)
            keep = 1 + 8 * (lat - 2) if lat >= 2 else 0
            assert torch.allclose(y2[:, :, :keep], y[:, :, :keep], atol=1e-4)


def test_latent_frame_helpers():
    v = vo.OracleVAE({}, vo.FULL_VAE)
    assert v.get_latent_num_frames(57) == 8 and v.get_latent_num_frames(121) == 16 and v.get_latent_num_frames(1) == 1
    assert v.get_pixel_num_frames(8) == 57 and v.get_pixel_num_frames(16) == 121
    with pytest.raises(ValueError):
        v.encode(torch.zeros(3, 9, 8, 8))


def test_unpinned_oracles_have_not_drifted():
    """regression fixtures produced by the oracle itself (tests/golden/make_unpinned_regression.py): NOT upstream parity"""
    import os

    import numpy as np

    from oracle import envmap_oracle as eo
    f = np.load(os.path.join(os.path.dirname(__file__), "golden", "unpinned_regression.npz"))
    sd = vo.make_vae_state_dict(vo.SMALL_VAE, seed=7)
    with torch.no_grad():
        z = vo.encode(sd, vo.SMALL_VAE, torch.from_numpy(f["x"]))
        y = vo.decode(sd, vo.SMALL_VAE, z)
    assert np.allclose(z.numpy(), f["z"], rtol=1e-4, atol=1e-5)
    assert np.allclose(y[:, :, ::4, ::8, ::8].numpy(), f["y_sub"], rtol=1e-4, atol=1e-5)
    env = eo.render_projection_from_panorama(torch.from_numpy(f["pano"]), (12, 20), 1.3, True, 180.0, cube_res=16)
    assert np.allclose(env["env_ldr"].numpy(), f["env_ldr"], atol=1e-5) and np.allclose(env["env_log"].numpy(), f["env_log"], atol=1e-5)
