"""CPU self-checks of the environment-map oracle (the nvdiffrast cube fetch is UNPINNED: see oracle/envmap_oracle.py)."""
import torch

from oracle import envmap_oracle as eo


def test_dir_to_cube_inverts_cube_to_dir():
    g = torch.Generator().manual_seed(0)
    x, y = torch.rand(200, generator=g) * 1.98 - 0.99, torch.rand(200, generator=g) * 1.98 - 0.99
    for s in range(6):
        s2, x2, y2 = eo.dir_to_cube(eo.cube_to_dir(s, x, y) * 3.7)
        assert (s2 == s).all() and torch.allclose(x2, x, atol=1e-6) and torch.allclose(y2, y, atol=1e-6)


def test_cube_fetch_reproduces_texel_centres_and_is_seamless():
    g = torch.Generator().manual_seed(1)
    R = 8
    cube = torch.rand(6, R, R, 3, generator=g)
    c = torch.linspace(-1 + 1 / R, 1 - 1 / R, R)
    gy, gx = torch.meshgrid(c, c, indexing="ij")
    for s in range(6):
        got = eo.cube_texture_linear(cube, eo.cube_to_dir(s, gx, gy))
        assert torch.allclose(got, cube[s], atol=1e-5)
    # a constant cube map stays constant everywhere, including across edges and corners
    const = torch.full((6, R, R, 3), 0.37)
    d = torch.randn(5000, 3, generator=g)
    assert torch.allclose(eo.cube_texture_linear(const, d), torch.full((5000, 3), 0.37), atol=1e-6)


def test_projection_of_a_constant_panorama_is_the_tone_mapped_constant():
    pano = torch.full((16, 32, 3), 2.0)
    out = eo.render_projection_from_panorama(pano, (12, 20), 1.0, True, 180.0, cube_res=16)
    want = eo.hdr_mapping(torch.full((1,), 2.0))
    assert torch.allclose(out["env_ldr"], want["env_ev0"].expand(12, 20, 3), atol=1e-5)
    assert torch.allclose(out["env_log"], want["env_log"].expand(12, 20, 3), atol=1e-5)
    assert out["env_ldr"].shape == (12, 20, 3)
