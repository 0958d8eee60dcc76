"""Pin the oracle against the REAL reference modules (bit-exact, fp32 and bf16, CPU).
Skipped where /root/reference is absent (the GPU box); the golden-vector test covers that case."""
import pytest
import torch

from oracle import ref_loader
from oracle.dit_oracle import dit_forward
from oracle.weights import MICRO_FORWARD, MICRO_INVERSE, make_state_dict, net_only

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.mark.parametrize("dims", [MICRO_INVERSE, MICRO_FORWARD])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dit_bit_exact(dims, dtype):
    dit, _, _, _ = ref_loader.load()
    sd = make_state_dict(dims, seed=7, dtype=dtype)
    net = dit.CleanDiffusionRendererGeneralDIT(**dims.net_kwargs())
    net.load_state_dict(net_only(sd), strict=True)
    net = net.to(dtype).eval()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 16, 3, 8, 12, generator=g).to(dtype)
    cond = torch.randn(1, dims.additional_concat_ch, 3, 8, 12, generator=g).to(dtype)
    ci = torch.full((1, 1), 2, dtype=torch.long).to(dtype)       # the pipeline casts it (pipeline:200-208)
    for sigma in (80.0, 0.7, 0.02):
        with torch.no_grad():
            ref = net(x=x, timesteps=torch.tensor(sigma), latent_condition=cond, context_index=ci)
            mine = dit_forward(net_only(sd), dims, x, torch.tensor(sigma), cond, ci)
        assert torch.equal(ref, mine)


def test_state_dict_keys_match_reference_model():
    _, cfgm, mdl, _ = ref_loader.load()
    for dims, getc in ((MICRO_INVERSE, cfgm.get_inverse_renderer_config), (MICRO_FORWARD, cfgm.get_forward_renderer_config)):
        config = getc(64, 96, 9)
        config["net"].update(model_channels=dims.model_channels, num_blocks=dims.num_blocks, num_heads=dims.num_heads)
        model = mdl.CleanDiffusionRendererModel(config)
        ref = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        mine = {k: tuple(v.shape) for k, v in make_state_dict(dims, seed=0).items()}
        assert ref == mine


def test_envmap_oracle_matches_reference_torch_stages():
    """apply_hdr_preprocessing, latlong_to_cubemap_official, latlong_vec, hdr_mapping_official: bit-exact on CPU"""
    from oracle import envmap_oracle as eo
    ref = ref_loader.load_envmap()
    g = torch.Generator().manual_seed(3)
    pano = torch.rand(24, 48, 3, generator=g) * 8.0
    pano[2, 5, 1] = float("nan")
    pano[3, 7, 0] = float("inf")
    for bright, flip, rot in ((1.0, True, 180.0), (1.7, False, 37.0), (0.5, True, 0.0)):
        a = ref.apply_hdr_preprocessing(pano.clone(), bright, flip, rot, "cpu")
        b = eo.apply_hdr_preprocessing(pano, bright, flip, rot)
        assert torch.equal(a, b)
        assert torch.equal(ref.latlong_to_cubemap_official(a, [16, 16]), eo.latlong_to_cubemap(b, [16, 16]))
    assert torch.equal(ref.latlong_vec((12, 20), device="cpu"), eo.latlong_vec((12, 20)))
    x = torch.rand(9, 11, 3, generator=g) * 30
    r, o = ref.hdr_mapping_official(x), eo.hdr_mapping(x)
    assert torch.equal(r["env_ev0"], o["env_ev0"]) and torch.equal(r["env_log"], o["env_log"])
    for s in range(6):
        gx, gy = torch.rand(5, generator=g) * 2 - 1, torch.rand(5, generator=g) * 2 - 1
        assert torch.equal(ref.cube_to_dir(s, gx, gy), eo.cube_to_dir(s, gx, gy))
