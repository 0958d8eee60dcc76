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
