"""Full-size parity (BASELINE.json configs[1] and [2]: the 7B inverse renderer and the 7B forward renderer — 136 condition
channels, K = 612 -> 616 patch GEMM, no context embedding — on a 57x704x1280 clip, S = 28 160 tokens) against the oracle run
in bf16 on the same GPU with the very same weight tensors: teacher-forced latent after one Euler step, relative L2
<= 1e-2 (north_star), at the first, a middle and the last sigma of the 15-step schedule.  The raw network output F is
printed but not gated: at 28 blocks the reference's own bf16-vs-fp32 noise on F is 1.2e-2..1.7e-2.
Also the size-independent properties the sampler offers at this size: determinism (bit-identical repeat) and linearity of
the Euler update in F."""
import pytest
import torch

from oracle import sampler_oracle as so
from oracle.dit_oracle import dit_forward
from oracle.weights import FULL_FORWARD, FULL_INVERSE
from tests.util import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", params=["inverse", "forward"])
def full_model(request):
    if torch.cuda.get_device_properties(0).total_memory < 60 * 2 ** 30:
        pytest.skip("needs ~40 GiB of device memory")
    from drb200 import diffusion_renderer_config as cfgm
    from drb200.model_diffusion_renderer import CleanDiffusionRendererModel
    cfg = cfgm.get_config_by_model_type(request.param, 704, 1280, 57)   # diffusion_renderer_config.py:131-170 / :191-233
    cfg["model_type"] = request.param
    with torch.device("meta"):
        model = CleanDiffusionRendererModel(cfg)
    model = model.to_empty(device=DEV).to(torch.bfloat16)
    model.net.init_weights_(seed=0)
    model.net._ensure_packed()                     # parameters now view the packed buffers: the oracle reads the same bytes
    model.kind = request.param
    yield model
    del model
    torch.cuda.empty_cache()


def test_full_size_euler_step_matches_oracle(full_model):
    model = full_model
    dims = FULL_INVERSE if model.kind == "inverse" else FULL_FORWARD
    sdn = {k: v for k, v in model.net.state_dict().items()}
    T, H, W = 8, 88, 160
    g = torch.Generator(device=DEV).manual_seed(1234)
    cond = (torch.randn(1, dims.additional_concat_ch, T, H, W, device=DEV, generator=g) * 0.5).bfloat16()
    if model.kind == "forward":                  # every 17th channel is a condition mask of ones (model:191-196)
        cond[:, 16::17] = 1.0
    ci = torch.full((1, 1), 3, dtype=torch.long, device=DEV)
    sig = so.sigma_schedule(15, device=DEV)
    model.scheduler.set_timesteps(15, device=DEV)
    noise = torch.randn(1, 16, T, H, W, device=DEV, generator=g).bfloat16()
    for i in (0, 7, 14):
        # a plausible x_t for this sigma: clean-latent-like signal (std 0.5) plus noise at sigma
        x_t = (cond[:, :16].float() + noise.float() * sig[i]).bfloat16()
        with torch.no_grad():
            x_in = so.scale_model_input(x_t, sig[i])
            f_ref = dit_forward(sdn, dims, x_in, sig[i], cond, ci)
            x_ref = so.euler_step(f_ref, sig[i], sig[i + 1], x_t)
            f_got = model.net(x=x_in, timesteps=sig[i], latent_condition=cond, context_index=ci)
            steps = []
            model.sample_latent(x_t, {"latent_condition": cond, "context_index": ci}, None, per_step=steps,
                                teacher=[x_t] * 15)     # teacher-forced: every step restarts from x_t; take step i
        e_x, e_f = rel_l2(steps[i], x_ref), rel_l2(f_got, f_ref)
        print(f"\n{model.kind} renderer, sigma[{i}] = {float(sig[i]):.3f}: latent after the Euler step rel-L2 {e_x:.3e}; network output F rel-L2 {e_f:.3e}")
        assert torch.isfinite(f_got.float()).all()
        assert e_x <= 1e-2


def test_full_size_forward_is_deterministic_and_euler_is_linear_in_f(full_model):
    from drb200 import ops
    model = full_model
    T, H, W = 8, 88, 160
    g = torch.Generator(device=DEV).manual_seed(7)
    x = torch.randn(1, 16, T, H, W, device=DEV, generator=g).bfloat16()
    cond = (torch.randn(1, model.net.additional_concat_ch, T, H, W, device=DEV, generator=g) * 0.5).bfloat16()
    ci = torch.full((1, 1), 1, dtype=torch.long, device=DEV)
    s = torch.tensor(1.26, device=DEV)
    with torch.no_grad():
        a = model.net(x=x, timesteps=s, latent_condition=cond, context_index=ci)
        b = model.net(x=x, timesteps=s, latent_condition=cond, context_index=ci)
    assert torch.equal(a, b)
    # Euler update: x' is affine in F (fp32 inside, ONE bf16 rounding of x'): step(x, 2F) - step(x, F) == step(x, F) - step(x, 0)
    # up to the three roundings involved: 2 bf16 ulp (ulp <= 2^-7 |x|) of the largest x' elementwise
    sn = torch.tensor(0.7, device=DEV)
    f = a[0].float()
    st = lambda ff: ops.edm_euler_step(ff.bfloat16().contiguous(), x[0].contiguous(), s.reshape(1), sn.reshape(1)).float()
    x0, x1, x2 = st(torch.zeros_like(f)), st(f), st(2 * f)
    ulp = 2.0 ** -7 * torch.maximum(torch.maximum(x0.abs(), x1.abs()), x2.abs()).clamp_min(2.0 ** -6)
    assert ((x2 - x1) - (x1 - x0)).abs().le(2.0 * ulp).all()


def test_full_size_tokenizer_matches_oracle():
    """CV8x8x8 tokenizer at the bench clip size (57 x 704 x 1280, full channel widths): encode and decode against the oracle
    in fp32 and in bf16 on the same GPU — the product may be no further from fp32 than 2x the bf16 oracle's own error."""
    if torch.cuda.get_device_properties(0).total_memory < 60 * 2 ** 30:
        pytest.skip("needs ~40 GiB of device memory")
    from oracle import vae_oracle as vo
    from drb200.CleanVAE import AutoencoderKLCosmos, CleanVAE
    sd = vo.make_vae_state_dict(vo.FULL_VAE, seed=11)
    model = AutoencoderKLCosmos()
    model.load_state_dict(sd, strict=True)
    vae = CleanVAE(model=model)
    vae.to(DEV)
    vae.reset_dtype(torch.bfloat16)
    sd16 = {k: v.to(DEV).bfloat16() for k, v in sd.items()}
    sd32 = {k: v.float() for k, v in sd16.items()}
    x = (torch.rand(1, 3, 57, 704, 1280, device=DEV, generator=torch.Generator(device=DEV).manual_seed(21)) * 2 - 1).bfloat16()
    with torch.no_grad():
        z = vae.encode(x)
        z16 = vo.encode(sd16, vo.FULL_VAE, x)
        z32 = vo.encode(sd32, vo.FULL_VAE, x.float())
        floor, err = rel_l2(z16, z32), rel_l2(z, z32)
        print(f"\nfull-size encode: product-vs-fp32 {err:.3e}   bf16-oracle-vs-fp32 {floor:.3e}")
        assert z.shape == (1, 16, 8, 88, 160)
        assert err <= 2 * floor + 2e-3
        zin = z32.bfloat16()
        del z16, z32
        torch.cuda.empty_cache()
        y = vae.decode(zin)
        y16 = vo.decode(sd16, vo.FULL_VAE, zin)
        y32 = vo.decode(sd32, vo.FULL_VAE, zin.float())
        floor, err = rel_l2(y16, y32), rel_l2(y, y32)
        print(f"full-size decode: product-vs-fp32 {err:.3e}   bf16-oracle-vs-fp32 {floor:.3e}")
        assert y.shape == (1, 3, 57, 704, 1280)
        assert err <= 2 * floor + 2e-3
