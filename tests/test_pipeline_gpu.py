"""End-to-end GPU parity: the product pipeline / node classes (B200 tokenizer + GeneralDIT + EDM sampler + uint8
post-process) against the oracle chain (vae_oracle -> sampler_oracle -> vae_oracle -> postprocess) in bf16 on the same
GPU, same weights, same clip, same seed.

Gates (BASELINE.json north_star): final decoded frames >= 40 dB PSNR; the final latent within 1e-2 relative L2.  The
decoded video tensor is also compared before quantisation (random-init decoders give low-contrast frames, for which
PSNR alone would be a weak check)."""
import sys
import types

import numpy as np
import pytest
import torch

from oracle import sampler_oracle as so
from oracle import vae_oracle as vo
from oracle.weights import MICRO_FORWARD, MICRO_INVERSE, TINY_INVERSE, net_only
from tests.util import build_product_model, psnr_u8, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"
VDIMS = vo.SMALL_VAE


def _vae(seed=7, target_std=0.45):
    """Product tokenizer + the same weights for the oracle.  The last decoder convolution is rescaled so that decoded
    frames have real-image contrast (std ~0.45 in [-1, 1]): a raw random-init decoder is either flat grey or, with a
    large gain, saturated far outside [-1, 1] where an 8-bit PSNR says nothing about relative accuracy."""
    from drb200.CleanVAE import AutoencoderKLCosmos, CleanVAE
    sd = vo.make_vae_state_dict(VDIMS, seed=seed)
    with torch.no_grad():
        probe = torch.randn(1, 16, 2, 8, 8, generator=torch.Generator().manual_seed(1))
        gain = target_std / float(vo.decode(sd, VDIMS, probe).std())
    for k in ("decoder.conv_out.conv_t.weight", "decoder.conv_out.conv_t.bias"):
        sd[k] = sd[k] * gain
    model = AutoencoderKLCosmos(encoder_block_out_channels=VDIMS.encoder_block_out_channels,
                                decode_block_out_channels=VDIMS.decode_block_out_channels)
    model.load_state_dict(sd, strict=True)
    vae = CleanVAE(model=model)
    vae.to(DEV)
    vae.reset_dtype(torch.bfloat16)
    return vae, {k: v.to(DEV).bfloat16() for k, v in sd.items()}


def _pipeline(dims, model_type, vae, steps, guidance=0.0):
    from drb200.diffusion_renderer_pipeline import CleanDiffusionRendererPipeline
    model, sd = build_product_model(dims, model_type, seed=3)
    pipe = CleanDiffusionRendererPipeline(checkpoint_dir="", checkpoint_name="", model_type=model_type, vae_instance=vae,
                                          model_instance=model, guidance=guidance, num_steps=steps, seed=42)
    return pipe, model, net_only(sd)


def _oracle_video(sdn, dims, vsd, batch, keys, append_mask, ci, steps, seed, thw, guidance=0.0):
    T, H, W = thw
    shape = (16, (T - 1) // 8 + 1, H // 8, W // 8)
    with torch.no_grad():
        cond = so.latent_conditions(batch, keys, append_mask, lambda x: vo.encode(vsd, VDIMS, x), (1, *shape))
        z = so.sample(sdn, dims, cond, ci, shape, steps, seed, guidance=guidance)
        video = vo.decode(vsd, VDIMS, z / so.SIGMA_DATA)
    return z, video


@pytest.mark.parametrize("dims,thw,steps", [(MICRO_INVERSE, (9, 64, 96), 3), (TINY_INVERSE, (17, 64, 64), 2)])
def test_inverse_generate_video_matches_oracle_chain(dims, thw, steps):
    vae, vsd = _vae()
    pipe, model, sdn = _pipeline(dims, "inverse", vae, steps)
    T, H, W = thw
    clip = (torch.rand(1, 3, T, H, W, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1234)) * 2 - 1)
    ci = torch.full((1, 1), 3, dtype=torch.long, device=DEV)
    for normalize in (False, True):
        got = pipe.generate_video({"rgb": clip, "video": clip, "context_index": ci}, normalize_normal=normalize, seed=42)
        clip16 = clip.bfloat16()
        z_ref, video_ref = _oracle_video(sdn, dims, vsd, {"rgb": clip16}, ["rgb"], False, ci, steps, 42, thw)
        ref = so.postprocess(video_ref, normalize)
        assert got.shape == ref.shape == (1, T, H, W, 3) and got.dtype == np.uint8
        p = psnr_u8(got, ref)
        print(f"\ninverse {thw} normalize_normal={normalize}: PSNR {p:.1f} dB, frame std {ref.std():.1f}")
        assert ref.std() > 8.0                       # the frames are not flat: the PSNR gate means something
        assert p >= 40.0
    # the tensors behind the frames
    batch = pipe._move_to_device({"rgb": clip, "video": clip, "context_index": ci})
    with torch.no_grad():
        z = model.generate_samples_from_batch(batch, guidance=0.0, seed=42, state_shape=list(z_ref.shape[1:]), num_steps=steps)
        video = model.decode(z)
    e_z, e_v = rel_l2(z, z_ref), rel_l2(video, video_ref)
    print(f"final latent rel-L2 {e_z:.3e}; decoded video rel-L2 {e_v:.3e}")
    assert e_z <= 1e-2
    assert e_v <= 3e-2


def _forward_batch(T, H, W, seed=77):
    g = torch.Generator(device=DEV).manual_seed(seed)
    keys = ["basecolor", "normal", "metallic", "roughness", "depth", "env_ldr", "env_log", "env_nrm"]
    return keys, {k: (torch.rand(1, 3, T, H, W, device=DEV, generator=g) * 2 - 1) for k in keys}


def test_forward_generate_video_matches_oracle_chain():
    """BASELINE configs[2] at micro size: 8 encoded conditions + mask channels -> 136-channel latent condition."""
    vae, vsd = _vae()
    steps, thw = 2, (9, 32, 48)
    pipe, model, sdn = _pipeline(MICRO_FORWARD, "forward", vae, steps)
    keys, batch = _forward_batch(*thw)
    batch["video"] = batch["depth"]
    got = pipe.generate_video(dict(batch), seed=42)
    b16 = {k: v.bfloat16() for k, v in batch.items()}
    z_ref, video_ref = _oracle_video(sdn, MICRO_FORWARD, vsd, b16, keys, True, None, steps, 42, thw)
    ref = so.postprocess(video_ref)
    p = psnr_u8(got, ref)
    print(f"\nforward {thw}: PSNR {p:.1f} dB, frame std {ref.std():.1f}")
    assert got.shape == ref.shape and p >= 40.0
    # a missing condition key is zeros + a zero mask channel (model_diffusion_renderer.py:181-186)
    part = {k: v for k, v in batch.items() if k != "env_log"}
    got2 = pipe.generate_video(dict(part), seed=42)
    _, video_ref2 = _oracle_video(sdn, MICRO_FORWARD, vsd, {k: v for k, v in b16.items() if k != "env_log"}, keys, True, None,
                                  steps, 42, thw)
    assert psnr_u8(got2, so.postprocess(video_ref2)) >= 40.0


def test_guidance_runs_two_forwards_per_step():
    vae, vsd = _vae()
    steps, thw = 2, (9, 32, 48)
    pipe, model, sdn = _pipeline(MICRO_INVERSE, "inverse", vae, steps, guidance=2.0)
    clip = (torch.rand(1, 3, *thw, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5)) * 2 - 1)
    ci = torch.full((1, 1), 1, dtype=torch.long, device=DEV)
    got = pipe.generate_video({"rgb": clip, "video": clip, "context_index": ci}, seed=7)
    _, video_ref = _oracle_video(sdn, MICRO_INVERSE, vsd, {"rgb": clip.bfloat16()}, ["rgb"], False, ci, steps, 7, thw, guidance=2.0)
    assert psnr_u8(got, so.postprocess(video_ref)) >= 40.0


def test_inverse_node_returns_five_gbuffers(monkeypatch):
    """Cosmos1InverseRenderer.run_inverse_pass (nodes.py:144-215): five passes over one clip, float32 (B*T,H,W,3) in [0,1]."""
    for name in ("comfy", "comfy.utils", "comfy.model_management", "folder_paths"):
        monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    from drb200 import nodes
    vae, vsd = _vae()
    steps, thw = 2, (9, 32, 48)
    pipe, model, sdn = _pipeline(MICRO_INVERSE, "inverse", vae, steps)
    T, H, W = thw
    image = torch.rand(1, T, H, W, 3, generator=torch.Generator().manual_seed(9))        # ComfyUI IMAGE, CPU, [0,1]
    outs = nodes.Cosmos1InverseRenderer().run_inverse_pass(pipe, image, guidance=0.0, seed=42)
    assert len(outs) == 5
    clip16 = (image.permute(0, 4, 1, 2, 3) * 2.0 - 1.0).to(DEV).bfloat16()
    for name, out in zip(nodes.INFERENCE_PASSES, outs):
        assert out.shape == (T, H, W, 3) and out.dtype == torch.float32 and out.device.type == "cpu"
        assert 0.0 <= float(out.min()) and float(out.max()) <= 1.0
        ci = torch.full((1, 1), nodes.GBUFFER_INDEX_MAPPING[name], dtype=torch.long, device=DEV)
        _, video_ref = _oracle_video(sdn, MICRO_INVERSE, vsd, {"rgb": clip16}, ["rgb"], False, ci, steps, 42, thw)
        ref = so.postprocess(video_ref, name == "normal")[0]
        assert psnr_u8((out.numpy() * 255).round().astype(np.uint8), ref) >= 40.0
    assert not torch.equal(outs[0], outs[1])          # the passes differ through the context embedding


def test_single_image_path_matches_oracle_chain():
    """T = 1 (a ComfyUI IMAGE of one frame): tokenizer image path, one latent frame through the DiT, decode to one frame"""
    vae, vsd = _vae()
    steps, thw = 2, (1, 64, 96)
    pipe, model, sdn = _pipeline(MICRO_INVERSE, "inverse", vae, steps)
    clip = (torch.rand(1, 3, *thw, device=DEV, generator=torch.Generator(device=DEV).manual_seed(11)) * 2 - 1)
    ci = torch.full((1, 1), 0, dtype=torch.long, device=DEV)
    got = pipe.generate_video({"rgb": clip, "video": clip, "context_index": ci}, seed=3)
    _, video_ref = _oracle_video(sdn, MICRO_INVERSE, vsd, {"rgb": clip.bfloat16()}, ["rgb"], False, ci, steps, 3, thw)
    assert got.shape == (1, 1, 64, 96, 3)
    assert psnr_u8(got, so.postprocess(video_ref)) >= 40.0


def test_batched_clips_are_rejected_like_the_reference():
    """the reference sampler is batch-1 (model_diffusion_renderer.py:222): B > 1 must fail loudly, not silently
    render the first clip"""
    vae, _ = _vae()
    pipe, model, _ = _pipeline(MICRO_INVERSE, "inverse", vae, 2)
    clip = torch.rand(2, 3, 9, 32, 48, device=DEV) * 2 - 1
    ci = torch.zeros(2, 1, dtype=torch.long, device=DEV)
    with pytest.raises((ValueError, RuntimeError)):
        pipe.generate_video({"rgb": clip, "video": clip, "context_index": ci}, seed=3)
    with pytest.raises(ValueError):
        pipe.generate_video({"context_index": ci})                      # no tensor to infer the clip shape from


def test_inverse_node_with_batched_passes_equals_the_pass_loop(monkeypatch):
    """pipeline.batch_passes: the five G-buffer passes as ONE sampler run (generate_video_passes, the passes batched along the
    token rows of every kernel) must return exactly the frames of the reference's pass-by-pass loop (nodes.py:187-205)"""
    for name in ("comfy", "comfy.utils", "comfy.model_management", "folder_paths"):
        monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    from drb200 import nodes
    vae, _ = _vae()
    pipe, model, _ = _pipeline(MICRO_INVERSE, "inverse", vae, 2)
    image = torch.rand(1, 9, 32, 48, 3, generator=torch.Generator().manual_seed(9))
    assert not pipe.wants_batched_passes()                # automatic: only when the net is context-parallel
    loop = nodes.Cosmos1InverseRenderer().run_inverse_pass(pipe, image, guidance=0.0, seed=42)
    pipe.batch_passes, pipe.pass_batch = True, 5
    batched = nodes.Cosmos1InverseRenderer().run_inverse_pass(pipe, image, guidance=0.0, seed=42)
    pipe.pass_batch = 2                                   # 2 + 2 + 1 passes per sampler run
    chunked = nodes.Cosmos1InverseRenderer().run_inverse_pass(pipe, image, guidance=0.0, seed=42)
    for a, b, c in zip(loop, batched, chunked):
        assert torch.equal(a, b) and torch.equal(a, c)
    pipe.fuse_postprocess = False                         # decode -> planar video -> post-process kernel: the same frames
    unfused = nodes.Cosmos1InverseRenderer().run_inverse_pass(pipe, image, guidance=0.0, seed=42)
    pipe.fuse_postprocess = True
    for a, b in zip(loop, unfused):
        assert torch.equal(a, b)
    loop_g = nodes.Cosmos1InverseRenderer().run_inverse_pass(pipe, image, guidance=1.5, seed=7)     # CFG: 10 sequences per step
    pipe.batch_passes = False
    ref_g = nodes.Cosmos1InverseRenderer().run_inverse_pass(pipe, image, guidance=1.5, seed=7)
    for a, b in zip(loop_g, ref_g):
        assert torch.equal(a, b)


def test_loader_node_to_inverse_pass_on_the_gpu(tmp_path, monkeypatch):
    """checkpoint files -> LoadDiffusionRendererModel.load_pipeline (meta skeleton, to_empty on the GPU, strict load,
    nodes.py:94-115) -> Cosmos1InverseRenderer: the same frames as a model that was handed its weights directly"""
    from drb200 import nodes
    from drb200.CleanVAE import AutoencoderKLCosmos, CleanVAE
    from drb200.diffusion_renderer_pipeline import CleanDiffusionRendererPipeline
    from drb200.model_diffusion_renderer import CleanDiffusionRendererModel
    from tests.test_loader_cpu import _stub_comfy, write_models
    from tests.util import model_config
    sd, vsd = write_models(str(tmp_path), MICRO_INVERSE, True, "inverse.pt")
    _stub_comfy(monkeypatch, str(tmp_path), DEV)
    (pipe,) = nodes.LoadDiffusionRendererModel().load_pipeline("inverse.pt")
    pipe.num_steps = 2
    image = torch.rand(1, 9, 32, 48, 3, generator=torch.Generator().manual_seed(4))
    got = nodes.Cosmos1InverseRenderer().run_inverse_pass(pipe, image, guidance=0.0, seed=11)
    model = CleanDiffusionRendererModel(model_config(MICRO_INVERSE, "inverse"))
    model.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True)
    model = model.to(device=DEV, dtype=torch.bfloat16)
    vm = AutoencoderKLCosmos(encoder_block_out_channels=VDIMS.encoder_block_out_channels, decode_block_out_channels=VDIMS.decode_block_out_channels)
    vm.load_state_dict(vsd, strict=True)
    vae = CleanVAE(model=vm)
    vae.to(DEV)
    vae.reset_dtype(torch.bfloat16)
    ref_pipe = CleanDiffusionRendererPipeline(checkpoint_dir="", checkpoint_name="", model_type=None, vae_instance=vae, model_instance=model,
                                              guidance=0.0, num_steps=2, seed=42)
    want = nodes.Cosmos1InverseRenderer().run_inverse_pass(ref_pipe, image, guidance=0.0, seed=11)
    for a, b in zip(got, want):
        assert torch.isfinite(a).all() and torch.equal(a, b)


def test_pinned_output_returns_the_same_frames():
    """pipeline.pinned_output: frames arrive in page-locked staging buffers owned by the pipeline (valid until the next
    call) — same bytes as the fresh-array path"""
    vae, _ = _vae()
    pipe, model, _ = _pipeline(MICRO_INVERSE, "inverse", vae, 2)
    clip = (torch.rand(1, 3, 9, 32, 48, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5)) * 2 - 1)
    batch = lambda k: {"rgb": clip, "video": clip, "context_index": torch.full((1, 1), k, dtype=torch.long, device=DEV)}
    fresh = [pipe.generate_video(batch(k), seed=7).copy() for k in (0, 3)]
    pipe.pinned_output = True
    a = pipe.generate_video(batch(0), seed=7)
    assert np.array_equal(a, fresh[0])
    b = pipe.generate_video(batch(3), seed=7)          # reuses the staging buffer: `a` now shows pass 3
    assert np.array_equal(b, fresh[1]) and np.shares_memory(a, b)
    outs = pipe.generate_video_passes({"rgb": clip, "video": clip}, [0, 3], seed=7)
    assert np.array_equal(outs[0], fresh[0]) and np.array_equal(outs[1], fresh[1]) and not np.shares_memory(outs[0], outs[1])
