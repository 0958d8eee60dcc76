"""CPU tier: host logic of the product package and the C-ABI surface (no kernel is launched here)."""
import ctypes
import os
import re

import pytest
import torch

from oracle.weights import FULL_FORWARD, FULL_INVERSE, MICRO_FORWARD, MICRO_INVERSE, TINY_INVERSE, net_param_shapes
from tests.util import model_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from drb200 import _lib
    header = open(os.path.join(ROOT, "include", "drb200.h")).read()
    declared = set(re.findall(r"\b(drb_[a-z0-9_]+)\s*\(", header))
    assert {"drb_gemm_bf16", "drb_attention_bf16", "drb_adaln_modulate", "drb_unpatchify_euler"} <= declared
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/drb200.h but not exported"
    bound = set(_lib.PROTOTYPES) | {"drb_last_error", "drb_version", "drb_device_supported"}
    assert declared == bound, f"ctypes binding out of sync with the header: {declared ^ bound}"
    assert _lib.load().drb_version() >= 100


def test_invalid_arguments_raise_valueerror_without_a_gpu():
    from drb200 import _lib
    with pytest.raises(ValueError):
        _lib.call("drb_gemm_bf16", None, 8, None, 8, None, 8, 1, 8, 8, 0, None, 0, None, 0, None)
    with pytest.raises(ValueError):
        _lib.call("drb_adaln_modulate", 16, 16, 16, 16, None, None, 4, 100, None)   # D not a multiple of 256


@pytest.mark.parametrize("dims,mt", [(MICRO_INVERSE, "inverse"), (MICRO_FORWARD, "forward"), (TINY_INVERSE, "inverse")])
def test_state_dict_keys_and_shapes_match_reference_layout(dims, mt):
    from drb200.model_diffusion_renderer import CleanDiffusionRendererModel
    model = CleanDiffusionRendererModel(model_config(dims, mt))
    sd = model.state_dict()
    want = {k: tuple(s) for k, s, _ in net_param_shapes(dims)}
    assert set(sd) == set(want)
    for k, s in want.items():
        assert tuple(sd[k].shape) == s, k


@pytest.mark.parametrize("dims,mt,n", [(FULL_INVERSE, "inverse", 572), (FULL_FORWARD, "forward", 571)])
def test_full_size_key_count_on_meta_device(dims, mt, n):
    from drb200.model_diffusion_renderer import CleanDiffusionRendererModel
    with torch.device("meta"):
        model = CleanDiffusionRendererModel(model_config(dims, mt, 704, 1280, 57))
    sd = model.state_dict()
    assert len(sd) == n                                    # SURVEY.md §8b [PROBED: 572 / 571 entries]
    net_params = sum(v.numel() for k, v in sd.items() if k.startswith("net.") and k != "net.pos_embedder.seq")
    assert net_params == (7_234_963_456 if mt == "inverse" else 7_236_913_152)


def test_cpu_forward_fails_loudly():
    from drb200.model_diffusion_renderer import CleanDiffusionRendererModel
    model = CleanDiffusionRendererModel(model_config(MICRO_INVERSE, "inverse"))
    x = torch.zeros(1, 16, 2, 8, 12)
    with pytest.raises(RuntimeError, match="CUDA"):
        model.net(x, torch.tensor(1.0), x, torch.zeros(1, 1, dtype=torch.long))


def test_config_matches_reference_values():
    from drb200 import diffusion_renderer_config as cfg
    inv = cfg.get_inverse_renderer_config()
    fwd = cfg.get_forward_renderer_config()
    assert inv["condition_keys"] == ["rgb"] and inv["append_condition_mask"] is False
    assert inv["net"]["additional_concat_ch"] == 16 and inv["net"]["use_context_embedding"] is True
    assert fwd["net"]["additional_concat_ch"] == 136 and fwd["net"]["use_context_embedding"] is False
    assert len(fwd["condition_keys"]) == 8 and fwd["append_condition_mask"] is True
    assert inv["latent_shape"] == [16, 8, 88, 160]
    assert inv["scheduler"]["sigma_max"] == 80.0 and inv["scheduler"]["sigma_min"] == 0.02
    cfg.validate_config(inv)
    with pytest.raises(ValueError):
        cfg.get_config_from_tensor_shape("inverse", (1, 3, 704, 1280))
    with pytest.raises(ValueError):
        cfg.validate_config({k: v for k, v in inv.items() if k != "net"})


def test_scheduler_sigmas_match_oracle():
    from drb200.model_diffusion_renderer import CleanEDMEulerScheduler
    from oracle.sampler_oracle import sigma_schedule
    s = CleanEDMEulerScheduler()
    s.set_timesteps(15)
    assert torch.equal(s.sigmas, sigma_schedule(15))
    assert s.timesteps.numel() == 15 and s.sigmas[-1] == 0
    with pytest.raises(RuntimeError):
        CleanEDMEulerScheduler().step(torch.zeros(1), torch.tensor(1.0), torch.zeros(1))


# ------------------------------------------------------------------------------------------------ tokenizer host logic
def test_tokenizer_state_dict_layout_matches_the_restated_upstream_class():
    from drb200.CleanVAE import AutoencoderKLCosmos
    from oracle.vae_oracle import FULL_VAE, SMALL_VAE, vae_param_shapes
    with torch.device("meta"):
        full = AutoencoderKLCosmos()
    sd = full.state_dict()
    want = dict(vae_param_shapes(FULL_VAE))
    assert set(sd) == set(want) and len(sd) == 310
    assert all(tuple(sd[k].shape) == want[k] for k in want)
    assert sum(v.numel() for v in sd.values()) == 105_653_696          # ~105.6 M
    small = AutoencoderKLCosmos(encoder_block_out_channels=SMALL_VAE.encoder_block_out_channels,
                                decode_block_out_channels=SMALL_VAE.decode_block_out_channels)
    assert set(small.state_dict()) == set(dict(vae_param_shapes(SMALL_VAE)))
    # only the first down block / the middle up block resample (8x = 4x Haar * 2x); attention only in the mid blocks
    assert [(b["spatial"], b["temporal"]) for b in full.enc_plan] == [(True, True), (False, False), (False, False)]
    assert [(b["spatial"], b["temporal"]) for b in full.dec_plan] == [(False, False), (True, True), (False, False)]
    assert not any("attentions" in k and "mid_block" not in k for k in sd)


def test_tokenizer_wrapper_surface_and_errors(tmp_path):
    from drb200.CleanVAE import AutoencoderKLCosmos, CleanVAE
    model = AutoencoderKLCosmos(encoder_block_out_channels=(64, 64, 64, 64), decode_block_out_channels=(64, 64, 64, 64))
    vae = CleanVAE(model=model)
    assert (vae.latent_ch, vae.spatial_compression_factor, vae.temporal_compression_factor) == (16, 8, 8)
    assert [vae.get_latent_num_frames(n) for n in (1, 9, 57, 121)] == [1, 2, 8, 16]
    assert [vae.get_pixel_num_frames(n) for n in (1, 2, 8, 16)] == [1, 9, 57, 121]
    with pytest.raises(ValueError):
        vae.encode(torch.zeros(3, 9, 32, 32))
    with pytest.raises(ValueError):
        vae.decode(torch.zeros(16, 2, 4, 4))
    with pytest.raises(RuntimeError, match="CUDA"):                      # no CPU fallback
        vae.encode(torch.zeros(1, 3, 9, 32, 32))
    with pytest.raises(ValueError):
        CleanVAE()
    with pytest.raises(FileNotFoundError):
        CleanVAE(model_path=str(tmp_path))
    with pytest.raises(ValueError):
        AutoencoderKLCosmos(patch_size=2)
    # diffusers checkpoint directory: config.json + diffusion_pytorch_model.safetensors, strict key match
    import json
    from safetensors.torch import save_file
    cfg = {"encoder_block_out_channels": [64, 64, 64, 64], "decode_block_out_channels": [64, 64, 64, 64], "_class_name": "AutoencoderKLCosmos"}
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    sd = {k: torch.randn_like(v) for k, v in model.state_dict().items()}
    save_file(sd, str(tmp_path / "diffusion_pytorch_model.safetensors"))
    loaded = CleanVAE(model_path=str(tmp_path))
    got = loaded.model.state_dict()
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    loaded.reset_dtype(torch.bfloat16)
    assert next(loaded.model.parameters()).dtype == torch.bfloat16


def test_conv_argument_block_matches_the_header():
    """ctypes mirror of drb_conv3d_args: same field order as include/drb200.h, and bad blocks are rejected without a GPU"""
    from drb200 import _lib
    header = open(os.path.join(ROOT, "include", "drb200.h")).read()
    body = header[header.index("typedef struct drb_conv3d_args {"):header.index("} drb_conv3d_args;")]
    names = re.findall(r"\b([A-Za-z_]+)(?=[,;])", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert names == [f[0] for f in _lib.Conv3dArgs._fields_]
    a = _lib.Conv3dArgs()
    with pytest.raises(ValueError):
        _lib.call("drb_conv3d_cl", a, None)                                # null pointers
    with pytest.raises(ValueError):
        _lib.call("drb_haar_patch", 8, 8, 3, 10, 32, 32, None)              # frames not 1 + 4k
    with pytest.raises(ValueError):
        _lib.call("drb_softmax_rows", 16, 12, 4, 12, 1.0, None)             # pitch not a multiple of 8


# ------------------------------------------------------------------------------------------------ node surface (no ComfyUI)
def test_node_registry_and_signatures_match_the_reference_surface():
    """reference nodes.py:62-72, :131-149, :219-247, :313-323, :335-347"""
    import inspect

    import drb200
    from drb200 import nodes
    assert set(drb200.NODE_CLASS_MAPPINGS) == {"LoadDiffusionRendererModel", "Cosmos1InverseRenderer", "Cosmos1ForwardRenderer",
                                               "LoadHDRImage"}
    assert set(drb200.NODE_DISPLAY_NAME_MAPPINGS) == set(drb200.NODE_CLASS_MAPPINGS)
    inv, fwd, load = nodes.Cosmos1InverseRenderer, nodes.Cosmos1ForwardRenderer, nodes.LoadDiffusionRendererModel
    assert inv.RETURN_TYPES == ("IMAGE",) * 5 and inv.RETURN_NAMES == ("base_color", "metallic", "roughness", "normal", "depth")
    assert inv.FUNCTION == "run_inverse_pass" and fwd.FUNCTION == "run_forward_pass" and load.FUNCTION == "load_pipeline"
    assert inv.CATEGORY == fwd.CATEGORY == load.CATEGORY == "Cosmos1"
    assert load.RETURN_TYPES == ("DIFFUSION_RENDERER_PIPELINE",) and fwd.RETURN_TYPES == ("IMAGE",)
    assert list(inspect.signature(inv.run_inverse_pass).parameters) == ["self", "pipeline", "image", "guidance", "seed"]
    assert list(inspect.signature(fwd.run_forward_pass).parameters) == [
        "self", "pipeline", "depth", "normal", "roughness", "metallic", "base_color", "env_map", "guidance", "seed", "env_format",
        "env_brightness", "env_flip_horizontal", "env_rotation"]
    req = inv.INPUT_TYPES()
    assert set(req["required"]) == {"pipeline", "image"} and set(req["optional"]) == {"guidance", "seed"}
    assert set(fwd.INPUT_TYPES()["required"]) == {"pipeline", "depth", "normal", "roughness", "metallic", "base_color", "env_map"}
    assert nodes.GBUFFER_INDEX_MAPPING == {"basecolor": 0, "metallic": 1, "roughness": 2, "normal": 3, "depth": 4}


def test_image_inputs_are_normalised_like_the_reference():
    """nodes.py:156-177: 3-D adds batch and time, 4-D adds time, lists are stacked, anything else raises"""
    from drb200.nodes import _to_5d, latlong_vec
    assert _to_5d(torch.zeros(8, 12, 3)).shape == (1, 1, 8, 12, 3)
    assert _to_5d(torch.zeros(5, 8, 12, 3)).shape == (5, 1, 8, 12, 3)
    assert _to_5d(torch.zeros(1, 9, 8, 12, 3)).shape == (1, 9, 8, 12, 3)
    assert _to_5d([torch.zeros(9, 8, 12, 3)]).shape == (1, 9, 8, 12, 3)
    with pytest.raises(ValueError):
        _to_5d(torch.zeros(8, 12))
    with pytest.raises(TypeError):
        _to_5d("clip.mp4")
    v = latlong_vec((6, 10))
    assert v.shape == (6, 10, 3) and torch.allclose(v.norm(dim=-1), torch.ones(6, 10), atol=1e-6)


def test_envmap_module_host_logic():
    from drb200 import preprocess_envmap as pe
    assert pe.process_comfyui_tensor(torch.zeros(2, 16, 32, 3)).shape == (16, 32, 3)
    assert pe.process_comfyui_tensor(torch.zeros(1, 4, 16, 32)).shape == (16, 32, 3)      # (B,C,H,W) with alpha
    assert pe.process_comfyui_tensor(torch.zeros(16, 32, 1)).shape == (16, 32, 3)
    a, b = torch.rand(1, 8, 16, 3), torch.rand(1, 8, 16, 3)
    assert pe.compute_tensor_hash(a) == pe.compute_tensor_hash(a.clone()) != pe.compute_tensor_hash(b)
    assert pe._key(a, (4, 8), "proj", 1.0, True, 180.0, "cuda") != pe._key(a, (4, 8), "proj", 1.5, True, 180.0, "cuda")
    assert pe._key(a, (4, 8), "proj", 1.0, True, 180.0, "cuda:0") != pe._key(a, (4, 8), "proj", 1.0, True, 180.0, "cuda:1")
    cache = pe.EnvironmentMapCache(max_size=2)
    for i in range(3):
        cache.put(("k", i), {"env_ldr": i})
    assert cache.get(("k", 0)) is None and cache.get(("k", 2)) == {"env_ldr": 2}
    # LRU, and overwriting an existing key evicts nothing
    assert cache.get(("k", 1)) == {"env_ldr": 1}          # 1 is now the most recently used
    cache.put(("k", 1), {"env_ldr": 10})
    assert len(cache.cache) == 2 and cache.get(("k", 2)) is not None
    cache.get(("k", 1))
    cache.put(("k", 3), {"env_ldr": 3})                   # evicts 2 (least recently used), not 1
    assert cache.get(("k", 2)) is None and cache.get(("k", 1)) == {"env_ldr": 10}
    with pytest.raises(ValueError):
        pe.render_projection_from_panorama(3.14, (8, 8))
    with pytest.raises(ValueError):                                                     # CPU tensor: the device path refuses
        pe.render_projection_from_panorama(torch.rand(8, 16, 3), (4, 8), device="cpu", use_cache=False)
    with pytest.raises(RuntimeError):
        pe.load_hdr_file("/nonexistent/file.hdr")


def test_pipeline_surface_and_model_selection_errors():
    """diffusion_renderer_pipeline.py:38-45, :99, :242"""
    import inspect

    from drb200.diffusion_renderer_pipeline import CleanDiffusionRendererPipeline as P
    sig = inspect.signature(P.__init__).parameters
    assert list(sig)[1:] == ["checkpoint_dir", "checkpoint_name", "model_type", "vae_instance", "model_instance", "guidance", "num_steps",
                             "height", "width", "num_video_frames", "seed", "dtype"]
    assert (sig["guidance"].default, sig["num_steps"].default, sig["seed"].default, sig["model_type"].default) == (2.0, 20, 42, "inverse")
    p = P("", "", model_type="Inverse", model_instance=None)
    assert p.model_type == "inverse"
    p.config, p.model = {"x": 1}, object()
    p.set_model_type("forward")
    assert p.model_type == "forward" and p.config is None and p.model is None
    with pytest.raises(ValueError):
        p.generate_video({"context_index": torch.zeros(1, 1)})
    with pytest.raises(RuntimeError):                      # no pre-loaded model: the reference's fallback is dead code (:227)
        p.generate_video({"depth": torch.zeros(1, 3, 9, 64, 96)})
