"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU (or same-device torch) restatement of the reference algorithm for the
DiffusionRenderer denoising hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package; the product
package (``diffusionrenderer-comfyui_b200``) never does.

Parity status
-------------
* ``dit_oracle`` / ``sampler_oracle``: PINNED.  Checked bit-for-bit (fp32, CPU)
  against the real reference modules imported from ``/root/reference`` with the
  one documented patch (flatten heads before ``to_out``); the
  outputs are committed under ``tests/golden/`` together with the generating
  script ``tests/golden/make_golden.py``.
* ``vae_oracle``: **parity unpinned** — the tokenizer arithmetic lives in
  ``diffusers.AutoencoderKLCosmos`` (diffusers >= 0.34, not installed, not
  vendored); the restatement follows ``VAE_config.json`` + SURVEY Appendix B and
  is only self-checked (shapes, causality, Haar round trip, parameter count).
* ``envmap_oracle``: the torch stages of ``preprocess_envmap.py`` are PINNED bit-for-bit against the reference's own
  functions (imported with stubbed nvdiffrast / imageio / cv2); the cube-map fetch (``nvdiffrast.torch.texture``) is
  **parity unpinned** and self-checked (texel-centre exactness, seamlessness).
"""
