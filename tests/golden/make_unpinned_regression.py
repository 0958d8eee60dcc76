"""Regression fixtures for the two UNPINNED oracle parts (tokenizer, nvdiffrast cube fetch): outputs of the oracle ITSELF
on fixed seeds, so that an accidental change of the restatement is caught.  They prove nothing about upstream parity (see
oracle/vae_oracle.py, oracle/envmap_oracle.py); the envmap torch stages ARE pinned, in tests/test_oracle_vs_reference.py.

    python tests/golden/make_unpinned_regression.py      # CPU, fp32
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import envmap_oracle as eo        # noqa: E402
from oracle import vae_oracle as vo           # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.manual_seed(0)
    sd = vo.make_vae_state_dict(vo.SMALL_VAE, seed=7)
    g = torch.Generator().manual_seed(21)
    x = torch.rand(1, 3, 9, 32, 48, generator=g) * 2 - 1
    with torch.no_grad():
        z = vo.encode(sd, vo.SMALL_VAE, x)
        y = vo.decode(sd, vo.SMALL_VAE, z)
    pano = torch.rand(16, 32, 3, generator=g) ** 3 * 20
    env = eo.render_projection_from_panorama(pano, (12, 20), 1.3, True, 180.0, cube_res=16)
    np.savez_compressed(os.path.join(HERE, "unpinned_regression.npz"), x=x.numpy(), z=z.numpy(), y_sub=y[:, :, ::4, ::8, ::8].numpy(),
                        pano=pano.numpy(), env_ldr=env["env_ldr"].numpy(), env_log=env["env_log"].numpy())
    print("wrote unpinned_regression.npz", z.shape, y.shape)


if __name__ == "__main__":
    main()
