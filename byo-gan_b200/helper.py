"""Optional drop-in for the reference's ``helper.py`` (SURVEY.md §8(f) rank 1 — host-side step overheads in the caller).

``train.py`` draws three latent batches per iteration with ``helper.get_truncated_noise`` (helper.py:36-45): scipy's
``truncnorm.rvs`` on the host, a float64 -> float32 conversion and a host->device copy — 3 / 7 / 15 ms for batch 16 / 32 /
64, i.e. as long as the whole B200 training step.  This module keeps the three public names and their semantics but
samples on the device.  Put ``byo-gan_b200/`` first on ``sys.path`` and ``import helper`` resolves here;
leave it out and the reference's own helper is used — nothing in the hot path depends on it.

* ``get_truncated_noise(n_samples, z_dim, trunc)``: standard normal truncated to [-trunc, trunc], float32, on the current
  CUDA device, ``requires_grad=True`` — the same distribution as truncnorm.rvs(-trunc, trunc) (inverse-CDF sampling;
  the random STREAM differs from scipy's, as it would between two scipy seeds).
* ``set_requires_grad(model, flag)``: helper.py:48-50.
* ``display_image(...)``: helper.py:8-33; matplotlib / torchvision are imported lazily so that training does not need
  them unless a preview is actually drawn.
"""
from math import sqrt

import torch

_SQRT2 = sqrt(2.0)


def get_truncated_noise(n_samples, z_dim, trunc):
    trunc = float(trunc)
    # inverse CDF of the normal restricted to [-t, t]: u ~ U(Phi(-t), Phi(t)), x = Phi^-1(u)
    lo = 0.5 * (1.0 + torch.erf(torch.tensor(-trunc / _SQRT2, dtype=torch.float64)).item())
    hi = 1.0 - lo
    u = torch.empty(n_samples, z_dim, device="cuda", dtype=torch.float32).uniform_(lo, hi)
    noise = torch.erfinv(2.0 * u - 1.0).mul_(_SQRT2).clamp_(-trunc, trunc)
    return noise.requires_grad_()


def set_requires_grad(model, requires_grad: bool):
    for p in model.parameters():
        p.requires_grad = requires_grad


def display_image(images, num_display=4, save_to_disk=False, save_dir="./output", filename="figure", title="Images"):
    import matplotlib.pyplot as plt
    from torchvision import utils

    if images.dim() == 3:
        plt.imshow(images.detach().cpu().permute(1, 2, 0))
    else:
        nrow = int(sqrt(num_display))
        grid = utils.make_grid(images.detach().cpu()[:num_display], nrow=nrow)
        plt.imshow(grid.permute(1, 2, 0).squeeze())
    plt.title(title)
    if save_to_disk:
        plt.savefig("{0}/{1}.png".format(save_dir, filename))
    else:
        plt.show()
