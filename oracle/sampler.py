"""CPU restatement of the predictor-corrector sampler (sampling.py), with the random draws passed IN.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference draws `torch.randn_like(x)` inside every predictor / corrector call (sampling.py:207,284);
for parity the noise tensors are arguments here (SURVEY.md §7 hard part 3).  Pinned by tests/golden/pc_*.npz
which were produced by the live reference with `torch.randn_like` patched to replay the same tensors.
"""
import torch


def reverse_diffusion_update(sde, score, x, t, z, probability_flow=False):
    """ReverseDiffusionPredictor.update_fn (sampling.py:205-210) with RSDE.discretize (sde_lib.py:105-118),
    next_t=None.  `score` = score_fn(x, t) already evaluated."""
    f, G = sde.discretize(x, t)
    rev_f = f - G[:, None, None, None] ** 2 * score * (0.5 if probability_flow else 1.)
    rev_G = torch.zeros_like(G) if probability_flow else G
    x_mean = x - rev_f
    return x_mean + rev_G[:, None, None, None] * z, x_mean


def langevin_update(sde, score, x, t, z, snr):
    """One inner step of LangevinCorrector.update_fn (sampling.py:272-292): batch-mean norms."""
    alpha = sde.alpha_for_corrector(t)
    grad_norm = torch.norm(score.reshape(score.shape[0], -1), dim=-1).mean()
    noise_norm = torch.norm(z.reshape(z.shape[0], -1), dim=-1).mean()
    step = (snr * noise_norm / grad_norm) ** 2 * 2 * alpha
    x_mean = x + step[:, None, None, None] * score
    return x_mean + torch.sqrt(step * 2)[:, None, None, None] * z, x_mean


def pc_sampler(sde, score_fn, x_init, noises, num_scales, eps, snr, corrector='none', n_steps=1, denoise=True):
    """pc_sampler loop (sampling.py:418-447) up to (not including) the flow inverse.

    noises: iterator of tensors consumed in the reference's draw order (corrector first, then predictor)."""
    x = x_init
    timesteps = torch.linspace(sde.T, eps, num_scales)
    it = iter(noises)
    x_mean = x
    for i in range(num_scales):
        vec_t = torch.ones(x.shape[0]) * timesteps[i]
        if corrector == 'langevin':
            for _ in range(n_steps):
                x, x_mean = langevin_update(sde, score_fn(x, vec_t), x, vec_t, next(it), snr)
        x, x_mean = reverse_diffusion_update(sde, score_fn(x, vec_t), x, vec_t, next(it))
    return x_mean if denoise else x
