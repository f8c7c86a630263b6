"""Drop-in for the reference's ``src/train/loss.py`` (PhaseNet loss): L1 on the reconstructed image plus the wrapped
phase difference ``mean |atan2(sin d, cos d)|`` per level and orientation (loss.py:5-25).  Differentiable on CUDA tensors;
the image term back-propagates through ``fvfi.pyramid.Pyramid.inv_filter`` (fvfi_pyr_reconstruct_backward)."""
import torch


def wrapped_phase_l1(phase_r, phase_g):
    """mean |atan2(sin(g - r), cos(g - r))| (loss.py:15-16)."""
    d = phase_g - phase_r
    return torch.atan2(torch.sin(d), torch.cos(d)).abs().mean()


def get_loss(vals_o, vals_t, output, target, pyr, weighting_factor=0.005):
    """PhaseNet special loss (loss.py:5-25); returns (total, l1 share in %, phase share in %)."""
    phase_loss = 0
    for phase_r, phase_g in zip(vals_o.phase, vals_t.phase):
        if isinstance(phase_r, (int, float)) or isinstance(phase_g, (int, float)):
            continue
        r = phase_r.reshape(-1, pyr.nbands, phase_r.shape[2], phase_r.shape[3])
        g = phase_g.reshape(-1, pyr.nbands, phase_r.shape[2], phase_r.shape[3])
        for b in range(pyr.nbands):                     # one mean per orientation, summed (loss.py:14-16)
            phase_loss = phase_loss + wrapped_phase_l1(r[:, b], g[:, b])
    l_1 = torch.nn.functional.l1_loss(output, target)
    total = l_1 + weighting_factor * phase_loss
    l_1_p = 100 * l_1.detach() / total
    phase_loss_p = 100 * weighting_factor * (phase_loss.detach() if torch.is_tensor(phase_loss) else phase_loss) / total
    return total, l_1_p, phase_loss_p
