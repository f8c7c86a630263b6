"""oracle/wrap_align.py -- TEST INFRASTRUCTURE (see oracle/__init__.py): the +-pi branch aligner of the parity runs.

The reference feeds WRAPPED phases ``imag(log z)`` in (-pi, pi] to PhaseNet (src/train/pyramid.py:63, src/phase_net/phase_net.py:66)
and subtracts wrapped phases in the uncertainty branch (src/train/utils.py:322-346).  Both are discontinuous where a coefficient
lies on the negative real axis: a change of z at the level of fp32 rounding (1e-7 of the level maximum -- what ANY FFT other than
the very build the reference ran on does; even the reference's own decomposition of an input that differs by 3e-7) flips the phase
between +pi and -pi, an O(1) change of a network input.  tests/test_models_oracle.py shows on the reference's modules that three
such flips among 2.6 M coefficients move PhaseNet's output by 1.4e-4 (4.5e-3 with the shipped phase_net.pt) while the same
perturbation without the flips moves it by 3e-7.

For a meaningful comparison the parity runs evaluate the GPU pipeline ON THE REFERENCE'S BRANCH: ``WrapAligner`` is installed as
``FusionPipeline.filter_hook``; it knows, per decomposition call of the recipe ("phasenet", "uncertainty"), the reference's phase at
every coefficient within a window of +-pi, and where the GPU phase of such a coefficient lies on the other side of the cut (differs
by more than 3 rad) it is moved by 2 pi onto the reference's side -- the value itself is kept, only the branch changes.  Nothing
else is touched; ``flips`` counts the moved values, ``coefficients`` all phase values seen."""
import math

import numpy as np

# radians from +-pi.  "phasenet": inputs agree to 3e-7, a coefficient that can flip is within rounding of the cut.  "uncertainty":
# the decomposed images are the recipe's own intermediate results (agreeing to ~1e-5), weak coefficients of the coarse levels move
# further; only the six coarsest levels are read there (src/train/utils.py:282-320 with use_levels = 6), so a wide window is cheap.
WINDOW = {"phasenet": 0.05, "uncertainty": 1.0}
UNCERTAINTY_LEVELS = 6


def wrap_lists(decomps):
    """{call: DecompValues} of the reference run -> {"wrap_<call>_<level>_idx": int32 flat indices, "..._val": float32 phases} of
    the coefficients within WINDOW of +-pi (what a fixture stores)."""
    out = {}
    for tag, vals in decomps.items():
        L = len(vals.phase)
        for l, p in enumerate(vals.phase):
            if tag == "uncertainty" and l < L - UNCERTAINTY_LEVELS:
                continue
            flat = np.asarray(p.detach().cpu().numpy(), dtype=np.float32).reshape(-1)
            idx = np.nonzero(np.abs(flat) > math.pi - WINDOW.get(tag, 0.05))[0].astype(np.int32)
            out["wrap_%s_%d_idx" % (tag, l)] = idx
            out["wrap_%s_%d_val" % (tag, l)] = flat[idx]
    return out


class WrapAligner:
    def __init__(self, lists):
        """``lists``: the dict of wrap_lists() (or an np.load'ed fixture holding those keys)."""
        self.lists = lists
        self.flips = 0
        self.coefficients = 0

    @classmethod
    def from_decomps(cls, decomps):
        return cls(wrap_lists(decomps))

    def __call__(self, tag, planes, vals):
        import torch
        keys = self.lists.files if hasattr(self.lists, "files") else self.lists
        phase = list(vals.phase)
        for l, p in enumerate(phase):
            if p is None:
                continue
            self.coefficients += p.numel()
            k = "wrap_%s_%d_idx" % (tag, l)
            if k not in keys:
                continue
            idx = torch.as_tensor(np.asarray(self.lists[k]), dtype=torch.long, device=p.device)
            if idx.numel() == 0:
                continue
            ref = torch.as_tensor(np.asarray(self.lists["wrap_%s_%d_val" % (tag, l)]), dtype=p.dtype, device=p.device)
            flat = p.contiguous().reshape(-1).clone()
            got = flat[idx]
            flip = (got - ref).abs() > 3.0
            self.flips += int(flip.sum())
            flat[idx[flip]] = got[flip] + 2 * math.pi * torch.sign(ref[flip] - got[flip])
            phase[l] = flat.reshape(p.shape)
        return vals._replace(phase=phase)
