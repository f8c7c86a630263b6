"""oracle/wrap_align.py -- TEST INFRASTRUCTURE (see oracle/__init__.py): the +-pi branch aligner of the parity runs."""


class WrapAligner:
    """``FusionPipeline.filter_hook`` for parity runs.

    The reference feeds WRAPPED phases ``imag(log z)`` in (-pi, pi] to PhaseNet (src/train/pyramid.py:63, phase_net.py:66) and
    subtracts wrapped phases in the uncertainty branch (src/train/utils.py:322-346).  Both are discontinuous where a coefficient
    lies on the negative real axis: a rounding-level change of z (1e-7 of the level maximum -- any FFT other than the very build the
    reference ran on) flips its phase between +pi and -pi, an O(1) change of the network input.  tests/test_models_oracle.py shows
    on the reference's own modules that ONE such flip moves the output by 1e-4 .. 5e-3.  For a meaningful comparison the parity
    runs therefore evaluate the GPU pipeline on the reference's branch: this hook decomposes the same planes with the CPU oracle
    and, at the (few) coefficients where the GPU phase and the oracle phase differ by ~2 pi, replaces the GPU phase by the
    oracle's.  Nothing else is touched; ``flips`` counts the replaced values, ``coefficients`` all phase values seen."""

    def __init__(self, height, nbands=4):
        import math
        from oracle import nets
        self.pyr = nets.Pyramid(height, nbands, math.sqrt(2))
        self.flips = 0
        self.coefficients = 0
        self.max_aligned_phase_diff = 0.0

    def __call__(self, tag, planes, vals):
        import torch
        ref = self.pyr.filter(planes.detach().float().cpu())
        phase = list(vals.phase)
        for l, p in enumerate(phase):
            if p is None:
                continue
            r = ref.phase[l].to(p.device)
            d = p - r
            flip = d.abs() > 3.0
            self.flips += int(flip.sum())
            self.coefficients += p.numel()
            phase[l] = torch.where(flip, r, p)
        return vals._replace(phase=phase)
