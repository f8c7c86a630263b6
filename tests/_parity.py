"""Stage-by-stage parity report against the round-2 golden fixtures (tests/golden/make_golden_r2.py).

Criterion per stage k (north star: max abs error <= 1e-4 on [0,1] images against the reference's own PyTorch run):
    |GPU - reference_fp32| <= 1e-4
 or |GPU - fp64| <= 2 * max|reference_fp32 - fp64|        (the fp64 ARBITER: the GPU result is no further from the exact result of the
                                                          recipe than the reference's own fp32 CPU run is -- the only meaningful bound
                                                          for stages the recipe amplifies by 100 / 150 before a clamp,
                                                          src/fusion_net/interpolate_twoframe.py:211,220,224, and for the phase of
                                                          coefficients whose amplitude is rounding noise)
Phases are compared as complex coefficients amp * exp(i phase): the angle of a (near-)zero coefficient is arbitrary.
"""
import numpy as np

TOL = 1e-4


def _sample(a, z, k):
    st = int(z[k + "__stride"]) if (k + "__stride") in z.files else 1
    return a[..., ::st, ::st] if st > 1 else a


def _f64(z, k):
    return z[k].astype(np.float64) - z[k + "__d64"].astype(np.float64) / 1e4


def stage_report(z, stages):
    """z: np.load of a fixture; stages: {name: torch tensor (any device)}.  -> {name: dict(err_ref, err_f64, budget, ok)}"""
    rep = {}
    names = [k for k in z.files if "__" not in k and k not in ("meta", "checksum")]
    get = lambda k: _sample(stages[k].detach().float().cpu().numpy(), z, k)
    for k in names:
        if k not in stages:
            continue
        if k.startswith("amp") and k[3:].isdigit():
            continue                                            # handled with its phase
        if k.startswith("phase") and k[5:].isdigit():
            l = k[5:]
            ph, am = get(k).astype(np.float64), get("amp" + l).astype(np.float64)
            g = am * np.exp(1j * ph)
            r = z["amp" + l].astype(np.float64) * np.exp(1j * z[k].astype(np.float64))
            f = _f64(z, "amp" + l) * np.exp(1j * _f64(z, k))
            budget = float(np.abs(r - f).max())
            name = "coeff" + l
        else:
            g, r, f = get(k).astype(np.float64), z[k].astype(np.float64), _f64(z, k)
            budget = float(z[k + "__budget"])
            name = k
        err_ref, err_f64 = float(np.abs(g - r).max()), float(np.abs(g - f).max())
        rep[name] = dict(err_ref=err_ref, err_f64=err_f64, budget=budget,
                         ok=bool(err_ref <= TOL or err_f64 <= 2 * budget + 1e-7))
    return rep


def fmt(rep):
    return ", ".join("%s %.1e(f64 %.1e/b %.1e)%s" % (k, v["err_ref"], v["err_f64"], v["budget"], "" if v["ok"] else " FAIL")
                     for k, v in rep.items())


def psnr(a, b):
    mse = float(((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2).mean())
    return 10 * np.log10(1.0 / max(mse, 1e-20))


from oracle.wrap_align import WrapAligner  # noqa: E402,F401  (checker infrastructure shared with bench.py's parity block)
