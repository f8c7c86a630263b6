"""Model / pipeline golden fixtures from the REAL reference modules (see make_golden.py)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import adacof as oa, fusion_pipeline as fp, nets, ref_import  # noqa: E402

torch.set_grad_enabled(False)


class _CpuWarp:
    """Stands in for FunctionAdaCoF.apply on CPU (the reference raises NotImplementedError there,
    adacof.py:356-357): the C oracle, which is pinned against the reference's own CUDA kernels."""

    @staticmethod
    def apply(inp, w, a, b, dilation):
        return torch.from_numpy(oa.forward(inp.numpy(), w.numpy(), a.numpy(), b.numpy(), dilation, threads=8))


def reference_backend(state, H, W):
    ref_import.install_stubs()
    from src.train.pyramid import Pyramid
    from src.phase_net.phase_net import PhaseNet
    from src.fusion_net.fusion_net import FusionNet
    import src.fusion_net.fusion_adacofnet as fa
    from src.train import utils as ru
    cpu = torch.device("cpu")
    pyr = Pyramid(height=ru.calc_pyr_height(torch.empty(3, H, W)), nbands=4, scale_factor=np.sqrt(2), device=cpu)
    pn = PhaseNet(pyr, cpu, 2).eval()
    pn.load_state_dict(state["phase_net"])
    fn = FusionNet().eval()
    fn.load_state_dict(state["fusion_net"])
    an = fa.AdaCoFNet(types.SimpleNamespace(kernel_size=5, dilation=1, gpu_id=0)).eval()
    an.load_state_dict(state["adacof"])
    an.moduleAdaCoF = _CpuWarp.apply
    return types.SimpleNamespace(pyr=pyr, phase_net=pn, fusion_net=fn, adacof=an, separate_vals=ru.separate_vals,
                                 get_concat_layers_inf=ru.get_concat_layers_inf,
                                 get_last_value_levels=ru.get_last_value_levels,
                                 get_first_value_levels=ru.get_first_value_levels,
                                 subtract_values=ru.subtract_values)


def make_models():
    for (B, H, W, seed) in [(1, 64, 64, 0), (2, 64, 96, 1)]:
        state = fp.seeded_state(seed)
        rgb1, rgb2 = fp.seeded_frames(B, H, W, seed)
        be = reference_backend(state, H, W)
        st = {}
        final = fp.interp(be, rgb1, rgb2, st)
        # the oracle restatement must agree with the real reference modules on the same weights
        ob = fp.oracle_backend(state, hw=(H, W), threads=8)
        so = {}
        fo = fp.interp(ob, rgb1, rgb2, so)
        err = {k: float((st[k] - so[k]).abs().max()) for k in st}
        print("oracle-vs-reference max abs diff per stage:", {k: "%.1e" % v for k, v in err.items()})
        assert max(err.values()) < 5e-6, err
        # op-level fixtures: PhaseNet on the first pyramid call, FusionNet on the final inputs
        name = "pipeline_ref_B%d_%dx%d_s%d.npz" % (B, H, W, seed)
        keep = {k: v.numpy().astype(np.float32) for k, v in st.items()}
        keep["meta"] = np.array([B, H, W, seed])
        keep["checksum"] = np.array([float(sum(v.double().sum() for v in state[n].values())) for n in
                                     ("phase_net", "fusion_net", "adacof")])
        np.savez_compressed(os.path.join(HERE, name), **keep)
        print("wrote", name, {k: v.shape for k, v in keep.items() if k not in ("meta", "checksum")})


if __name__ == "__main__":
    make_models()
