"""FusionNet training step -- mirror of ``Trainer.predict`` + ``Trainer.train`` in the reference's
``src/fusion_net/trainer.py`` (:65-220 predict under no_grad, :246-259 L1 loss / backward / Adam step),
data-parallel over ranks with one flat-bucket gradient all-reduce (fvfi.dist)."""
import torch

from . import conv as tc
from .dist import FlatGradBucket, all_reduce_mean_buffers, broadcast_module


class FusionTrainer:
    def __init__(self, pipeline, lr=1e-4, group=None, graph_frozen=False):
        self.pipe = pipeline
        self.graph_frozen = graph_frozen        # replay the frozen PhaseNet / AdaCoF part as a CUDA graph (fixed crop shape)
        self.check_range = False                # True: read the 3xFP16 range flag after every step (one stream sync per step)
        self.net = pipeline.fusion_net
        self.net.train()
        # replicas start from rank 0's weights: the trained FusionNet and the frozen networks that produce its inputs
        for m in (self.net, pipeline.phase_net, pipeline.adacof):
            broadcast_module(m, 0, group)
        for n, p in self.net.named_parameters():
            p.requires_grad_(not n.startswith("net."))          # dead weights (fusion_net.py:11-20)
        self.bucket = FlatGradBucket(self.net.live_parameters())
        self.optimizer = torch.optim.Adam(self.net.live_parameters(), lr=lr)   # trainer.py:46
        self.group = group

    def step(self, rgb1, rgb2, target):
        """One optimisation step on this rank's shard; returns the (local) loss tensor."""
        with torch.no_grad():                                    # frozen PhaseNet + AdaCoF (trainer.py:68-159)
            if self.graph_frozen:
                inputs = self.pipe.graphed("fusion_inputs", rgb1, rgb2)(rgb1, rgb2)
            else:
                inputs = self.pipe.fusion_inputs(rgb1, rgb2)
        self.bucket.zero()
        pred = self.net(*inputs, variant=0)                      # trainer.py:215
        loss = torch.nn.functional.l1_loss(target, torch.clip(pred, 0, 1))     # trainer.py:246-254
        loss.backward()                                          # grads land in the flat bucket
        self.bucket.all_reduce_mean(self.group)
        self.optimizer.step()                                    # trainer.py:257-259
        if self.check_range and tc.overflow_pending():           # FusionNet's training forward runs on the 3xFP16 kernels too
            raise FloatingPointError("FusionTrainer: activation beyond the 3xFP16 range; use fvfi.conv.forced_precision('tf32x3')")
        return loss.detach()


class PhaseNetTrainer:
    """PhaseNet training step -- mirror of ``Trainer.predict`` + ``Trainer.train`` of the reference's ``src/train/trainer.py``
    in ``mode == 'phase'`` (:65-165): decompose frame 1, frame 2 and the target (Lab planes), PhaseNet on the two input
    pyramids, reconstruct, PhaseNet loss (src/train/loss.py) = L1 on the image + wrapped phase difference per level, Adam.
    The decomposition needs no gradient; the reconstruction back-propagates through fvfi_pyr_reconstruct_backward; the network's
    convolutions run forward and backward on libfvfi (conv._ConvTC), BatchNorm's batch statistics are torch's.  Data parallel like FusionTrainer (one flat gradient bucket)."""

    def __init__(self, pyr, phase_net, lr=1e-3, weight_decay=0.0, group=None):
        from .dist import FlatGradBucket
        self.pyr, self.net, self.group = pyr, phase_net, group
        self.net.train()
        broadcast_module(self.net, 0, group)                                             # parameters + BatchNorm statistics
        params = [p for p in self.net.parameters() if p.requires_grad]
        self.bucket = FlatGradBucket(params)
        self.optimizer = torch.optim.Adam(params, lr=lr, weight_decay=weight_decay)       # trainer.py:40-42

    def predict(self, lab1, lab2, target, m=None):
        """lab1 / lab2 / target: [P,H,W] Lab planes.  Returns (prediction [P,H,W], vals_pred, vals_target)."""
        from . import utils
        P = lab1.shape[0]
        with torch.no_grad():                                                            # inputs carry no gradient
            vals = self.pyr.filter(torch.cat((lab1, lab2, target), 0))
            v1, v2, vt = utils.separate_vals(vals, 3)
            vals_in = self.net.normalize_vals(utils.get_concat_layers_inf(self.pyr, [v1, v2]))
        vals_pred = self.net(vals_in, m)
        prediction = self.pyr.inv_filter(vals_pred)
        return prediction, vals_pred, vt

    def step(self, lab1, lab2, target, m=None):
        from .loss import get_loss
        self.bucket.zero()
        prediction, vals_pred, vals_target = self.predict(lab1, lab2, target, m)
        loss, p1, p2 = get_loss(vals_pred, vals_target, prediction, target, self.pyr)      # trainer.py:127-128
        loss.backward()
        self.bucket.all_reduce_mean(self.group)
        self.optimizer.step()
        all_reduce_mean_buffers(self.net, self.group)          # BatchNorm running statistics stay identical across replicas
        return loss.detach()
