"""FusionNet training step -- mirror of ``Trainer.predict`` + ``Trainer.train`` in the reference's
``src/fusion_net/trainer.py`` (:65-220 predict under no_grad, :246-259 L1 loss / backward / Adam step),
data-parallel over ranks with one flat-bucket gradient all-reduce (fvfi.dist)."""
import torch

from .dist import FlatGradBucket


class FusionTrainer:
    def __init__(self, pipeline, lr=1e-4, group=None):
        self.pipe = pipeline
        self.net = pipeline.fusion_net
        self.net.train()
        for n, p in self.net.named_parameters():
            p.requires_grad_(not n.startswith("net."))          # dead weights (fusion_net.py:11-20)
        self.bucket = FlatGradBucket(self.net.live_parameters())
        self.optimizer = torch.optim.Adam(self.net.live_parameters(), lr=lr)   # trainer.py:46
        self.group = group

    def step(self, rgb1, rgb2, target):
        """One optimisation step on this rank's shard; returns the (local) loss tensor."""
        with torch.no_grad():                                    # frozen PhaseNet + AdaCoF (trainer.py:68-159)
            inputs = self.pipe.fusion_inputs(rgb1, rgb2)
        self.bucket.zero()
        pred = self.net(*inputs, variant=0)                      # trainer.py:215
        loss = torch.nn.functional.l1_loss(target, torch.clip(pred, 0, 1))     # trainer.py:246-254
        loss.backward()                                          # grads land in the flat bucket
        self.bucket.all_reduce_mean(self.group)
        self.optimizer.step()                                    # trainer.py:257-259
        return loss.detach()
