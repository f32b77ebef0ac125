"""One data-parallel training step of aadensenet121, the way chexpert.py:152-165 runs it, on synthetic
radiograph-shaped batches.

    model(x) -> BCE-with-logits on the 5 competition classes, sum over classes, mean over batch (chexpert.py:160,530)
    -> backward -> [bucketed all-reduce over ranks] -> SGD(momentum 0.9, nesterov) + MultiStepLR (chexpert.py:479-480)

The loop around it (tqdm, tensorboard, checkpoint tracker, periodic eval; chexpert.py:167-193) is outside the hot
path and not rebuilt; `TrainStep` is what bench.py times and what a user's own loop would call per batch.
"""
import torch
import torch.distributed as dist

from .dataparallel import GradientBuckets
from .densenet import aadensenet121
from .loss import BCEWithLogitsLoss

# dataset normalisation constants of the reference transform (chexpert.py:70-72)
PIXEL_MEAN, PIXEL_STD = 0.5330, 0.0349


def synthetic_batch(batch, size=320, seed=0, device='cpu', raw_labels=True):
    """Radiograph-shaped synthetic input: one-channel U[0,1] image normalised and replicated to 3 channels
    (chexpert.py:64-72), labels either raw 14-wide rows in {nan,-1,0,1} (dataset.py:20-23) or (B,5) {0,1} targets."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(batch, 1, size, size, generator=g)
    x = ((u - PIXEL_MEAN) / PIXEL_STD).expand(-1, 3, -1, -1).contiguous()
    if raw_labels:
        vals = torch.tensor([float('nan'), -1.0, 0.0, 1.0])
        idx = torch.multinomial(torch.tensor([0.55, 0.10, 0.15, 0.20]), batch * 14, replacement=True, generator=g)
        t = vals[idx].reshape(batch, 14)
    else:
        t = (torch.rand(batch, 5, generator=g) < 0.3).float()
    return x.to(device), t.to(device)


class TrainStep:
    """Owns model, loss kernel, optimizer, scheduler and (when world > 1) the gradient buckets."""

    def __init__(self, device, size=320, precision='bf16', lr=1e-4, raw_labels=True, bucket_mb=25.0, seed=0,
                 autocast=True, channels_last=False):
        torch.manual_seed(seed)
        self.device = torch.device(device)
        self.model = aadensenet121(5, (size, size), precision=precision).to(self.device)
        if channels_last:
            self.model = self.model.to(memory_format=torch.channels_last)
        self.loss_fn = BCEWithLogitsLoss('train', raw_labels=raw_labels).to(self.device)
        self.opt = torch.optim.SGD(self.model.parameters(), lr=lr, momentum=0.9, nesterov=True)
        self.sched = torch.optim.lr_scheduler.MultiStepLR(self.opt, [40000, 60000])
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.buckets = GradientBuckets(self.model, bucket_mb=bucket_mb) if self.world > 1 else None
        # the dense blocks are torch/cuDNN (outside the hot path); bf16 autocast only decides THEIR arithmetic
        self.autocast = bool(autocast and precision == 'bf16' and self.device.type == 'cuda')
        self.model.train()

    def __call__(self, x, target):
        """-> loss (0-dim device tensor; no host sync, unlike loss.item() at chexpert.py:167)."""
        if self.buckets is not None:
            self.buckets.reset()
        else:
            self.opt.zero_grad(set_to_none=True)
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.autocast):
            out = self.model(x)
        loss = self.loss_fn(out.float(), target)
        loss.backward()
        if self.buckets is not None:
            self.buckets.finish()
        self.opt.step()
        self.sched.step()
        return loss.detach()
