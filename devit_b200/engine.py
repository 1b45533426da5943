"""Evaluation loops around the forward path: drop-ins for ``engine.evaluate`` (engine.py:17-45)
and ``engine.evaluate_ens_disjoint`` (engine.py:212-243) of the reference.

What changes against the reference loop:

* the per-batch tail -- CrossEntropyLoss + timm ``accuracy(topk=(1, 5))`` followed by three
  ``.item()`` host synchronisations (engine.py:34-41 / 229-238) -- is one device call
  (``devit_eval_tail``) that adds into five float64 meters living in HBM; the host reads them
  once, after the last batch;
* a data loader may hand over *decoded uint8* batches ([B,C,H,W] or [B,H,W,3]); ToTensor +
  Normalize (data/get_dataset.py:107-108) then run inside the patch extraction on the device
  (``devit_im2col_tokens_u8``), so the host->device copy moves a quarter of the bytes.  fp32
  batches (the reference's loader output) keep working unchanged.

The returned dict has the reference's keys and meaning: ``loss`` = mean over batches of the batch
mean loss (MetricLogger.update(loss=...) with n = 1), ``acc1`` / ``acc5`` = percentages weighted
by batch size (meters updated with n = batch_size), summed over ranks when torch.distributed is
initialised (utils/dist_utils.py:35-46).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib as L


class EvalMeters:
    """Device-resident ``loss`` / ``acc1`` / ``acc5`` meters (the three SmoothedValue meters the
    reference's MetricLogger keeps, utils/dist_utils.py:16-60, reduced to total and count)."""

    def __init__(self, device, topk: int = 5):
        self.topk = topk
        # {sum of batch-mean losses, #batches, #correct@1, #correct@k, #samples}
        self.acc = torch.zeros(5, device=device, dtype=torch.float64)

    def update(self, logits: torch.Tensor, target: torch.Tensor, want_batch: bool = False):
        """Adds one batch; no host synchronisation.  Returns the batch's float32
        {mean loss, #correct@1, #correct@k} on the device when `want_batch`."""
        return L.eval_tail(logits.float(), target, self.acc, self.topk, want_batch)

    def synchronize_between_processes(self):
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
            dist.all_reduce(self.acc)

    def result(self) -> dict:
        """One device->host read; the reference's ``{k: meter.global_avg}`` dict."""
        return meters_to_dict(self.acc.tolist())


def meters_to_dict(acc) -> dict:
    loss_total, batches, c1, ck, samples = (float(v) for v in acc)
    if batches == 0 or samples == 0:
        return {}
    return {'loss': loss_total / batches, 'acc1': 100.0 * c1 / samples,
            'acc5': 100.0 * ck / samples}


def _to_device(images, target, device):
    return images.to(device, non_blocking=True), target.to(device, non_blocking=True)


@torch.no_grad()
def evaluate(data_loader, model, device):
    """engine.py:17-45 for one (sub-)model."""
    model.eval()
    meters = EvalMeters(device)
    for images, target in data_loader:
        images, target = _to_device(images, target, device)
        meters.update(model(images), target)
    meters.synchronize_between_processes()
    return meters.result()


@torch.no_grad()
def evaluate_ens_disjoint(data_loader, model, ens_model, device):
    """engine.py:212-243: MultiViT features -> EnsMLP fusion head -> loss / top-1 / top-5."""
    model.eval()
    ens_model.eval()
    meters = EvalMeters(device)
    for images, target in data_loader:
        images, target = _to_device(images, target, device)
        meters.update(ens_model(model(images)), target)
    meters.synchronize_between_processes()
    return meters.result()
