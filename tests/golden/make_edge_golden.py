"""Generates tests/golden/edge_golden.npz: fixtures for the two steps either side of the forward
path, produced by the code the reference itself runs:

* `norm_*`   : torchvision.transforms ToTensor + Normalize(IMAGENET_DEFAULT_MEAN/STD) exactly as
               composed by the reference's eval transform (data/get_dataset.py:107-108), applied
               to seeded uint8 HWC images;
* `eval_*`   : the reference's UNMODIFIED engine.evaluate (engine.py:17-45; its own MetricLogger
               and CrossEntropyLoss; timm's accuracy() restated in oracle/ref_shim.py) run over
               seeded (logits, target) batches with a ragged last batch.

Run in the build container (the reference is not present on the GPU box):
    python tests/golden/make_edge_golden.py
"""
import contextlib
import io
import sys
from pathlib import Path

import numpy as np
import torch
from torchvision import transforms

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_shim  # noqa: E402
from devit_b200 import synth  # noqa: E402
from devit_b200.models import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD  # noqa: E402

OUT = Path(__file__).resolve().parent


class _Replay(torch.nn.Module):
    """Returns the pre-computed logits of the batch whose first pixel carries its index."""

    def __init__(self, logits):
        super().__init__()
        self.logits = logits

    def forward(self, images):
        return self.logits[int(images.flatten()[0].item())]


def main():
    out = {}
    # ---- ToTensor + Normalize
    tf = transforms.Compose([transforms.ToTensor(),
                             transforms.Normalize(IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD)])
    u8 = synth.images_u8(2, img=32, nhwc=True)  # HWC like a decoded PIL image
    out['norm_u8_nhwc'] = u8.numpy()
    out['norm_out'] = torch.stack([tf(img.numpy()) for img in u8]).numpy()
    # every byte value once per channel (256 pixels x 3 channels)
    ramp = torch.arange(256, dtype=torch.uint8).view(16, 16, 1).expand(16, 16, 3).contiguous()
    out['norm_ramp_out'] = tf(ramp.numpy()).numpy()

    # ---- engine.evaluate
    engine = ref_shim.load_reference_engine()
    for name, sizes, classes in (('c100', (8, 8, 5), 100), ('c1000', (16, 3), 1000),
                                 ('c3', (4, 4), 3)):
        batches = synth.eval_batches(sizes, classes)
        loader = [(torch.full((b[0].shape[0], 1), float(i)), b[1]) for i, b in enumerate(batches)]
        with contextlib.redirect_stdout(io.StringIO()):
            res = engine.evaluate(loader, _Replay([b[0] for b in batches]), torch.device('cpu'))
        out[f'eval_{name}'] = np.array([res['loss'], res['acc1'], res['acc5']], dtype=np.float64)
        print(name, res)
    np.savez_compressed(OUT / 'edge_golden.npz', **out)
    print('wrote', OUT / 'edge_golden.npz')


if __name__ == '__main__':
    main()
