"""Ensemble wrappers: MultiViT (N decomposed sub-models) and the EnsMLP feature-fusion head,
drop-in for models/ensemble_models.py:13-90.

Data layout.  The reference builds ``torch.stack(list, 1).view(B, -1)`` (sub-model-major within
a sample) and feeds two Linears (models/ensemble_models.py:76-85).  Here every sub-model's
final-norm kernel writes its cls / dist rows straight into one slab

        feats[s, j, b, :]      s = sub-model, j = 0 (cls) | 1 (dist), b = image

and the first fusion GEMM walks the slab as K-segments: sub-model s contributes
A = feats[s, j] ([B, D]) against weight columns [s*D, (s+1)*D) -- the same sum as the stacked
matmul, without the stack/transpose copies.  With one sub-model per GPU the slab is exactly what
an all-gather of the per-rank [2, B, D] blocks produces (devit_b200/parallel.py).
"""
from __future__ import annotations

import contextlib
import os

import torch
import torch.nn as nn

from . import _lib as L
from . import models  # noqa: F401  (registers the entrypoints)
from .models import _PREC, default_precision
from .registry import create_model


def sub_streams() -> int:
    """Number of CUDA streams MultiViT spreads a rank's kernel chains over (DEVIT_SUB_STREAMS)."""
    return int(os.environ.get('DEVIT_SUB_STREAMS', '4'))


def collapse_head() -> bool:
    """EnsMLP evaluates its two Linears per token kind as one pre-multiplied GEMM (see
    EnsMLP._collapsed); DEVIT_COLLAPSE_HEAD=0 keeps the four separate GEMMs."""
    return os.environ.get('DEVIT_COLLAPSE_HEAD', '1') != '0'


def batch_chunks(n_sub_local: int, batch: int) -> int:
    """How many chunks the batch is cut into.  Default 1: with the current kernels a rank's
    sub-models are enough independent chains, and cutting one sub-model's batch only shrinks its
    grids (measured, profiles/r2_time_chain_v2.txt: 1 sub-model x 256 images 2.12 ms as one chain,
    2.20 ms as 4 x 64).  DEVIT_CHAINS=<n> asks for about n chains per rank (chunks never smaller
    than DEVIT_MIN_CHUNK images, default 64) for experiments."""
    want = int(os.environ.get('DEVIT_CHAINS', '0'))
    if want <= 0:
        return 1
    min_chunk = max(1, int(os.environ.get('DEVIT_MIN_CHUNK', '64')))
    n = max(1, -(-want // max(1, n_sub_local)))
    return max(1, min(n, batch // min_chunk))


def chain_sm_budget(device, n_chains: int = 4) -> int:
    """SMs each concurrent chain's grids are sized for: half the chip when at least four chains
    run (two at a time side by side), the whole chip for two or three (measured: 2 sub-models
    3.87 ms with whole-chip grids on two streams, 4.10 ms on half-chip grids; 4 sub-models 8.08 vs
    7.79 ms).  DEVIT_SM_SHARE=<d> forces chip / d."""
    share = int(os.environ.get('DEVIT_SM_SHARE', '0'))
    if share <= 0:
        share = 2 if n_chains >= 4 else 1
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    return 0 if share == 1 else max(2, (sms // share) & ~1)


_SIDE_STREAMS = {}


def _side_streams(device, n):
    """Side streams of the CURRENT stream: every parent stream gets its own set, so two forwards
    running concurrently on two streams (parallel.BatchPipeline slots) never share a side stream
    -- and with it a workspace, which packing.workspace keys by stream."""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    pool = _SIDE_STREAMS.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


def chain_tasks(subs, batch, on_cuda=True):
    """[(slab index, sub-model id, (b0, b1) | None)]: the (sub-model, batch chunk) chains of one
    forward of the sub-models `subs` over `batch` images."""
    n_chunk = batch_chunks(len(subs), batch) if on_cuda and not L.profiling() else 1
    if n_chunk <= 1:
        return [(i, s, None) for i, s in enumerate(subs)]
    per = batch // n_chunk
    cuts = [c * per for c in range(n_chunk)] + [batch]
    return [(i, s, (cuts[c], cuts[c + 1])) for c in range(n_chunk) for i, s in enumerate(subs)]


def run_chains(device, tasks, launch):
    """Runs launch(task) for every task, spreading the tasks round-robin over up to
    `sub_streams()` CUDA streams forked from / joined to the current stream.

    The sub-models are independent until the fusion head (models/ensemble_models.py:33 is a plain
    loop) and so are the images of a batch: each task is an independent kernel chain with its own
    workspace (packing.workspace is keyed by stream), and while more than one stream is in use
    every persistent grid is sized for HALF the chip (devit_set_sm_budget), so two chains run side
    by side.  Each kernel alternates tensor-bound and HBM-bound phases (operand refill, fp32
    residual read / write) that do not overlap inside one CTA; two unrelated kernels on the two
    halves of the chip fill each other's gaps, and a chain's partly filled last round of tiles no
    longer idles the SMs it does not use.  Measured on the 4-way bs-256 step: 9.93 -> 8.7 ms.
    Results are bit-identical to the single-chain order (every row's arithmetic is unchanged).
    The SM budget is a process-wide setting of the library: forwards issued from several host
    threads at once would see each other's budget (a performance matter only, never correctness)."""
    on_cuda = device.type == 'cuda'
    n_st = max(1, min(len(tasks), sub_streams())) if on_cuda and not L.profiling() else 1
    if n_st == 1:
        for t in tasks:
            launch(t)
        return
    cur = torch.cuda.current_stream(device)
    side = _side_streams(device, n_st - 1)
    fork = torch.cuda.Event()
    fork.record(cur)
    for st in side:
        st.wait_event(fork)
    prev_budget = L.load().devit_set_sm_budget(chain_sm_budget(device, len(tasks)))
    try:
        for k, t in enumerate(tasks):
            st = cur if k % n_st == 0 else side[k % n_st - 1]
            with torch.cuda.stream(st):
                launch(t)
    finally:
        L.load().devit_set_sm_budget(prev_budget)
    for st in side:  # join: whatever the tasks used is next touched (and freed) on `cur`
        cur.wait_stream(st)


class FeatureList(list):
    """A plain list of [B, D] tensors (what the reference's MultiViT returns) that remembers
    the slab its entries are views of, so EnsMLP can skip the stack."""
    slab_f32 = None   # [n_sub, n_kind, B, D] fp32
    slab_op = None    # same in GEMM-operand format (bf16, or [2, n_sub, n_kind, B, D] split)
    kind = 0          # 0 = cls, 1 = dist
    slab_version = -1  # slab_f32._version when the list was built (in-place edits bump it)


class MultiViT(nn.Module):
    """models/ensemble_models.py:13-40."""

    def __init__(self, model='dedevit', drop=0, drop_path=0.1, num_classes_list=[25, 25, 25, 25],
                 num_div=4):
        super().__init__()
        self.model = model
        assert len(num_classes_list) == num_div, 'num of classes is not match num of sub-models'
        self.backbones = nn.ModuleList([])
        for i, num_class in enumerate(num_classes_list):
            self.backbones.append(create_model(model_name=self.model, num_classes=int(num_class),
                                               drop_rate=drop, drop_path_rate=drop_path,
                                               drop_block_rate=None))
            del self.backbones[i].head
            if 'deit' in self.model:
                del self.backbones[i].head_dist
        self.precision = default_precision()

    def set_precision(self, precision: str):
        self.precision = precision
        for b in self.backbones:
            b.set_precision(precision)
        return self

    def set_input_norm(self, mean, std):
        """Device-side ToTensor + Normalize for uint8 batches (see VisionTransformer)."""
        for b in self.backbones:
            b.set_input_norm(mean, std)
        return self

    @torch.no_grad()
    def forward_slab(self, x, subs=None):
        """Runs the sub-models in `subs` (default: all) and returns (slab_f32, slab_op) with
        shape [len(subs), n_kind, B, D] (slab_op: bf16, or [2, ...] split planes in fp32 mode)."""
        subs = list(range(len(self.backbones))) if subs is None else list(subs)
        bb0 = self.backbones[subs[0]]
        prec = _PREC[self.precision]
        B, D, T = x.shape[0], bb0.embed_dim, bb0.num_tokens
        f32 = torch.empty(len(subs), T, B, D, device=x.device)
        if prec == L.DEVIT_BF16:
            op = torch.empty(len(subs), T, B, D, device=x.device, dtype=torch.bfloat16)
        else:
            op = torch.empty(2, len(subs), T, B, D, device=x.device)
        # every sub-model embeds the SAME images (models/ensemble_models.py:33): extract the
        # patch matrix once when several of them run here and share the patch geometry
        patches = None
        if (len(subs) > 1 or x.dtype == torch.uint8) and x.is_cuda and x.dim() == 4:
            geo = {(bb.patch_embed.img_size, bb.patch_embed.proj.weight.shape[1:], bb.num_tokens,
                    bb.input_norm) for bb in (self.backbones[s] for s in subs)}
            if len(geo) == 1:
                bb0._check_input(x, convert=False)
                if bb0.precision != self.precision:
                    bb0.set_precision(self.precision)
                patches = bb0.patches_of(x)
        # independent (sub-model, batch chunk) kernel chains on several streams: see run_chains
        def launch(task):
            i, s, rows = task
            bb = self.backbones[s]
            if bb.precision != self.precision:
                bb.set_precision(self.precision)
            bb.features_into(x, feats_f32=f32[i],
                             feats_op=op[i] if prec == L.DEVIT_BF16 else op[:, i],
                             patches=patches, rows=rows)

        run_chains(x.device, chain_tasks(subs, B, x.is_cuda), launch)
        return f32, op


    def forward(self, x):
        f32, op = self.forward_slab(x)
        n = f32.shape[0]
        if 'vit' in self.model:  # e.g. 'devit' (note: 'vit' is not a substring of 'dedeit')
            out = FeatureList(f32[s, 0] for s in range(n))
            out.slab_f32, out.slab_op, out.kind, out.slab_version = f32, op, 0, f32._version
            return out
        cls = FeatureList(f32[s, 0] for s in range(n))
        dist = FeatureList(f32[s, 1] for s in range(n))
        cls.slab_f32, cls.slab_op, cls.kind, cls.slab_version = f32, op, 0, f32._version
        dist.slab_f32, dist.slab_op, dist.kind, dist.slab_version = f32, op, 1, f32._version
        return cls, dist


class EnsMLP(nn.Module):
    """models/ensemble_models.py:43-90: per token kind, Linear(n*sub_size -> teacher_size) then
    Linear(teacher_size -> num_class) (no activation between them), logits = (cls + dist) / 2."""

    def __init__(self, model='dedevit', num_class=100, sub_size=192,
                 num_classes_list=[25, 25, 25, 25], teacher_size=None):
        super().__init__()
        self.model = model
        self.sub_size = sub_size
        self.teacher_size = teacher_size
        self.num_classes = num_class
        self.num_sub = len(num_classes_list)
        self.sum_feature_dim = self.sub_size * self.num_sub
        if self.teacher_size is None:
            self.cls_classifier = nn.Linear(self.sum_feature_dim, self.num_classes)
            if 'deit' in self.model:
                self.dist_classifier = nn.Linear(self.sum_feature_dim, self.num_classes)
        else:
            self.cls_mlp = nn.Linear(self.sum_feature_dim, self.teacher_size)
            self.cls_classifier = nn.Linear(self.teacher_size, self.num_classes)
            if 'deit' in self.model:
                self.dist_mlp = nn.Linear(self.sum_feature_dim, self.teacher_size)
                self.dist_classifier = nn.Linear(self.teacher_size, self.num_classes)
        self.precision = default_precision()
        self._packs = None

    def set_precision(self, precision: str):
        self.precision = precision
        return self

    # ------------------------------------------------------------------ packed weights
    def _packed(self, device):
        from . import packing
        ver = (self.precision, str(device), packing.module_version(self))
        if self._packs is None or self._packs[0] != ver:
            prec = _PREC[self.precision]
            pk = {name: packing.PackedLinear(m, prec, device)
                  for name, m in self.named_children() if isinstance(m, nn.Linear)}
            self._packs = (ver, pk)
        return self._packs[1]

    def _slab_of(self, lst, device):
        """Operand slab [n, kinds, B, D] for a list of [B, D] tensors (+ which kind to read)."""
        prec = _PREC[self.precision]
        if isinstance(lst, FeatureList) and lst.slab_op is not None:
            # the slab is a second copy of the features: use it only while every list entry still
            # IS the slab's view (same storage, untouched since) and the operand format matches
            # this head's precision; otherwise stack the tensors like the reference does
            f32, op = lst.slab_f32, lst.slab_op
            fresh = len(lst) == f32.shape[0] and all(
                torch.is_tensor(t) and t.data_ptr() == f32[j, lst.kind].data_ptr()
                for j, t in enumerate(lst)) and f32._version == lst.slab_version
            fmt_ok = (op.dim() == 4 and op.dtype == torch.bfloat16) if prec == L.DEVIT_BF16 \
                else (op.dim() == 5 and op.dtype == torch.float32)
            if fresh and fmt_ok:
                return op, lst.kind
        if not lst[0].is_cuda:
            raise L.DevitError("EnsMLP runs on CUDA (sm_100) tensors only; no CPU fallback")
        stacked = torch.stack([t.float() for t in lst], 0).unsqueeze(1).contiguous()  # [n,1,B,D]
        if prec == L.DEVIT_BF16:
            return stacked.to(torch.bfloat16), 0
        return L.split_tf32(stacked), 0

    def _fuse(self, slab, kind, first: str, out_kind, order=None, **epi):
        """GEMM over the slab with one K-segment per sub-model (weight = Linear `first`).
        order[j] = sub-model id held by slab entry j (identity when None)."""
        prec = _PREC[self.precision]
        pk = self._packed(slab.device)[first]
        s4 = slab if prec == L.DEVIT_BF16 else slab[0]
        n, kinds, B, D = s4.shape
        order = list(range(n)) if order is None else list(order)
        if sorted(order) != list(range(n)):
            raise L.DevitError(f"EnsMLP: slab order {order} is not a permutation of 0..{n - 1}")
        if n * D != pk.in_features:
            raise L.DevitError(f"EnsMLP: {n} sub-models x {D} features != Linear in_features "
                               f"{pk.in_features} (reference passes sub_size from a stale table, "
                               f"ensemble.py:224; use sub_size={D})")
        if n > 8:
            raise L.DevitError("EnsMLP: at most 8 sub-models per fusion GEMM")
        a = slab.view(n * kinds * B, D) if prec == L.DEVIT_BF16 else slab.view(2, n * kinds * B, D)
        # segments are issued in SUB-MODEL order whatever the slab layout, so the fp32
        # accumulation order -- hence every logit bit -- is the same for any world size
        segs = sorted(((j * kinds + kind) * B, 0, order[j] * D, D) for j in range(n))
        segs.sort(key=lambda sg: sg[2])
        return L.gemm(a, pk.w, precision=prec, m=B, segs=segs, bias=pk.b, out_kind=out_kind,
                      tag=6, **epi)

    def _collapsed(self, device):
        """Eval-time algebra: the two Linears of a token kind have NO activation between them
        (models/ensemble_models.py:79-84), so logits = (Wc fc + Wd fd + b) / 2 with
        Wk = classifier_k.weight @ mlp_k.weight (computed in fp64 here), b = sum_k
        (classifier_k.weight @ mlp_k.bias + classifier_k.bias).  The whole head becomes ONE
        K-segmented GEMM over the gathered slab (two when 2 n > 8 segments) instead of four
        launches, and the 768-wide intermediate is never rounded."""
        from . import packing
        ver = (self.precision, str(device), packing.module_version(self))
        hit = self.__dict__.get('_collapsed_pack')
        if hit is None or hit[0] != ver:
            prec = _PREC[self.precision]
            ws, b = [], 0
            for kind in ('cls', 'dist'):
                cl = getattr(self, f'{kind}_classifier')
                w, bb = cl.weight.detach().double(), cl.bias.detach().double()
                if self.teacher_size is not None:
                    ml = getattr(self, f'{kind}_mlp')
                    bb = w @ ml.bias.detach().double() + bb
                    w = w @ ml.weight.detach().double()
                ws.append(w)
                b = b + bb
            wcat = torch.cat(ws, 1).float().to(device)                    # [C, 2 n D]
            hit = (ver, (L.to_operand(wcat, prec), b.float().to(device).contiguous(),
                         [L.to_operand(w.float().to(device), prec) for w in ws],
                         [(bb_.float().to(device).contiguous()) for bb_ in (b * 0, b)]))
            self.__dict__['_collapsed_pack'] = hit
        return hit[1]

    def _head_collapsed(self, slab, order):
        """logits from an operand slab [n, 2, B, D] holding both token kinds."""
        prec = _PREC[self.precision]
        s4 = slab if prec == L.DEVIT_BF16 else slab[0]
        n, kinds, B, D = s4.shape
        order = list(range(n)) if order is None else list(order)
        if sorted(order) != list(range(n)) or kinds != 2 or n * D != self.sum_feature_dim:
            raise L.DevitError(f"EnsMLP: slab {tuple(s4.shape)} / order {order} does not match "
                               f"{self.num_sub} sub-models x {self.sub_size} features")
        wcat, bias, wk, bk = self._collapsed(slab.device)
        a = slab.view(n * kinds * B, D) if prec == L.DEVIT_BF16 else slab.view(2, n * kinds * B, D)
        if 2 * n <= 8:
            segs = sorted((((j * kinds + k) * B, 0, (k * n + order[j]) * D, D)
                           for k in range(2) for j in range(n)), key=lambda sg: sg[2])
            return L.gemm(a, wcat, precision=prec, m=B, segs=segs, bias=bias, out_kind=L.OUT_F32,
                          alpha=0.5, tag=6)
        out = None
        for k in range(2):  # 2 n > 8 K-segments: one GEMM per token kind, the second averages
            segs = sorted((((j * kinds + k) * B, 0, order[j] * D, D) for j in range(n)),
                          key=lambda sg: sg[2])
            out = L.gemm(a, wk[k], precision=prec, m=B, segs=segs, bias=bk[k], out_kind=L.OUT_F32,
                         tag=6, **({} if k == 0 else dict(resid=out, alpha=0.5)))
        return out

    def _lin(self, name, a, out_kind, **epi):
        prec = _PREC[self.precision]
        pk = self._packed(a.device)[name]
        return L.gemm(a, pk.w, precision=prec, bias=pk.b, out_kind=out_kind, tag=6, **epi)

    @torch.no_grad()
    def forward_gathered(self, slab, order=None):
        """Fusion head straight from a gathered operand slab [n, 2, B, D] (bf16; or
        [2, n, 2, B, D] split planes in fp32 mode) whose entry j holds sub-model order[j] --
        the layout an all-gather of the per-rank [2, B, D] blocks produces.  'deit' models."""
        prec = _PREC[self.precision]
        opk = L.OUT_BF16 if prec == L.DEVIT_BF16 else L.OUT_F32_SPLIT
        if collapse_head() and 'deit' in self.model and 'vit' not in self.model:
            return self._head_collapsed(slab, order)
        if self.teacher_size is not None:
            hc = self._fuse(slab, 0, 'cls_mlp', opk, order)
            hd = self._fuse(slab, 1, 'dist_mlp', opk, order)
            cls_logits = self._lin('cls_classifier', hc, L.OUT_F32)
            return self._lin('dist_classifier', hd, L.OUT_F32, resid=cls_logits, alpha=0.5)
        cls_logits = self._fuse(slab, 0, 'cls_classifier', L.OUT_F32, order)
        return self._fuse(slab, 1, 'dist_classifier', L.OUT_F32, order, resid=cls_logits,
                          alpha=0.5)

    @torch.no_grad()
    def forward(self, x, distill=False):
        prec = _PREC[self.precision]
        opk = L.OUT_BF16 if prec == L.DEVIT_BF16 else L.OUT_F32_SPLIT
        want_tokens = distill and self.training and self.teacher_size is not None
        if 'vit' in self.model:
            slab, kind = self._slab_of(x, x[0].device)
            if self.teacher_size is not None:
                h = self._fuse(slab, kind, 'cls_mlp', L.OUT_F32 if want_tokens else opk)
                ens_token = h
                logits = self._lin('cls_classifier', L.to_operand(h, prec) if want_tokens else h,
                                   L.OUT_F32)
            else:
                ens_token = None
                logits = self._fuse(slab, kind, 'cls_classifier', L.OUT_F32)
        elif 'deit' in self.model:
            cls_list, dist_list = x
            cslab, ckind = self._slab_of(cls_list, cls_list[0].device)
            dslab, dkind = self._slab_of(dist_list, dist_list[0].device)
            if collapse_head() and not want_tokens and cslab is dslab and (ckind, dkind) == (0, 1):
                return self._head_collapsed(cslab, None)
            if self.teacher_size is not None:
                hk = L.OUT_F32 if want_tokens else opk
                hc = self._fuse(cslab, ckind, 'cls_mlp', hk)
                hd = self._fuse(dslab, dkind, 'dist_mlp', hk)
                ens_token = (hc, hd)
                ac = L.to_operand(hc, prec) if want_tokens else hc
                ad = L.to_operand(hd, prec) if want_tokens else hd
                cls_logits = self._lin('cls_classifier', ac, L.OUT_F32)
                # (cls_logits + dist_logits) / 2 in the last GEMM's epilogue
                logits = self._lin('dist_classifier', ad, L.OUT_F32, resid=cls_logits, alpha=0.5)
            else:
                ens_token = None
                cls_logits = self._fuse(cslab, ckind, 'cls_classifier', L.OUT_F32)
                logits = self._fuse(dslab, dkind, 'dist_classifier', L.OUT_F32, resid=cls_logits,
                                    alpha=0.5)
        else:
            raise L.DevitError(f"EnsMLP: model name {self.model!r} contains neither 'vit' nor "
                               f"'deit' (models/ensemble_models.py:66,73)")
        if want_tokens:
            return ens_token, logits
        return logits
