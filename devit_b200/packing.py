"""Gate-aware weight packing: nn.Parameters (fp32 masters) -> compacted GEMM operands.

The reference computes every head / neuron densely and multiplies the result by a 0/1 gate
(models/de_vit.py:41-43, :77-79).  Dropping a gated unit is numerically identical, so the pack
physically removes it:
  * kept heads  (gate != 0, ascending index) select rows of qkv.weight/bias and columns of
    proj.weight (scaled by the gate value, so non-binary gates stay exact);
  * kept neurons select rows of fc1.weight/bias and columns of fc2.weight (scaled likewise);
    the kept count is zero-padded to a multiple of 16 (padded neurons give gelu(0) = 0).
The kept-index lists are integer work and are exposed (``kept_heads`` / ``kept_neurons``) so the
tests can compare them bit-exactly with the oracle's.

Operand formats follow include/devit_b200.h: bf16 arrays, or fp32 hi/lo split planes for the
3xTF32 parity mode.  The per-layer compaction + LayerNorm folding runs on the device through the
C ABI (devit_pack_layer, csrc/pack.cu); the same packing written with torch ops is kept as its
cross-check (DEVIT_PACK_TORCH=1, tests/test_pack_gpu.py).  One-off work when a gate or a
parameter changes; not on the per-batch path.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib as L


def kept_indices(gate: torch.Tensor) -> torch.Tensor:
    """Ascending indices of the units whose gate is non-zero (int64, CPU)."""
    return torch.nonzero(gate.detach().float().cpu() != 0, as_tuple=False).flatten()


class PackedVit:
    """Device-resident packed weights + the ctypes descriptor passed to devit_vit_forward."""

    def __init__(self, model, precision: int, device: torch.device, fold_ln: bool = True):
        self.precision = precision
        self.device = device
        self.fold_ln = fold_ln
        self._keep = []  # tensors referenced by raw pointers in the descriptors
        self.kept_heads = []
        self.kept_neurons = []
        dim = model.embed_dim
        depth = len(model.blocks)
        self.layers = (L.LayerDesc * depth)()

        def op(t):  # GEMM operand in the mode's format
            t = L.to_operand(t.detach().to(device=device, dtype=torch.float32), precision)
            self._keep.append(t)
            return t.data_ptr()

        def f32(t):
            t = t.detach().to(device=device, dtype=torch.float32).contiguous()
            self._keep.append(t)
            return t.data_ptr()

        fold = precision == L.DEVIT_BF16 and fold_ln and dim % 128 == 0 and dim <= 768
        use_torch = os.environ.get('DEVIT_PACK_TORCH', '0') == '1' or device.type != 'cuda'
        for i, blk in enumerate(model.blocks):
            if blk.attn.num_heads * 64 != dim:
                raise L.DevitError(f"head_dim {dim // blk.attn.num_heads} unsupported (the "
                                   f"attention kernel is built for head_dim 64)")
            if use_torch:
                self._pack_layer_torch(i, blk, dim, precision, fold, op, f32)
            else:
                self._pack_layer_c(i, blk, dim, precision, fold, f32)

        pe = model.patch_embed
        desc = L.VitDesc()
        desc.precision, desc.dim, desc.depth = precision, dim, depth
        desc.img, desc.chans = int(pe.img_size[0]), int(pe.proj.weight.shape[1])
        desc.num_prefix = model.num_tokens
        desc.ln_eps = float(model.norm.eps)
        desc.w_patch = op(pe.proj.weight.detach().float().reshape(dim, -1))
        desc.b_patch = f32(pe.proj.bias)
        prefix = [model.cls_token.detach().reshape(1, dim)]
        if model.dist_token is not None:
            prefix.append(model.dist_token.detach().reshape(1, dim))
        desc.prefix = f32(torch.cat(prefix, 0))
        desc.pos = f32(model.pos_embed.detach().reshape(-1, dim))
        # wrapped table pos (+ cls/dist - patch bias) for the periodic-residual patch GEMM
        # (devit_vit_desc.tok_table); DEVIT_TOK_TABLE=0 falls back to devit_token_init
        desc.tok_table = None
        if os.environ.get('DEVIT_TOK_TABLE', '1') != '0':
            desc.tok_table = f32(token_table(model.pos_embed.detach().float().cpu().reshape(-1, dim),
                                             torch.cat(prefix, 0).float().cpu(),
                                             pe.proj.bias.detach().float().cpu()))
        desc.norm_g, desc.norm_b = f32(model.norm.weight), f32(model.norm.bias)
        desc.layers = C.cast(self.layers, C.POINTER(L.LayerDesc))
        desc.w_plane_stride_unused = 0
        self.desc = desc
        self.tokens = (desc.img // 16) ** 2 + desc.num_prefix
        self.dim = dim

    def _pack_layer_c(self, i, blk, dim, precision, fold, f32):
        """Gate compaction + LayerNorm folding on the device through the C ABI
        (devit_pack_layer, csrc/pack.cu)."""
        attn, mlp = blk.attn, blk.mlp
        w = L.BlockWeights()
        w.dim, w.num_heads, w.hidden = dim, attn.num_heads, mlp.hidden_features
        tmp = []  # fp32 device copies only devit_pack_layer reads (it synchronises before returning)

        def t32(t):
            t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
            tmp.append(t)
            return t.data_ptr()

        # the descriptor keeps pointing at these (LayerNorm parameters, proj / fc2 bias): kept alive
        w.ln1_g, w.ln1_b = f32(blk.norm1.weight), f32(blk.norm1.bias)
        w.ln2_g, w.ln2_b = f32(blk.norm2.weight), f32(blk.norm2.bias)
        w.b_proj, w.b_fc2 = f32(attn.proj.bias), f32(mlp.fc2.bias)
        # read once by the pack kernels
        w.w_qkv = t32(attn.qkv.weight)
        w.b_qkv = t32(attn.qkv.bias) if attn.qkv.bias is not None else None
        w.w_proj = t32(attn.proj.weight)
        w.w_fc1, w.b_fc1 = t32(mlp.fc1.weight), t32(mlp.fc1.bias)
        w.w_fc2 = t32(mlp.fc2.weight)
        hg = attn.gate.detach().float().cpu().contiguous()
        ng = mlp.gate.detach().float().cpu().contiguous()
        if hg.numel() != attn.num_heads or ng.numel() != mlp.hidden_features:
            raise L.DevitError("gate length does not match num_heads / hidden_features")
        w.head_gate = C.cast(hg.data_ptr(), C.POINTER(C.c_float))
        w.neuron_gate = C.cast(ng.data_ptr(), C.POINTER(C.c_float))
        lib = L.load()
        nbytes = lib.devit_pack_layer_bytes(C.byref(w), precision)
        if nbytes == 0:
            L.check(1)
        buf = torch.empty(nbytes, device=self.device, dtype=torch.uint8)
        self._keep.append(buf)
        kh = (C.c_int32 * attn.num_heads)()
        kn = (C.c_int32 * mlp.hidden_features)()
        nh, nn_ = C.c_int32(0), C.c_int32(0)
        with torch.cuda.device(self.device):
            L.check(lib.devit_pack_layer(C.byref(w), precision, 1 if fold else 0, buf.data_ptr(),
                                         nbytes, C.byref(self.layers[i]), kh, C.byref(nh), kn,
                                         C.byref(nn_), L.stream_ptr(self.device)))
        self.kept_heads.append(torch.tensor(list(kh[:nh.value]), dtype=torch.int64))
        self.kept_neurons.append(torch.tensor(list(kn[:nn_.value]), dtype=torch.int64))

    def _pack_layer_torch(self, i, blk, dim, precision, fold, op, f32):
        """The same packing with torch ops on the host (the cross-check of devit_pack_layer in
        tests/test_pack_gpu.py; DEVIT_PACK_TORCH=1 selects it)."""
        attn, mlp = blk.attn, blk.mlp
        nh = attn.num_heads
        hg = attn.gate.detach().float().cpu()
        hk = kept_indices(hg)
        hscale = hg[hk]
        if hk.numel() == 0:  # every head gated off: keep one, with zeroed proj columns
            hk, hscale = torch.tensor([0]), torch.tensor([0.0])
        self.kept_heads.append(hk.clone())
        hd = dim // nh
        cols = (hk[:, None] * hd + torch.arange(hd)[None, :]).flatten()  # kept q/k/v columns
        rows = torch.cat([cols + w * dim for w in range(3)])
        wq = attn.qkv.weight.detach().float().cpu()
        bq = attn.qkv.bias.detach().float().cpu() if attn.qkv.bias is not None \
            else torch.zeros(3 * dim)
        wp = attn.proj.weight.detach().float().cpu()[:, cols] * \
            hscale.repeat_interleave(hd)[None, :]
        ng = mlp.gate.detach().float().cpu()
        nk = kept_indices(ng)
        nscale = ng[nk]
        self.kept_neurons.append(nk.clone())
        f = int(nk.numel())
        f_ld = max(16, (f + 15) // 16 * 16)
        w1 = torch.zeros(f_ld, dim)
        b1 = torch.zeros(f_ld)
        w2 = torch.zeros(dim, f_ld)
        if f:
            w1[:f] = mlp.fc1.weight.detach().float().cpu()[nk]
            b1[:f] = mlp.fc1.bias.detach().float().cpu()[nk]
            w2[:, :f] = mlp.fc2.weight.detach().float().cpu()[:, nk] * nscale[None, :]
        d = self.layers[i]
        d.heads, d.hidden, d.hidden_ld = int(hk.numel()), max(f, 1), f_ld
        d.ln1_g, d.ln1_b = f32(blk.norm1.weight), f32(blk.norm1.bias)
        d.ln2_g, d.ln2_b = f32(blk.norm2.weight), f32(blk.norm2.bias)
        wqk, bqk = wq[rows], bq[rows]
        if fold:
            # LayerNorm folding (include/devit_b200.h, devit_gemm_args.ln_stats):
            # LN(x) W^T + b = rstd (x (gamma.W)^T - mean c1) + c2 with c1 = rowsum of the
            # folded weights AS THE TENSOR CORE SEES THEM (bf16-rounded), c2 = b + W beta.
            wqk, d.cs_qkv, bqk = self._fold(wqk, bqk, blk.norm1, f32)
            w1, d.cs_fc1, b1 = self._fold(w1, b1, blk.norm2, f32)
        d.w_qkv, d.b_qkv = op(wqk), f32(bqk)
        d.w_proj, d.b_proj = op(wp), f32(attn.proj.bias)
        d.w_fc1, d.b_fc1 = op(w1), f32(b1)
        d.w_fc2, d.b_fc2 = op(w2), f32(mlp.fc2.bias)

    def layer_arrays(self, i):
        """The packed arrays of layer i as tensors (views of the device buffers the descriptor
        points into): {'w_qkv', 'b_qkv', 'cs_qkv', 'w_proj', 'w_fc1', 'b_fc1', 'cs_fc1', 'w_fc2'}.
        Used by the tests that compare devit_pack_layer with the torch packing."""
        d = self.layers[i]
        hd, f_ld, dim = d.heads * 64, d.hidden_ld, self.dim
        bf = self.precision == L.DEVIT_BF16

        def view(ptr, rows, cols, weight):
            if not ptr:
                return None
            dt = torch.bfloat16 if (weight and bf) else torch.float32
            planes = 2 if (weight and not bf) else 1
            n = planes * rows * cols
            nbytes = n * (2 if dt == torch.bfloat16 else 4)
            for t in self._keep:
                off = ptr - t.data_ptr()
                if 0 <= off and off + nbytes <= t.numel() * t.element_size():
                    flat = t.view(-1).view(torch.uint8)[off:off + nbytes].view(dt)
                    return flat.view(planes, rows, cols) if planes == 2 else flat.view(rows, cols)
            raise L.DevitError("descriptor pointer outside every kept buffer")

        return {'w_qkv': view(d.w_qkv, 3 * hd, dim, True), 'b_qkv': view(d.b_qkv, 1, 3 * hd, False),
                'cs_qkv': view(d.cs_qkv, 1, 3 * hd, False), 'w_proj': view(d.w_proj, dim, hd, True),
                'w_fc1': view(d.w_fc1, f_ld, dim, True), 'b_fc1': view(d.b_fc1, 1, f_ld, False),
                'cs_fc1': view(d.cs_fc1, 1, f_ld, False), 'w_fc2': view(d.w_fc2, dim, f_ld, True)}

    @staticmethod
    def _fold(w, b, norm, f32):
        gamma = norm.weight.detach().float().cpu()
        beta = norm.bias.detach().float().cpu()
        wf = w * gamma[None, :]
        c1 = wf.to(torch.bfloat16).float().sum(1)
        c2 = b + w @ beta
        return wf, f32(c1), c2

    def workspace_bytes(self, batch: int) -> int:
        n = L.load().devit_vit_workspace_bytes(C.byref(self.desc), batch)
        if n == 0:
            L.check(1)
        return n


def token_table(pos, prefix, patch_bias, wrap=31):
    """[tokens + wrap, D]: row j = pos[j] + (prefix[j] - patch_bias if j < num_prefix else 0), the
    value the residual stream holds before the patch GEMM adds `A W^T + bias`
    (models/de_vit.py:259-264); the first `wrap` rows are repeated at the end so that a 32-row
    box starting at any row j < tokens never wraps."""
    t = pos.clone()
    n = prefix.shape[0]
    t[:n] += prefix - patch_bias[None, :]
    return torch.cat([t, t[:wrap]], 0).contiguous()


class PackedLinear:
    """One nn.Linear as a GEMM B operand (+ fp32 bias)."""

    def __init__(self, linear, precision: int, device: torch.device):
        self.w = L.to_operand(linear.weight.detach().to(device=device, dtype=torch.float32),
                              precision)
        self.b = None if linear.bias is None else \
            linear.bias.detach().to(device=device, dtype=torch.float32).contiguous()
        self.out_features, self.in_features = linear.weight.shape


_WORKSPACES = {}


def workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Per-(device, stream) scratch buffer, grown on demand and reused across calls."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, device=device, dtype=torch.uint8)
        _WORKSPACES[key] = buf
    return buf


def module_version(module) -> tuple:
    """Cheap fingerprint of everything a pack depends on: parameter storage + in-place version
    counters + the gate epochs.  A changed fingerprint invalidates the pack."""
    v = []
    for p in module.parameters():
        v.append(p.data_ptr())
        v.append(p._version)
    for m in module.modules():
        e = getattr(m, '_gate_epoch', None)
        if e is not None:
            v.append(e)
            g = m._gate
            v.append(g._version)
    return tuple(v)
