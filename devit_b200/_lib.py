"""ctypes binding of ``libdevit_b200.so`` (the C ABI declared in ``include/devit_b200.h``).

This module is deliberately thin: it only turns torch tensors into raw device pointers and
forwards the current CUDA stream.  There is no fallback -- if the shared library is missing the
import of any compute entry point raises, and on a non-sm_100 device every call returns
``DEVIT_ERR_DEVICE`` which is re-raised here as ``DevitError``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libdevit_b200.so"

ABI_VERSION = 4
DEVIT_BF16, DEVIT_FP32 = 0, 1
OUT_BF16, OUT_F32, OUT_F32_SPLIT = 0, 1, 2
ACT_NONE, ACT_GELU_ERF, ACT_RELU = 0, 1, 2
LAYOUT_NCHW, LAYOUT_NHWC = 0, 1


class DevitError(RuntimeError):
    pass


class GemmSeg(C.Structure):
    _fields_ = [("a_row_off", C.c_int32), ("a_k_off", C.c_int32), ("b_k_off", C.c_int32),
                ("k_len", C.c_int32)]


class GemmArgs(C.Structure):
    _fields_ = [
        ("precision", C.c_int32), ("m", C.c_int32), ("n", C.c_int32),
        ("a", C.c_void_p), ("a_rows", C.c_int32), ("a_cols", C.c_int32), ("lda", C.c_int64),
        ("a_plane_stride", C.c_int64),
        ("b", C.c_void_p), ("b_rows", C.c_int32), ("b_cols", C.c_int32), ("ldb", C.c_int64),
        ("b_plane_stride", C.c_int64),
        ("num_segs", C.c_int32), ("segs", GemmSeg * 8),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out_kind", C.c_int32),
        ("out_plane_stride", C.c_int64),
        ("bias", C.c_void_p), ("resid", C.c_void_p), ("ldr", C.c_int64),
        ("resid_period", C.c_int32),
        ("rowbias", C.c_void_p), ("ld_rowbias", C.c_int64),
        ("act", C.c_int32), ("alpha", C.c_float),
        ("rowmap_period", C.c_int32), ("rowmap_stride", C.c_int32), ("rowmap_off", C.c_int32),
        ("block_n", C.c_int32), ("profile_tag", C.c_int32), ("cluster_m", C.c_int32),
        # LayerNorm folding (see include/devit_b200.h)
        ("ln_stats", C.c_void_p), ("ln_parts", C.c_int32), ("ln_dim", C.c_int32),
        ("ln_eps", C.c_float), ("ln_colsum", C.c_void_p),
        ("out_bf16", C.c_void_p), ("ld_out_bf16", C.c_int64), ("stats_out", C.c_void_p),
    ]


class LayerDesc(C.Structure):
    _fields_ = [
        ("heads", C.c_int32), ("hidden", C.c_int32), ("hidden_ld", C.c_int32),
        ("ln1_g", C.c_void_p), ("ln1_b", C.c_void_p),
        ("w_qkv", C.c_void_p), ("b_qkv", C.c_void_p),
        ("w_proj", C.c_void_p), ("b_proj", C.c_void_p),
        ("ln2_g", C.c_void_p), ("ln2_b", C.c_void_p),
        ("w_fc1", C.c_void_p), ("b_fc1", C.c_void_p),
        ("w_fc2", C.c_void_p), ("b_fc2", C.c_void_p),
        ("cs_qkv", C.c_void_p), ("cs_fc1", C.c_void_p),
    ]


class MlpArgs(C.Structure):
    _fields_ = [
        ("m", C.c_int32), ("dim", C.c_int32), ("hidden_ld", C.c_int32),
        ("xb", C.c_void_p), ("w1", C.c_void_p), ("c1", C.c_void_p), ("c2", C.c_void_p),
        ("ln_stats", C.c_void_p), ("ln_parts", C.c_int32), ("ln_eps", C.c_float),
        ("w2", C.c_void_p), ("b2", C.c_void_p), ("x", C.c_void_p), ("xb_out", C.c_void_p),
        ("stats_out", C.c_void_p),
        ("o", C.c_void_p), ("w_proj", C.c_void_p), ("b_proj", C.c_void_p), ("proj_k", C.c_int32),
        ("x_lo_in", C.c_void_p), ("x_lo_out", C.c_void_p),
    ]


class BlockWeights(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("num_heads", C.c_int32), ("hidden", C.c_int32),
        ("ln1_g", C.c_void_p), ("ln1_b", C.c_void_p),
        ("w_qkv", C.c_void_p), ("b_qkv", C.c_void_p),
        ("w_proj", C.c_void_p), ("b_proj", C.c_void_p),
        ("ln2_g", C.c_void_p), ("ln2_b", C.c_void_p),
        ("w_fc1", C.c_void_p), ("b_fc1", C.c_void_p),
        ("w_fc2", C.c_void_p), ("b_fc2", C.c_void_p),
        ("head_gate", C.POINTER(C.c_float)), ("neuron_gate", C.POINTER(C.c_float)),
    ]


class VitDesc(C.Structure):
    _fields_ = [
        ("precision", C.c_int32), ("dim", C.c_int32), ("depth", C.c_int32),
        ("img", C.c_int32), ("chans", C.c_int32), ("num_prefix", C.c_int32),
        ("ln_eps", C.c_float),
        ("w_patch", C.c_void_p), ("b_patch", C.c_void_p),
        ("prefix", C.c_void_p), ("pos", C.c_void_p),
        ("norm_g", C.c_void_p), ("norm_b", C.c_void_p),
        ("layers", C.POINTER(LayerDesc)),
        ("w_plane_stride_unused", C.c_int64),
        ("tok_table", C.c_void_p),
    ]


class VitExports(C.Structure):
    _fields_ = [("qkv", C.c_void_p * 32), ("feats_kind_rows", C.c_int32)]


class CctDesc(C.Structure):
    _fields_ = [
        ("precision", C.c_int32), ("dim", C.c_int32), ("depth", C.c_int32),
        ("img", C.c_int32), ("chans", C.c_int32), ("n_conv", C.c_int32),
        ("conv_chans", C.c_int32 * 3), ("w_conv", C.c_void_p * 3), ("conv_kpad", C.c_int32 * 3),
        ("pos", C.c_void_p), ("ln_eps", C.c_float),
        ("norm_g", C.c_void_p), ("norm_b", C.c_void_p),
        ("pool_w", C.c_void_p), ("pool_b", C.c_float),
        ("layers", C.POINTER(LayerDesc)),
    ]


# name -> (restype, argtypes); this table is also what tests/test_host_cpu.py checks against the header
_SIGS = {
    "devit_abi_version": (C.c_int, []),
    "devit_last_error": (C.c_char_p, []),
    "devit_device_check": (C.c_int, []),
    "devit_launch_count": (C.c_longlong, []),
    "devit_set_sm_budget": (C.c_int, [C.c_int]),
    "devit_tmap_cache_stats": (C.c_int, [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "devit_profile_enable": (C.c_int, [C.c_int]),
    "devit_profile_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "devit_gemm": (C.c_int, [C.POINTER(GemmArgs), C.c_void_p]),
    "devit_debug_set_trace": (C.c_int, [C.c_void_p]),
    "devit_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_int32, C.c_float, C.c_int32, C.c_int64, C.c_void_p]),
    "devit_mlp_fused": (C.c_int, [C.POINTER(MlpArgs), C.c_void_p]),
    "devit_rowstats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                 C.c_void_p]),
    "devit_attention": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                  C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p]),
    "devit_im2col_patch16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_int64, C.c_void_p]),
    "devit_im2col_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_int64, C.c_void_p]),
    "devit_im2col_tokens_u8": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_float),
                                         C.POINTER(C.c_float), C.c_void_p, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p]),
    "devit_eval_tail_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "devit_eval_tail": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "devit_token_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "devit_vit_forward_patches": (C.c_int, [C.POINTER(VitDesc), C.c_void_p, C.c_int64, C.c_int32,
                                            C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                            C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "devit_vit_forward_ex": (C.c_int, [C.POINTER(VitDesc), C.c_void_p, C.c_void_p, C.c_int64,
                                       C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                       C.c_int64, C.c_void_p, C.c_int32, C.POINTER(VitExports),
                                       C.c_void_p]),
    "devit_cct_workspace_bytes": (C.c_size_t, [C.POINTER(CctDesc), C.c_int32]),
    "devit_cct_forward": (C.c_int, [C.POINTER(CctDesc), C.c_void_p, C.c_int32, C.c_void_p,
                                    C.c_size_t, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "devit_im2col3x3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                  C.c_int32, C.c_int64, C.c_void_p]),
    "devit_maxpool3x3s2_cl": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_void_p]),
    "devit_seqpool": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int32,
                                C.c_int32, C.c_int32, C.c_void_p]),
    "devit_token_prefix": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_void_p]),
    "devit_gather_ln": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "devit_vit_workspace_bytes": (C.c_size_t, [C.POINTER(VitDesc), C.c_int32]),
    "devit_pack_layer_bytes": (C.c_size_t, [C.POINTER(BlockWeights), C.c_int32]),
    "devit_pack_layer": (C.c_int, [C.POINTER(BlockWeights), C.c_int32, C.c_int32, C.c_void_p,
                                   C.c_size_t, C.POINTER(LayerDesc), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32), C.c_void_p]),
    "devit_vit_forward": (C.c_int, [C.POINTER(VitDesc), C.c_void_p, C.c_int32, C.c_void_p,
                                    C.c_size_t, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_int32, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises DevitError with build instructions if absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("DEVIT_B200_LIB", LIB_PATH))
    if not path.exists():
        raise DevitError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; "
            f"g.build()'` or `make -C devit_b200/csrc`. devit_b200 has no CPU/PyTorch fallback.")
    lib = C.CDLL(str(path))
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.devit_abi_version() != ABI_VERSION:
        raise DevitError("libdevit_b200.so ABI version mismatch")
    _lib = lib
    return lib


def exported_symbols() -> list[str]:
    return list(_SIGS)


def check(rc: int) -> None:
    if rc != 0:
        msg = load().devit_last_error().decode(errors="replace")
        raise DevitError(f"devit_b200 error {rc}: {msg}")


def ptr(t: torch.Tensor | None) -> int | None:
    if t is None:
        return None
    if not t.is_cuda:
        raise DevitError("devit_b200 kernels need CUDA tensors (there is no CPU fallback)")
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    """cudaStream_t of torch's current stream on `device` (default: the current device).  The C
    side launches on the CURRENT device, so callers whose operands may live on another device wrap
    the call in ``torch.cuda.device(t.device)`` and pass that device here."""
    return torch.cuda.current_stream(device).cuda_stream


def on_operand_device(fn):
    """Runs a wrapper with the device of its first CUDA tensor argument made current: the C ABI
    launches on the current device and on torch's current stream OF THAT DEVICE, so a tensor on
    cuda:1 while cuda:0 is current would otherwise be handed to device 0's stream.  Mixed-device
    operands are rejected."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            if torch.is_tensor(a) and a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise DevitError(f"{fn.__name__}: operands on {dev} and {a.device}")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


# ------------------------------------------------------------------------------------------
# Convenience wrappers used by the modules and the tests.  Operand tensors:
#   DEVIT_BF16: torch.bfloat16 [rows, cols]
#   DEVIT_FP32: torch.float32  [2, rows, cols]   (hi plane, lo plane)
# ------------------------------------------------------------------------------------------
def split_tf32(t: torch.Tensor) -> torch.Tensor:
    """fp32 [..] -> [2, ..] (hi exactly representable in tf32, hi + lo == t)."""
    t = t.contiguous().float()
    bits = t.view(torch.int32)
    hi = ((bits + 0x1000) & ~0x1FFF).view(torch.float32)
    return torch.stack([hi, t - hi], 0).contiguous()


def to_operand(t: torch.Tensor, precision: int) -> torch.Tensor:
    if precision == DEVIT_BF16:
        return t.to(torch.bfloat16).contiguous()
    return split_tf32(t)


def operand_to_f32(t: torch.Tensor, precision: int) -> torch.Tensor:
    if precision == DEVIT_BF16:
        return t.float()
    return t[0] + t[1]


@on_operand_device
def gemm(a, b, *, precision=DEVIT_BF16, m=None, n=None, segs=None, out=None,
         out_kind=OUT_BF16, bias=None, resid=None, rowbias=None, act=ACT_NONE, alpha=1.0,
         rowmap=(0, 0, 0), block_n=0, out_rows=None, tag=0, cluster_m=0,
         ln_stats=None, ln_colsum=None, ln_dim=0, ln_eps=0.0, out_bf16=None, stats_out=None,
         resid_period=0):
    """out = epilogue(sum_s A_s B_s^T); see include/devit_b200.h (devit_gemm)."""
    lib = load()
    planes = 1 if precision == DEVIT_BF16 else 2
    a2 = a if planes == 1 else a[0]
    b2 = b if planes == 1 else b[0]
    a_rows, a_cols = a2.shape
    b_rows, b_cols = b2.shape
    if segs is None:
        segs = [(0, 0, 0, a_cols)]
    m = a_rows if m is None else m
    n = b_rows if n is None else n
    if out is None:
        rows = m if out_rows is None else out_rows
        if out_kind == OUT_BF16:
            out = torch.empty(rows, n, device=a.device, dtype=torch.bfloat16)
        elif out_kind == OUT_F32:
            out = torch.empty(rows, n, device=a.device, dtype=torch.float32)
        else:
            out = torch.empty(2, rows, n, device=a.device, dtype=torch.float32)
    o2 = out[0] if out_kind == OUT_F32_SPLIT else out
    g = GemmArgs()
    g.precision, g.m, g.n = precision, m, n
    g.a, g.a_rows, g.a_cols, g.lda = ptr(a), a_rows, a_cols, a2.stride(0)
    g.a_plane_stride = a.stride(0) if planes == 2 else 0
    g.b, g.b_rows, g.b_cols, g.ldb = ptr(b), b_rows, b_cols, b2.stride(0)
    g.b_plane_stride = b.stride(0) if planes == 2 else 0
    g.num_segs = len(segs)
    for i, s in enumerate(segs):
        g.segs[i] = GemmSeg(*s)
    g.out, g.ldo, g.out_kind = ptr(out), o2.stride(0), out_kind
    g.out_plane_stride = out.stride(0) if out_kind == OUT_F32_SPLIT else 0
    g.bias = ptr(bias)
    g.resid = ptr(resid)
    g.ldr = resid.stride(0) if resid is not None else 0
    g.resid_period = resid_period
    g.rowbias = ptr(rowbias)
    g.ld_rowbias = rowbias.stride(0) if rowbias is not None else 0
    g.act, g.alpha = act, alpha
    g.rowmap_period, g.rowmap_stride, g.rowmap_off = rowmap
    g.block_n = block_n
    g.profile_tag = tag
    g.cluster_m = cluster_m
    if ln_stats is not None:  # [parts, m, 2] partial (sum, sum^2) per row
        g.ln_stats, g.ln_parts = ptr(ln_stats), ln_stats.shape[0]
        g.ln_dim, g.ln_eps, g.ln_colsum = ln_dim, ln_eps, ptr(ln_colsum)
    if out_bf16 is not None:
        g.out_bf16, g.ld_out_bf16 = ptr(out_bf16), out_bf16.stride(0)
    g.stats_out = ptr(stats_out)
    check(lib.devit_gemm(C.byref(g), stream_ptr()))
    return out


TAGS = ["gemm_other", "gemm_patch", "gemm_qkv", "gemm_proj", "gemm_fc1", "gemm_fc2",
        "gemm_fusion", "gemm_head", "attention", "layernorm", "gather_ln", "im2col",
        "token_prefix", "gemm_mlp_fused", "eval_tail"]


def tmap_cache_stats() -> dict:
    """{'entries', 'hits', 'misses'} of the library's tensor-map (TMA descriptor) cache."""
    h, m = C.c_longlong(0), C.c_longlong(0)
    n = load().devit_tmap_cache_stats(C.byref(h), C.byref(m))
    return {"entries": n, "hits": h.value, "misses": m.value}


_profiling = False


def profiling() -> bool:
    """True while the per-launch CUDA-event profile is on (callers then keep to one stream)."""
    return _profiling


def profile_enable(on: bool) -> None:
    global _profiling
    check(load().devit_profile_enable(1 if on else 0))
    _profiling = bool(on)


def profile_collect() -> dict:
    """{tag: (milliseconds, launches)} accumulated since profile_enable(True)."""
    ms = (C.c_double * 16)()
    cnt = (C.c_longlong * 16)()
    check(load().devit_profile_collect(ms, cnt))
    return {TAGS[i]: (ms[i], cnt[i]) for i in range(len(TAGS)) if cnt[i]}


@on_operand_device
def im2col_tokens(images, num_prefix, precision=DEVIT_BF16):
    """Token-row patch matrix of an NCHW fp32 batch (see devit_im2col_tokens): bf16
    [B*tokens, C*256], or fp32 hi/lo planes [2, B*tokens, C*256] in the DEVIT_FP32 mode."""
    B, Cn, H, W = images.shape
    rows, k = B * (num_prefix + (H // 16) * (W // 16)), Cn * 256
    if precision == DEVIT_BF16:
        a = torch.empty(rows, k, device=images.device, dtype=torch.bfloat16)
        kind, plane = OUT_BF16, 0
    else:
        a = torch.empty(2, rows, k, device=images.device, dtype=torch.float32)
        kind, plane = OUT_F32_SPLIT, rows * k
    check(load().devit_im2col_tokens(ptr(images), ptr(a), B, Cn, H, num_prefix, kind, plane,
                                     stream_ptr()))
    return a


@on_operand_device
def im2col_tokens_u8(images, mean, std, num_prefix, precision=DEVIT_BF16, layout=LAYOUT_NCHW):
    """Token-row patch matrix of a uint8 batch ([B,C,H,W], or [B,H,W,3] with LAYOUT_NHWC) with
    ToTensor + Normalize(mean, std) applied on the device (see devit_im2col_tokens_u8)."""
    if images.dtype != torch.uint8 or images.dim() != 4:
        raise DevitError("im2col_tokens_u8 takes a 4-d uint8 tensor")
    images = images.contiguous()
    if layout == LAYOUT_NCHW:
        B, Cn, H, W = images.shape
    else:
        B, H, W, Cn = images.shape
    if H != W or len(mean) != Cn or len(std) != Cn:
        raise DevitError(f"im2col_tokens_u8: square images and {Cn} mean/std values expected")
    rows, k = B * (num_prefix + (H // 16) * (W // 16)), Cn * 256
    if precision == DEVIT_BF16:
        a = torch.empty(rows, k, device=images.device, dtype=torch.bfloat16)
        kind, plane = OUT_BF16, 0
    else:
        a = torch.empty(2, rows, k, device=images.device, dtype=torch.float32)
        kind, plane = OUT_F32_SPLIT, rows * k
    m = (C.c_float * Cn)(*[float(v) for v in mean])
    s = (C.c_float * Cn)(*[float(v) for v in std])
    check(load().devit_im2col_tokens_u8(ptr(images), layout, m, s, ptr(a), B, Cn, H, num_prefix,
                                        kind, plane, stream_ptr()))
    return a


@on_operand_device
def eval_tail(logits, target, acc=None, topk=5, want_batch=True):
    """Device-side CrossEntropy + top-1 / top-k counts of one batch (see devit_eval_tail).
    `acc`: float64 [5] running meters updated in place.  Returns float32 [3]
    {mean loss, #correct@1, #correct@k} of this batch (or None when want_batch is False)."""
    if logits.dtype != torch.float32 or logits.dim() != 2 or logits.stride(1) != 1:
        raise DevitError("eval_tail takes fp32 logits [batch, classes] with unit column stride")
    if target.dtype != torch.int64 or target.shape != (logits.shape[0],):
        raise DevitError("eval_tail takes int64 targets [batch]")
    if acc is not None and (acc.dtype != torch.float64 or acc.numel() != 5):
        raise DevitError("eval_tail: acc must be a float64 tensor of 5 elements")
    B, Cn = logits.shape
    lib = load()
    ws = torch.empty(lib.devit_eval_tail_workspace_bytes(B), device=logits.device,
                     dtype=torch.uint8)
    out = torch.empty(3, device=logits.device, dtype=torch.float32) if want_batch else None
    check(lib.devit_eval_tail(ptr(logits), logits.stride(0), ptr(target.contiguous()), B, Cn,
                              topk, ptr(ws), ws.numel(), ptr(acc), ptr(out), stream_ptr()))
    return out


@on_operand_device
def rowstats(x):
    """bf16 copy + per-row (sum, sum of squares) of an fp32 matrix -> (xb, stats[1, rows, 2])."""
    rows, dim = x.shape
    xb = torch.empty(rows, dim, device=x.device, dtype=torch.bfloat16)
    stats = torch.empty(1, rows, 2, device=x.device, dtype=torch.float32)
    check(load().devit_rowstats(ptr(x), ptr(xb), ptr(stats), rows, dim, stream_ptr()))
    return xb, stats


@on_operand_device
def mlp_fused(x, xb, ln_stats, w1, c1, c2, w2, b2, eps, xb_out=None, stats_out=None,
              o=None, w_proj=None, b_proj=None, x_lo_in=None, x_lo_out=None):
    """In-place x += gelu(LN(x) W1^T + b1) W2^T + b2 (LayerNorm folded); see devit_mlp_fused.
    With `o` / `w_proj` / `b_proj` the attention-output projection runs in front in the same
    kernel (x1 = x + o Wp^T + bp, then the MLP on x1; `xb` / `ln_stats` are not used)."""
    a = MlpArgs()
    a.m, a.dim, a.hidden_ld = x.shape[0], x.shape[1], w1.shape[0]
    a.xb, a.w1, a.c1, a.c2 = ptr(xb), ptr(w1), ptr(c1), ptr(c2)
    a.ln_stats, a.ln_parts, a.ln_eps = ptr(ln_stats), (ln_stats.shape[0] if ln_stats is not None
                                                      else 0), eps
    a.w2, a.b2, a.x = ptr(w2), ptr(b2), ptr(x)
    a.xb_out, a.stats_out = ptr(xb_out), ptr(stats_out)
    if o is not None:
        a.o, a.w_proj, a.b_proj, a.proj_k = ptr(o), ptr(w_proj), ptr(b_proj), o.shape[1]
    a.x_lo_in, a.x_lo_out = ptr(x_lo_in), ptr(x_lo_out)
    check(load().devit_mlp_fused(C.byref(a), stream_ptr()))
    return x


@on_operand_device
def layernorm(x, gamma, beta, eps, out_kind=OUT_BF16):
    lib = load()
    rows, dim = x.shape
    if out_kind == OUT_BF16:
        y = torch.empty(rows, dim, device=x.device, dtype=torch.bfloat16)
    elif out_kind == OUT_F32:
        y = torch.empty(rows, dim, device=x.device, dtype=torch.float32)
    else:
        y = torch.empty(2, rows, dim, device=x.device, dtype=torch.float32)
    check(lib.devit_layernorm(ptr(x), ptr(gamma), ptr(beta), ptr(y), rows, dim, eps, out_kind,
                              rows * dim, stream_ptr()))
    return y


@on_operand_device
def attention(qkv, batch, tokens, heads, scale, precision=DEVIT_BF16):
    lib = load()
    rows = batch * tokens
    if precision == DEVIT_BF16:
        out = torch.empty(rows, heads * 64, device=qkv.device, dtype=torch.bfloat16)
        check(lib.devit_attention(precision, ptr(qkv), 0, ptr(out), 0, batch, tokens, heads,
                                  scale, stream_ptr()))
    else:
        out = torch.empty(2, rows, heads * 64, device=qkv.device, dtype=torch.float32)
        check(lib.devit_attention(precision, ptr(qkv), qkv.stride(0), ptr(out), out.stride(0),
                                  batch, tokens, heads, scale, stream_ptr()))
    return out
