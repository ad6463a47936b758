"""ctypes binding of ``libdistb200.so`` (the C ABI declared in ``include/distb200.h``).

There is deliberately no fallback: if the shared library is missing or fails to load, importing
:func:`lib` raises.  Every wrapper here turns torch tensors into raw pointers / sizes and returns a
*prepared call* - a ``(c_function, argument tuple)`` pair - so that the engine can build the whole
forward once and replay it with negligible host work (and capture it into a CUDA graph).
"""

import ctypes as C
import os

import torch

from . import build as _build

F32, BF16 = 0, 1
ACT_NONE, ACT_QUICKGELU = 0, 1
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05, IMPL_TCGEN05_1CTA, IMPL_TCGEN05_2CTA = 0, 1, 2, 3, 4
MAX_TAPS = 9
STAT_SLOTS = 16

_TORCH2ENUM = {torch.float32: F32, torch.bfloat16: BF16}


class GemmDesc(C.Structure):
    """Mirror of ``distb200_gemm_desc`` (field order and types must match the header)."""
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p),
        ("dtype", C.c_int32), ("impl", C.c_int32),
        ("a_dim", C.c_int64 * 4), ("a_stride", C.c_int64 * 4),
        ("img_w", C.c_int32), ("num_taps", C.c_int32),
        ("tap_off", (C.c_int32 * 3) * MAX_TAPS),
        ("ldb", C.c_int64), ("b_tap_stride", C.c_int64),
        ("n", C.c_int32), ("k", C.c_int32),
        ("groups", C.c_int64), ("rows_per_group", C.c_int64),
        ("bias", C.c_void_p), ("res", C.c_void_p),
        ("ld_res", C.c_int64), ("res_gstride", C.c_int64), ("res_roff", C.c_int64), ("res_rep_stride", C.c_int64),
        ("out", C.c_void_p), ("out_dtype", C.c_int32), ("out_rep", C.c_int32),
        ("ld_out", C.c_int64), ("out_gstride", C.c_int64), ("out_roff", C.c_int64), ("out_rep_stride", C.c_int64),
        ("out2", C.c_void_p), ("out2_dtype", C.c_int32), ("act", C.c_int32),
        ("ld_out2", C.c_int64), ("block_n", C.c_int32), ("group_dim", C.c_int32),
        ("ln_stats", C.c_void_p), ("ln_wsum", C.c_void_p), ("stat_partials", C.c_void_p),
        ("act_from", C.c_int32), ("out2_gdiv", C.c_int32), ("out2_cstep", C.c_int32), ("act_to", C.c_int32),
        ("out2_gstride", C.c_int64), ("out2_roff", C.c_int64),
    ]


class WgradDesc(C.Structure):
    """Mirror of ``distb200_wgrad_desc``."""
    _fields_ = [
        ("x", C.c_void_p), ("dy", C.c_void_p),
        ("dtype", C.c_int32), ("impl", C.c_int32),
        ("a_dim", C.c_int64 * 4), ("a_stride", C.c_int64 * 4),
        ("img_w", C.c_int32), ("num_taps", C.c_int32),
        ("tap_off", (C.c_int32 * 3) * MAX_TAPS),
        ("n", C.c_int32), ("k", C.c_int32),
        ("groups", C.c_int64), ("rows_per_group", C.c_int64),
        ("group_dim", C.c_int32), ("reserved", C.c_int32),
        ("ld_dy", C.c_int64), ("dy_gstride", C.c_int64), ("dy_roff", C.c_int64),
        ("dw", C.c_void_p), ("ld_dw", C.c_int64), ("dw_tap_stride", C.c_int64),
    ]


class TemporalNetDesc(C.Structure):
    """Mirror of ``distb200_temporalnet_desc``."""
    _fields_ = [
        ("x", C.c_void_p), ("u", C.c_void_p), ("alpha", C.c_int32), ("dtype", C.c_int32),
        ("ln_g", C.c_void_p), ("ln_b", C.c_void_p), ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("out", C.c_void_p), ("out2", C.c_void_p), ("ld_out2", C.c_int64),
        ("out2_gdiv", C.c_int32), ("out2_cstep", C.c_int32), ("out2_gstride", C.c_int64), ("out2_roff", C.c_int64),
        ("clips", C.c_int32), ("frames", C.c_int32), ("grid", C.c_int32), ("channels", C.c_int32),
        ("eps", C.c_float), ("max_ctas", C.c_int32),
    ]


_LIB = None


def lib():
    """Load the CUDA library, building it first when the sources are newer (needs nvcc)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    alt = os.environ.get("DISTB200_LIB")          # A/B timing of two builds of the library on one GPU box (tools/); never a fallback
    if alt:
        path = alt
    elif _build.is_stale():
        try:
            _build.build()
        except Exception as exc:  # no nvcc on the box and no prebuilt library
            if not os.path.exists(path):
                raise RuntimeError(
                    "dist_b200: the CUDA extension {} is missing and could not be built ({}). There is no "
                    "CPU or PyTorch fallback for this path.".format(path, exc)) from exc
            # a prebuilt library exists but is OLDER than the sources and the rebuild failed: running it would silently test stale
            # kernels.  Refuse unless explicitly allowed (a GPU box without nvcc that received a library built elsewhere in time).
            if os.environ.get("DISTB200_ALLOW_STALE", "0") != "1":
                raise RuntimeError(
                    "dist_b200: {} is older than its sources and rebuilding failed ({}). Rebuild with `python -m dist_b200.build` "
                    "or set DISTB200_ALLOW_STALE=1 to load the stale library knowingly.".format(path, str(exc)[:300])) from exc
            import warnings
            warnings.warn("dist_b200: loading a STALE {} (rebuild failed: {})".format(path, str(exc)[:200]))
    try:
        L = C.CDLL(path)
    except OSError as exc:
        raise RuntimeError("dist_b200: cannot load {}: {} (no fallback exists)".format(path, exc)) from exc
    i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p
    L.distb200_version.restype = C.c_int
    L.distb200_arch.restype = C.c_int
    L.distb200_last_error.restype = C.c_char_p
    L.distb200_gemm.argtypes = [C.POINTER(GemmDesc), vp]
    L.distb200_temporalnet.argtypes = [C.POINTER(TemporalNetDesc), vp]
    L.distb200_layernorm.argtypes = [vp, i64, vp, i64, i64, i64, i32, f32, vp, vp, vp, i64, vp, vp, vp, i64, i32, vp]
    L.distb200_row_stats.argtypes = [vp, i32, i64, i64, i32, f32, vp, vp]
    L.distb200_row_stats_finalize.argtypes = [vp, i64, i32, f32, vp, vp]
    L.distb200_attention.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    L.distb200_cross_attention.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    L.distb200_attention_causal.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    L.distb200_embed_tokens.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp]
    L.distb200_gather_eot.argtypes = [vp, vp, i64, i32, i32, vp, vp]
    L.distb200_patchify.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i64, i32, vp]
    L.distb200_patchify_u8.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i64, i32, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3), vp]
    L.distb200_view_ensemble.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, i64, vp]
    L.distb200_topk_correct.argtypes = [vp, vp, i64, i32, vp, i32, vp, vp]
    L.distb200_rows_bcast.argtypes = [vp, i64, i64, i32, vp, i64, i32, vp, i64, vp]
    L.distb200_mean_rows.argtypes = [vp, i64, i32, i64, i32, vp, i32, vp]
    L.distb200_class_head.argtypes = [vp, vp, f32, i32, i32, i32, vp, vp, vp]
    L.distb200_class_head_fused.argtypes = [vp, vp, i32, f32, vp, f32, i32, i32, i32, vp, vp, vp, vp]
    # fine-tuning step
    L.distb200_gemm_wgrad.argtypes = [C.POINTER(WgradDesc), vp]
    L.distb200_quickgelu.argtypes = [vp, i32, vp, vp, i32, i64, vp]
    L.distb200_quickgelu_bwd.argtypes = [vp, i32, vp, i32, vp, vp, i32, i64, vp]
    L.distb200_cast.argtypes = [vp, vp, i32, i64, vp]
    L.distb200_group_sum.argtypes = [vp, i32, i64, i32, i64, vp, i32, vp]
    L.distb200_colsum.argtypes = [vp, i32, i64, i64, i64, i64, i64, i64, i32, vp, vp, vp]
    L.distb200_layernorm_bwd.argtypes = [vp, i64, vp, i64, i64, i64, i32, f32, vp, vp, i64, vp, vp, i64, i32, vp, i64, vp, i64, i32,
                                         vp, i64, i32, vp, vp, vp, vp, vp]
    L.distb200_cross_attention_bwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    L.distb200_softce_head.argtypes = [vp, vp, f32, vp, i32, i32, i32, vp, vp, vp, vp]
    L.distb200_adamw.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, f32, vp]
    L.distb200_pack_weight.argtypes = [vp, i64, i32, i32, vp, i64, vp, i64, i32, vp]
    for name in TRAIN_EXPORTS:
        getattr(L, name).restype = C.c_int
    for name in ("row_stats", "gemm", "layernorm", "attention", "cross_attention", "patchify", "patchify_u8", "rows_bcast", "mean_rows", "class_head", "view_ensemble",
                 "topk_correct", "attention_causal", "embed_tokens", "gather_eot", "row_stats_finalize", "temporalnet", "class_head_fused"):
        getattr(L, "distb200_" + name).restype = C.c_int
    assert L.distb200_version() == 100 and L.distb200_arch() == 100
    _LIB = L
    return L


TRAIN_EXPORTS = ("distb200_gemm_wgrad", "distb200_quickgelu", "distb200_quickgelu_bwd", "distb200_cast", "distb200_group_sum",
                 "distb200_colsum", "distb200_layernorm_bwd", "distb200_cross_attention_bwd", "distb200_softce_head",
                 "distb200_adamw", "distb200_pack_weight")

EXPORTS = TRAIN_EXPORTS + ("distb200_version", "distb200_arch", "distb200_last_error", "distb200_gemm", "distb200_row_stats", "distb200_layernorm",
           "distb200_attention", "distb200_cross_attention", "distb200_patchify", "distb200_patchify_u8", "distb200_view_ensemble", "distb200_topk_correct", "distb200_rows_bcast",
           "distb200_mean_rows", "distb200_class_head", "distb200_attention_causal", "distb200_embed_tokens", "distb200_gather_eot",
           "distb200_row_stats_finalize", "distb200_temporalnet", "distb200_class_head_fused")


class DistB200Error(RuntimeError):
    pass


def _check(code, what):
    if code != 0:
        raise DistB200Error("{} failed ({}): {}".format(what, code, lib().distb200_last_error().decode()))


def _ptr(t):
    return None if t is None else t.data_ptr()


def enum_of(t):
    return _TORCH2ENUM[t.dtype]


class Call:
    """A prepared C call: ``launch(stream_ptr)`` runs it on the given CUDA stream."""
    __slots__ = ("fn", "args", "name", "keep", "flops", "bytes")

    def __init__(self, fn, args, name, keep=(), flops=0, nbytes=0):
        self.fn, self.args, self.name, self.keep, self.flops, self.bytes = fn, args, name, keep, flops, nbytes

    def launch(self, stream):
        code = self.fn(*self.args, stream)
        if code != 0:
            _check(code, self.name)


def gemm(a, b, n, k, *, a_dim=None, a_stride=None, taps=((0, 0, 0),), b_tap_stride=0, ldb=None, img_w=0,
         groups=1, rows_per_group=None, group_dim=2, bias=None, res=None, ld_res=0, res_gstride=0, res_roff=0,
         res_rep_stride=0, out=None, ld_out=0, out_gstride=None, out_roff=0, out_rep=1, out_rep_stride=0,
         out2=None, ld_out2=0, act=ACT_NONE, impl=IMPL_AUTO, block_n=0, ln_stats=None, ln_wsum=None, stat_partials=None, act_from=0, act_to=0, out2_gdiv=0, out2_cstep=0, out2_gstride=0, out2_roff=0, name="gemm"):
    """Prepare one ``distb200_gemm`` (see the header for the exact definition).

    Defaults describe a plain ``out[M, n] = a[M, k] @ b[n, k]^T``: ``a`` is a 2-D row-major tensor,
    one group of ``M`` rows.  ``a_dim`` / ``a_stride`` override the logical A tensor (elements).
    """
    d = GemmDesc()
    d.a, d.b = a.data_ptr(), b.data_ptr()
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    d.dtype, d.impl = enum_of(a), impl
    if a_dim is None:
        assert a.dim() == 2 and a.stride(1) == 1
        a_dim = (k, a.shape[0], 1, 1)
        a_stride = (1, a.stride(0), a.stride(0) * a.shape[0], a.stride(0) * a.shape[0])
        if rows_per_group is None:
            rows_per_group = a.shape[0]
    for i in range(4):
        d.a_dim[i], d.a_stride[i] = int(a_dim[i]), int(a_stride[i])
    d.img_w, d.num_taps = int(img_w), len(taps)
    assert 1 <= len(taps) <= MAX_TAPS
    for j, tp in enumerate(taps):
        for i in range(3):
            d.tap_off[j][i] = int(tp[i])
    d.ldb = int(ldb if ldb is not None else b.stride(-2))
    d.b_tap_stride = int(b_tap_stride)
    d.n, d.k = int(n), int(k)
    d.groups, d.rows_per_group, d.group_dim = int(groups), int(rows_per_group), int(group_dim)
    d.bias, d.res = _ptr(bias), _ptr(res)
    if bias is not None:
        assert bias.dtype == torch.float32
    if res is not None:
        assert res.dtype == torch.float32
    d.ld_res, d.res_gstride, d.res_roff, d.res_rep_stride = int(ld_res), int(res_gstride), int(res_roff), int(res_rep_stride)
    d.out = _ptr(out)
    d.out_dtype = enum_of(out) if out is not None else F32
    d.out_rep = int(out_rep)
    d.ld_out = int(ld_out)
    d.out_gstride = int(out_gstride if out_gstride is not None else rows_per_group)
    d.out_roff, d.out_rep_stride = int(out_roff), int(out_rep_stride)
    d.out2 = _ptr(out2)
    d.out2_dtype = enum_of(out2) if out2 is not None else F32
    d.act, d.ld_out2, d.block_n = int(act), int(ld_out2), int(block_n)
    d.ln_stats, d.ln_wsum = _ptr(ln_stats), _ptr(ln_wsum)
    assert (ln_stats is None) == (ln_wsum is None)
    d.stat_partials = _ptr(stat_partials)
    d.act_from, d.act_to = int(act_from), int(act_to)
    d.out2_gdiv, d.out2_cstep, d.out2_gstride, d.out2_roff = int(out2_gdiv), int(out2_cstep), int(out2_gstride), int(out2_roff)
    if stat_partials is not None:
        assert stat_partials.dtype == torch.float32 and stat_partials.numel() >= int(groups) * int(rows_per_group) * STAT_SLOTS * 2
    rows = int(groups) * int(rows_per_group)
    flops = 2 * rows * int(n) * int(k) * len(taps)
    esz = a.element_size()
    nbytes = rows * int(k) * esz + int(n) * int(k) * len(taps) * esz
    if out is not None:
        nbytes += rows * int(n) * out.element_size() * int(out_rep)
    if out2 is not None:
        nbytes += rows * int(n) * out2.element_size() * int(out_rep)
    if res is not None:
        nbytes += rows * int(n) * 4 * int(out_rep)
    return Call(lib().distb200_gemm, (C.byref(d),), name, keep=(d, a, b, bias, res, out, out2, ln_stats, ln_wsum, stat_partials), flops=flops, nbytes=nbytes)


def temporalnet(x, ln_g, ln_b, w1, b1, w2, b2, *, clips, frames, grid, u=None, alpha=1, out=None, out2=None, ld_out2=0,
                out2_gdiv=0, out2_cstep=0, out2_gstride=0, out2_roff=0, eps=1e-5, max_ctas=0, name="temporalnet"):
    """Prepare one fused TemporalNet block (``distb200_temporalnet``; dist.py:48-65 + the i2t add of dist.py:231).

    ``x`` fp32 ``[clips, frames, grid*grid, C]`` channels-last, ``w1`` bf16 ``[3, C, C]``, ``w2`` bf16 ``[9, C, C]`` (per-tap K-major),
    ``u`` optional bf16 ``[clips, frames/alpha, grid*grid, C]``."""
    ch = int(w1.shape[-1])
    assert x.dtype == torch.float32 and w1.dtype == w2.dtype == torch.bfloat16
    assert tuple(w1.shape) == (3, ch, ch) and tuple(w2.shape) == (9, ch, ch) and w1.is_contiguous() and w2.is_contiguous()
    assert u is None or u.dtype == torch.bfloat16
    assert out is None or out.dtype == torch.float32
    assert out2 is None or out2.dtype == torch.bfloat16
    d = TemporalNetDesc()
    d.x, d.u, d.alpha, d.dtype = x.data_ptr(), _ptr(u), int(alpha), BF16
    d.ln_g, d.ln_b, d.w1, d.b1, d.w2, d.b2 = ln_g.data_ptr(), ln_b.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
    d.out, d.out2, d.ld_out2 = _ptr(out), _ptr(out2), int(ld_out2 if out2 is not None else 0)
    d.out2_gdiv, d.out2_cstep, d.out2_gstride, d.out2_roff = int(out2_gdiv), int(out2_cstep), int(out2_gstride), int(out2_roff)
    d.clips, d.frames, d.grid, d.channels = int(clips), int(frames), int(grid), ch
    d.eps, d.max_ctas = float(eps), int(max_ctas)
    rows = int(clips) * int(frames) * int(grid) * int(grid)
    flops = 2 * rows * ch * ch * 12
    nbytes = rows * ch * (4 + (4 if out is not None else 0) + (2 if out2 is not None else 0)) + (rows // int(alpha) * ch * 2 if u is not None else 0)
    return Call(lib().distb200_temporalnet, (C.byref(d),), name, keep=(d, x, u, ln_g, ln_b, w1, b1, w2, b2, out, out2), flops=flops, nbytes=nbytes)


def layernorm(x, g1, b1, y1, *, in2=None, in2_period=1, g2=None, b2=None, y2=None, rows=None, cols=None,
              ld_in1=None, ld_in2=None, ld_y1=None, ld_y2=None, eps=1e-5, name="layernorm"):
    assert x.dtype == torch.float32
    cols = int(cols if cols is not None else x.shape[-1])
    rows = int(rows if rows is not None else x.numel() // cols)
    ld_in1 = int(ld_in1 if ld_in1 is not None else cols)
    ld_y1 = int(ld_y1 if ld_y1 is not None else cols)
    if y2 is not None:
        assert y2.dtype == y1.dtype
    args = (x.data_ptr(), ld_in1, _ptr(in2), int(ld_in2 if ld_in2 is not None else cols), int(in2_period), rows, cols,
            float(eps), g1.data_ptr(), b1.data_ptr(), y1.data_ptr(), ld_y1, _ptr(g2), _ptr(b2), _ptr(y2),
            int(ld_y2 if ld_y2 is not None else cols), enum_of(y1))
    nbytes = rows * cols * (4 + (4 if in2 is not None else 0) + y1.element_size() * (2 if y2 is not None else 1))
    return Call(lib().distb200_layernorm, args, name, keep=(x, in2, g1, b1, y1, g2, b2, y2), nbytes=nbytes)


def row_stats(x, stats, rows=None, cols=None, ld=None, eps=1e-5, name="row_stats"):
    """(mean, rstd) per row of ``x`` -> ``stats`` [rows, 2] fp32 (the folded-LayerNorm input of :func:`gemm`)."""
    cols = int(cols if cols is not None else x.shape[-1])
    rows = int(rows if rows is not None else x.numel() // cols)
    assert stats.dtype == torch.float32
    if ld is None:
        ld = x.stride(0) if x.dim() == 2 else cols
    args = (x.data_ptr(), enum_of(x), int(ld), rows, cols, float(eps), stats.data_ptr())
    return Call(lib().distb200_row_stats, args, name, keep=(x, stats), nbytes=rows * cols * x.element_size())


def row_stats_finalize(partials, cols, stats, rows=None, eps=1e-5, name="row_stats_finalize"):
    """(mean, rstd) per row from the slots a GEMM epilogue emitted (``gemm(..., stat_partials=)``)."""
    rows = int(rows if rows is not None else stats.numel() // 2)
    assert partials.dtype == stats.dtype == torch.float32 and partials.numel() >= rows * STAT_SLOTS * 2
    args = (partials.data_ptr(), rows, int(cols), float(eps), stats.data_ptr())
    return Call(lib().distb200_row_stats_finalize, args, name, keep=(partials, stats), nbytes=rows * (STAT_SLOTS * 8 + 8))


def attention(qkv, out, frames, tokens, heads, impl=IMPL_AUTO, name="attention"):
    assert qkv.dtype == out.dtype
    args = (qkv.data_ptr(), out.data_ptr(), int(frames), int(tokens), int(heads), enum_of(qkv), int(impl))
    flops = 4 * frames * heads * tokens * tokens * 64
    nbytes = frames * tokens * heads * 64 * 4 * qkv.element_size()
    return Call(lib().distb200_attention, args, name, keep=(qkv, out), flops=flops, nbytes=nbytes)


def attention_causal(qkv, out, seqs, tokens, heads, name="attention_causal"):
    """Causal self attention of the CLIP text transformer (clip.py:404-410,122-124)."""
    assert qkv.dtype == out.dtype
    args = (qkv.data_ptr(), out.data_ptr(), int(seqs), int(tokens), int(heads), enum_of(qkv))
    return Call(lib().distb200_attention_causal, args, name, keep=(qkv, out), flops=2 * seqs * heads * tokens * (tokens + 1) * 64,
                nbytes=seqs * tokens * heads * 64 * 4 * qkv.element_size())


def embed_tokens(ids, table, pos, out, name="embed_tokens"):
    """out[s*ctx + i] = table[ids[s, i]] + pos[i]  (clip.py:420-421)"""
    assert ids.dtype == torch.int64 and ids.is_contiguous() and table.dtype == pos.dtype == out.dtype == torch.float32
    seqs, ctx = ids.shape
    width = table.shape[1]
    args = (ids.data_ptr(), table.data_ptr(), pos.data_ptr(), int(seqs), int(ctx), int(width), out.data_ptr())
    return Call(lib().distb200_embed_tokens, args, name, keep=(ids, table, pos, out), nbytes=seqs * ctx * width * 12)


def gather_eot(x, ids, out, name="gather_eot"):
    """out[s] = x[s*ctx + argmax_i ids[s, i]]  (clip.py:429)"""
    assert ids.dtype == torch.int64 and ids.is_contiguous() and x.dtype == out.dtype == torch.float32
    seqs, ctx = ids.shape
    width = out.shape[-1]
    args = (x.data_ptr(), ids.data_ptr(), int(seqs), int(ctx), int(width), out.data_ptr())
    return Call(lib().distb200_gather_eot, args, name, keep=(x, ids, out), nbytes=seqs * width * 8)


def cross_attention(q, kv, out, batch, keys, heads, name="cross_attention"):
    assert q.dtype == kv.dtype == out.dtype
    args = (q.data_ptr(), kv.data_ptr(), out.data_ptr(), int(batch), int(keys), int(heads), enum_of(q))
    return Call(lib().distb200_cross_attention, args, name, keep=(q, kv, out),
                flops=4 * batch * keys * heads * 64, nbytes=batch * keys * heads * 128 * q.element_size())


def patchify(video, out, clips, T, H, W, p, first, step, n_sel, ld_out, name="patchify"):
    assert video.dtype == torch.float32 and video.is_contiguous()
    args = (video.data_ptr(), out.data_ptr(), int(clips), int(T), int(H), int(W), int(p), int(first), int(step), int(n_sel),
            int(ld_out), enum_of(out))
    px = clips * n_sel * 3 * H * W
    return Call(lib().distb200_patchify, args, name, keep=(video, out), nbytes=px * (4 + out.element_size()))


def patchify_u8(frames, out, clips, T, H, W, p, first, step, n_sel, ld_out, mean, std, name="patchify_u8"):
    """frames uint8 [clips, T, H, W, 3]; mean / std: three floats each (DATA.MEAN / DATA.STD)"""
    assert frames.dtype == torch.uint8 and frames.is_contiguous()
    m3, s3 = (C.c_float * 3)(*[float(v) for v in mean]), (C.c_float * 3)(*[float(v) for v in std])
    args = (frames.data_ptr(), out.data_ptr(), int(clips), int(T), int(H), int(W), int(p), int(first), int(step), int(n_sel),
            int(ld_out), enum_of(out), C.byref(m3), C.byref(s3))
    px = clips * n_sel * 3 * H * W
    return Call(lib().distb200_patchify_u8, args, name, keep=(frames, out, m3, s3), nbytes=px * (1 + out.element_size()))


def view_ensemble(preds, labels, clip_ids, num_clips, method, video_preds, video_labels, clip_count, stream):
    """Immediate launch (batch sizes vary): accumulate the class scores of a batch of clips into their videos."""
    assert preds.dtype == torch.float32 and preds.is_contiguous() and labels.dtype == clip_ids.dtype == torch.int64
    n, c = preds.shape
    _check(lib().distb200_view_ensemble(preds.data_ptr(), labels.data_ptr(), clip_ids.data_ptr(), int(n), int(c), int(num_clips), int(method),
                                        video_preds.data_ptr(), video_labels.data_ptr(), clip_count.data_ptr(), int(video_preds.shape[0]),
                                        stream), "view_ensemble")


def topk_correct(video_preds, video_labels, ks_dev, correct, stream):
    v, c = video_preds.shape
    _check(lib().distb200_topk_correct(video_preds.data_ptr(), video_labels.data_ptr(), int(v), int(c), ks_dev.data_ptr(), int(ks_dev.numel()),
                                       correct.data_ptr(), stream), "topk_correct")


def rows_bcast(dst, row_stride, n_rows, cols, table, period, accumulate, dst2=None, row_stride2=0, name="rows_bcast"):
    assert dst.dtype == torch.float32 and table.dtype == torch.float32
    assert dst2 is None or dst2.dtype == torch.bfloat16
    args = (dst.data_ptr(), int(row_stride), int(n_rows), int(cols), table.data_ptr(), int(period), int(bool(accumulate)),
            _ptr(dst2), int(row_stride2))
    return Call(lib().distb200_rows_bcast, args, name, keep=(dst, table, dst2), nbytes=n_rows * cols * 8)


def mean_rows(src, row_stride, count, batch, cols, out, name="mean_rows"):
    assert src.dtype == torch.float32
    args = (src.data_ptr(), int(row_stride), int(count), int(batch), int(cols), out.data_ptr(), enum_of(out))
    return Call(lib().distb200_mean_rows, args, name, keep=(src, out), nbytes=batch * cols * (4 * count + out.element_size()))


def class_head(emb, text_n, scale, batch, embed_dim, classes, logits, probs, name="class_head"):
    args = (emb.data_ptr(), text_n.data_ptr(), float(scale), int(batch), int(embed_dim), int(classes), _ptr(logits), _ptr(probs))
    return Call(lib().distb200_class_head, args, name, keep=(emb, text_n, logits, probs))


def class_head_fused(emb, img, frames_per_clip, w, text_n, scale, batch, embed_dim, classes, logits, probs, img_n=None, name="class_head_fused"):
    """Class scores with the zero-shot / prediction-fusion average of clip.py:519-527 (video logits and mean per-frame CLIP logits)."""
    assert emb.dtype == img.dtype == text_n.dtype == torch.float32
    args = (emb.data_ptr(), img.data_ptr(), int(frames_per_clip), float(w), text_n.data_ptr(), float(scale), int(batch), int(embed_dim), int(classes),
            _ptr(logits), _ptr(probs), _ptr(img_n))
    return Call(lib().distb200_class_head_fused, args, name, keep=(emb, img, text_n, logits, probs, img_n))


# ---- fine-tuning step ------------------------------------------------------------------------------

def wgrad(x, dy, dw, n, k, *, a_dim=None, a_stride=None, taps=((0, 0, 0),), img_w=0, groups=1, rows_per_group=None, group_dim=2,
          ld_dy=None, dy_gstride=None, dy_roff=0, ld_dw=None, dw_tap_stride=None, impl=IMPL_AUTO, name="wgrad"):
    """Prepare ``distb200_gemm_wgrad``: dw[tap, n, k] += sum_rows dy[row, n] * A_tap(row, k) (A addressed as in :func:`gemm`)."""
    d = WgradDesc()
    assert x.dtype == dy.dtype and dw.dtype == torch.float32
    d.x, d.dy, d.dw = x.data_ptr(), dy.data_ptr(), dw.data_ptr()
    d.dtype, d.impl = enum_of(x), impl
    if a_dim is None:
        assert x.dim() == 2 and x.stride(1) == 1
        a_dim = (k, x.shape[0], 1, 1)
        a_stride = (1, x.stride(0), x.stride(0) * x.shape[0], x.stride(0) * x.shape[0])
        if rows_per_group is None:
            rows_per_group = x.shape[0]
    for i in range(4):
        d.a_dim[i], d.a_stride[i] = int(a_dim[i]), int(a_stride[i])
    d.img_w, d.num_taps = int(img_w), len(taps)
    for j, tp in enumerate(taps):
        for i in range(3):
            d.tap_off[j][i] = int(tp[i])
    d.n, d.k = int(n), int(k)
    d.groups, d.rows_per_group, d.group_dim = int(groups), int(rows_per_group), int(group_dim)
    d.ld_dy = int(ld_dy if ld_dy is not None else n)
    d.dy_gstride = int(dy_gstride if dy_gstride is not None else rows_per_group)
    d.dy_roff = int(dy_roff)
    d.ld_dw = int(ld_dw if ld_dw is not None else k)
    d.dw_tap_stride = int(dw_tap_stride if dw_tap_stride is not None else int(n) * d.ld_dw)
    rows = int(groups) * int(rows_per_group)
    flops = 2 * rows * int(n) * int(k) * len(taps)
    nbytes = rows * (int(n) + int(k)) * x.element_size()
    return Call(lib().distb200_gemm_wgrad, (C.byref(d),), name, keep=(d, x, dy, dw), flops=flops, nbytes=nbytes)


def quickgelu(z, y_f32=None, y_lp=None, n=None, name="quickgelu"):
    n = int(n if n is not None else z.numel())
    args = (z.data_ptr(), enum_of(z), _ptr(y_f32), _ptr(y_lp), enum_of(y_lp) if y_lp is not None else F32, n)
    nb = n * (z.element_size() + (4 if y_f32 is not None else 0) + (y_lp.element_size() if y_lp is not None else 0))
    return Call(lib().distb200_quickgelu, args, name, keep=(z, y_f32, y_lp), nbytes=nb)


def quickgelu_bwd(dy, z, dz_f32=None, dz_lp=None, n=None, name="quickgelu_bwd"):
    n = int(n if n is not None else z.numel())
    args = (dy.data_ptr(), enum_of(dy), z.data_ptr(), enum_of(z), _ptr(dz_f32), _ptr(dz_lp), enum_of(dz_lp) if dz_lp is not None else F32, n)
    nb = n * (dy.element_size() + z.element_size() + (4 if dz_f32 is not None else 0) + (dz_lp.element_size() if dz_lp is not None else 0))
    return Call(lib().distb200_quickgelu_bwd, args, name, keep=(dy, z, dz_f32, dz_lp), nbytes=nb)


def cast(src, dst, n=None, name="cast"):
    assert src.dtype == torch.float32
    n = int(n if n is not None else src.numel())
    return Call(lib().distb200_cast, (src.data_ptr(), dst.data_ptr(), enum_of(dst), n), name, keep=(src, dst), nbytes=n * (4 + dst.element_size()))


def group_sum(src, dst, groups, alpha, inner, name="group_sum"):
    args = (src.data_ptr(), enum_of(src), int(groups), int(alpha), int(inner), dst.data_ptr(), enum_of(dst))
    return Call(lib().distb200_group_sum, args, name, keep=(src, dst), nbytes=groups * inner * (alpha * src.element_size() + dst.element_size()))


def colsum(src, out, cols, *, ld=None, groups=1, rows_per_group=None, gstride=None, roff=0, period=1, out2=None, name="colsum"):
    assert out.dtype == torch.float32
    ld = int(ld if ld is not None else cols)
    if rows_per_group is None:
        rows_per_group = src.numel() // ld
    gstride = int(gstride if gstride is not None else rows_per_group)
    args = (src.data_ptr(), enum_of(src), ld, int(groups), int(rows_per_group), gstride, int(roff), int(period), int(cols), out.data_ptr(),
            _ptr(out2))
    return Call(lib().distb200_colsum, args, name, keep=(src, out, out2), nbytes=int(groups) * int(rows_per_group) * int(cols) * src.element_size())


def layernorm_bwd(x, g1, dy1, *, rows=None, cols=None, in2=None, in2_period=1, g2=None, dy2=None, add=None, dx=None, accumulate=False,
                  dx_lp=None, dg1=None, db1=None, dg2=None, db2=None, eps=1e-5, ld_x=None, ld_in2=None, ld_dy1=None, ld_dy2=None,
                  ld_add=None, ld_dx=None, ld_dx_lp=None, name="layernorm_bwd"):
    assert x.dtype == torch.float32
    cols = int(cols if cols is not None else x.shape[-1])
    rows = int(rows if rows is not None else x.numel() // cols)
    d = lambda v: int(v if v is not None else cols)
    if dy2 is not None:
        assert dy2.dtype == dy1.dtype
    args = (x.data_ptr(), d(ld_x), _ptr(in2), d(ld_in2), int(in2_period), rows, cols, float(eps),
            g1.data_ptr(), dy1.data_ptr(), d(ld_dy1), _ptr(g2), _ptr(dy2), d(ld_dy2), enum_of(dy1),
            _ptr(add), d(ld_add), _ptr(dx), d(ld_dx), int(bool(accumulate)),
            _ptr(dx_lp), d(ld_dx_lp), enum_of(dx_lp) if dx_lp is not None else F32,
            _ptr(dg1), _ptr(db1), _ptr(dg2), _ptr(db2))
    nb = rows * cols * (4 + dy1.element_size() * (2 if dy2 is not None else 1) + (4 if dx is not None else 0))
    return Call(lib().distb200_layernorm_bwd, args, name, keep=(x, in2, g1, dy1, g2, dy2, add, dx, dx_lp, dg1, db1, dg2, db2), nbytes=nb)


def cross_attention_bwd(q, kv, d_out, dq, dkv, batch, keys, heads, name="cross_attention_bwd"):
    assert q.dtype == kv.dtype == d_out.dtype == dq.dtype == dkv.dtype
    args = (q.data_ptr(), kv.data_ptr(), d_out.data_ptr(), dq.data_ptr(), dkv.data_ptr(), int(batch), int(keys), int(heads), enum_of(q))
    return Call(lib().distb200_cross_attention_bwd, args, name, keep=(q, kv, d_out, dq, dkv),
                nbytes=batch * keys * heads * 256 * q.element_size())


def softce_head(emb, text_n, scale, target, batch, embed_dim, classes, logits, loss, d_emb, name="softce_head"):
    args = (emb.data_ptr(), text_n.data_ptr(), float(scale), target.data_ptr(), int(batch), int(embed_dim), int(classes),
            _ptr(logits), loss.data_ptr(), d_emb.data_ptr())
    return Call(lib().distb200_softce_head, args, name, keep=(emb, text_n, target, logits, loss, d_emb))


def pack_weight(w, batch, n, k, out=None, out_t=None, ld_out=None, ld_out_t=None, name="pack_weight"):
    assert w.dtype == torch.float32
    ref = out if out is not None else out_t
    args = (w.data_ptr(), int(batch), int(n), int(k), _ptr(out), int(ld_out if ld_out is not None else k), _ptr(out_t),
            int(ld_out_t if ld_out_t is not None else n), enum_of(ref))
    return Call(lib().distb200_pack_weight, args, name, keep=(w, out, out_t), nbytes=int(batch) * n * k * (4 + 2 * ref.element_size()))


def adamw(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, stream):
    """Immediate launch (the scalar arguments change every step)."""
    _check(lib().distb200_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), int(n), float(lr), float(beta1), float(beta2),
                                float(eps), float(weight_decay), int(step), float(grad_scale), stream), "adamw")
