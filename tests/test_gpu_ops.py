"""Kernel-level parity on the GPU: every C ABI entry point against an fp64 torch restatement of its definition."""
import math

import pytest
import torch

from conftest import rel_l2
from helpers import gemm_reference

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from dist_b200 import ops
    return ops


def _run(call):
    call.launch(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()


def _qgelu(x):
    return x * torch.sigmoid(1.702 * x)


# scenario: a_shape = logical A as a torch tensor [c3, c2, c1, k_pitch]
SCENARIOS = {
    "plain_ragged": dict(a_shape=(1, 1, 300, 104), k=100, n=80, taps=[(0, 0, 0)], groups=1, rpg=300),
    "plain_vit": dict(a_shape=(1, 1, 1000, 768), k=768, n=256, taps=[(0, 0, 0)], groups=1, rpg=1000),
    "plain_wide": dict(a_shape=(1, 1, 520, 256), k=256, n=2304, taps=[(0, 0, 0)], groups=1, rpg=520),
    "n384": dict(a_shape=(1, 1, 394, 768), k=768, n=384, taps=[(0, 0, 0)], groups=1, rpg=394),
    "k96_n96": dict(a_shape=(1, 1, 777, 96), k=96, n=96, taps=[(0, 0, 0)], groups=1, rpg=777),
    "long_k": dict(a_shape=(1, 1, 256, 3072), k=3072, n=768, taps=[(0, 0, 0)], groups=1, rpg=256),
    "small_m": dict(a_shape=(1, 1, 2, 384), k=384, n=1536, taps=[(0, 0, 0)], groups=1, rpg=2),
    # (kt,1,1) conv over a clip: 3 taps, row shift +-P, zero padded inside each group
    "conv_t": dict(a_shape=(1, 3, 4 * 49, 96), k=96, n=96, taps=[(-49, 0, 0), (0, 0, 0), (49, 0, 0)], groups=3, rpg=4 * 49),
    # stem-like: 5 taps, padded K pitch
    "stem": dict(a_shape=(1, 2, 6 * 16, 592), k=588, n=96, taps=[((k - 2) * 16, 0, 0) for k in range(5)], groups=2, rpg=6 * 16),
    # (1,3,3) conv on a 14x14 grid: image mode, 9 taps
    "conv_s14": dict(a_shape=(5, 14, 14, 96), k=96, n=96, img_w=14, taps=[(j - 1, i - 1, 0) for i in range(3) for j in range(3)], groups=5, rpg=196),
    "conv_s16": dict(a_shape=(3, 16, 16, 96), k=96, n=96, img_w=16, taps=[(j - 1, i - 1, 0) for i in range(3) for j in range(3)], groups=3, rpg=256),
    # temporal -> integration: taps along dim 2, group on dim 3, rows written behind a class row
    "t2i": dict(a_shape=(6, 2, 196, 96), k=96, n=384, taps=[(0, 0, 0), (0, 1, 0)], groups=6, rpg=196, group_dim=3,
                out_gstride=197, out_roff=1, out_rows=6 * 197),
    # integration -> temporal: skip the class row on the input, replicate rows on the output
    "i2t": dict(a_shape=(1, 6, 197, 384), k=384, n=96, taps=[(1, 0, 0)], groups=6, rpg=196, out_gstride=2 * 196, out_rep=2,
                out_rep_stride=196, out_rows=12 * 196),
}


def _build(sc, dtype, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = torch.randn(*sc["a_shape"], generator=g)
    k, n, taps = sc["k"], sc["n"], sc["taps"]
    kp = sc["a_shape"][-1]
    a[..., k:] = 7.0        # pad columns must never be read
    b = torch.randn(len(taps), n, kp, generator=g) / math.sqrt(k * len(taps))
    b[..., k:] = 7.0
    bias = torch.randn(n, generator=g)
    out_rows = sc.get("out_rows", sc["groups"] * sc["rpg"])
    res = torch.randn(out_rows, n, generator=g)
    return a.to(DEV, dtype), b.to(DEV, dtype), bias.to(DEV), res.to(DEV)


def _expected(sc, a, b, bias, res, use_bias, use_res, act):
    shp = sc["a_shape"]
    a_dim = (sc["k"], shp[2], shp[1], shp[0])
    acc = gemm_reference(a.double(), a_dim, sc["taps"], b.double(), sc["groups"], sc["rpg"], sc.get("img_w", 0), sc.get("group_dim", 2))
    out_rows = sc.get("out_rows", sc["groups"] * sc["rpg"])
    gs, ro = sc.get("out_gstride", sc["rpg"]), sc.get("out_roff", 0)
    rep, rs = sc.get("out_rep", 1), sc.get("out_rep_stride", 0)
    exp = torch.full((out_rows, sc["n"]), float("nan"), dtype=torch.float64, device=a.device)
    gi = torch.arange(sc["groups"], device=a.device).repeat_interleave(sc["rpg"])
    r = torch.arange(sc["rpg"], device=a.device).repeat(sc["groups"])
    for q in range(rep):
        rows = gi * gs + ro + r + q * rs
        v = acc.clone()
        if use_bias:
            v = v + bias.double()
        if use_res:
            v = v + res.double()[rows]
        if act:
            v = _qgelu(v)
        exp[rows] = v
    return exp


def _call(ops, sc, a, b, bias, res, out, out2, use_bias, use_res, act, impl):
    shp = sc["a_shape"]
    kp = shp[-1]
    a_dim = (sc["k"], shp[2], shp[1], shp[0])
    a_stride = (1, kp, kp * shp[2], kp * shp[2] * shp[1])
    n = sc["n"]
    return ops.gemm(a, b, n, sc["k"], a_dim=a_dim, a_stride=a_stride, taps=sc["taps"], b_tap_stride=n * kp, ldb=kp,
                    img_w=sc.get("img_w", 0), groups=sc["groups"], rows_per_group=sc["rpg"], group_dim=sc.get("group_dim", 2),
                    bias=bias if use_bias else None, res=res if use_res else None, ld_res=n,
                    res_gstride=sc.get("out_gstride", sc["rpg"]), res_roff=sc.get("out_roff", 0), res_rep_stride=sc.get("out_rep_stride", 0),
                    out=out, ld_out=n, out_gstride=sc.get("out_gstride", sc["rpg"]), out_roff=sc.get("out_roff", 0),
                    out_rep=sc.get("out_rep", 1), out_rep_stride=sc.get("out_rep_stride", 0), out2=out2, ld_out2=n,
                    act=ops.ACT_QUICKGELU if act else ops.ACT_NONE, impl=impl)


@pytest.mark.parametrize("name", list(SCENARIOS))
@pytest.mark.parametrize("mode", ["simt_fp32", "simt_bf16", "tcgen05", "tcgen05_pair"])
def test_gemm(name, mode):
    ops = _ops()
    sc = SCENARIOS[name]
    dtype = torch.float32 if mode == "simt_fp32" else torch.bfloat16
    impl = {"tcgen05": ops.IMPL_TCGEN05_1CTA, "tcgen05_pair": ops.IMPL_TCGEN05_2CTA}.get(mode, ops.IMPL_SIMT)
    if mode.startswith("tcgen05") and sc["n"] % (32 if mode == "tcgen05_pair" else 16):
        pytest.skip("tcgen05 path needs n % 16 == 0 (n % 32 for CTA pairs)")
    a, b, bias, res = _build(sc, dtype)
    out_rows = sc.get("out_rows", sc["groups"] * sc["rpg"])
    for use_bias, use_res, act in [(False, False, False), (True, True, True), (True, False, True), (True, True, False)]:
        out = torch.full((out_rows, sc["n"]), float("nan"), device=DEV, dtype=torch.float32)
        out2 = torch.full((out_rows, sc["n"]), float("nan"), device=DEV, dtype=dtype)
        _run(_call(ops, sc, a, b, bias, res, out, out2, use_bias, use_res, act, impl))
        exp = _expected(sc, a, b, bias, res, use_bias, use_res, act)
        written = ~torch.isnan(exp)
        assert torch.equal(torch.isnan(out), ~written), "rows outside the mapping were touched (or rows were skipped)"
        tol = 2e-6 if mode == "simt_fp32" else 1e-5     # same (rounded) operands, fp32 accumulation
        assert rel_l2(out[written], exp[written]) < tol, (name, mode, use_bias, use_res, act)
        assert rel_l2(out2.float()[written], exp[written]) < (tol if dtype == torch.float32 else 3e-3)


def test_gemm_in_place_residual_bf16_out():
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    a = torch.randn(640, 768, generator=g).to(DEV, torch.bfloat16)
    w = (torch.randn(768, 768, generator=g) / 28).to(DEV, torch.bfloat16)
    bias = torch.randn(768, generator=g).to(DEV)
    h = torch.randn(640, 768, generator=g).to(DEV)
    exp = h.double() + a.double() @ w.double().t() + bias.double()
    tap = torch.empty(640, 768, device=DEV, dtype=torch.bfloat16)
    _run(ops.gemm(a, w, 768, 768, bias=bias, res=h, ld_res=768, out=h, ld_out=768, out2=tap, ld_out2=768))
    assert rel_l2(h, exp) < 1e-5
    assert rel_l2(tap.float(), exp) < 3e-3


def test_gemm_rejects_bad_arguments():
    ops = _ops()
    a = torch.zeros(8, 24, device=DEV, dtype=torch.bfloat16)
    w = torch.zeros(24, 24, device=DEV, dtype=torch.bfloat16)
    out = torch.zeros(8, 24, device=DEV)
    with pytest.raises(ops.DistB200Error):          # n % 16 != 0 on the tensor-core path
        _run(ops.gemm(a, w, 24, 24, out=out, ld_out=24))
    with pytest.raises(ops.DistB200Error):          # no output
        _run(ops.gemm(a, w, 24, 24))


@pytest.mark.parametrize("cols,rows", [(768, 1000), (1024, 77), (384, 515), (96, 3000), (128, 5), (32, 64)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm(cols, rows, dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(cols)
    x = (torch.randn(rows, cols, generator=g) * 3 + 1).to(DEV)
    add = torch.randn(7, cols, generator=g).to(DEV)
    g1, b1, g2, b2 = [torch.randn(cols, generator=g).to(DEV) for _ in range(4)]
    y1 = torch.empty(rows, cols, device=DEV, dtype=dtype)
    y2 = torch.empty(rows, cols, device=DEV, dtype=dtype)
    _run(ops.layernorm(x, g1, b1, y1, in2=add, in2_period=7, g2=g2, b2=b2, y2=y2))
    xx = x.double() + add.double()[torch.arange(rows, device=DEV) % 7]
    n = (xx - xx.mean(-1, keepdim=True)) / torch.sqrt(xx.var(-1, unbiased=False, keepdim=True) + 1e-5)
    tol = 2e-6 if dtype == torch.float32 else 3e-3
    assert rel_l2(y1.float(), n * g1.double() + b1.double()) < tol
    assert rel_l2(y2.float(), n * g2.double() + b2.double()) < tol
    if dtype == torch.float32:      # in place
        _run(ops.layernorm(x, g1, b1, x))
        n = (xx - add.double()[torch.arange(rows, device=DEV) % 7])
        n = (n - n.mean(-1, keepdim=True)) / torch.sqrt(n.var(-1, unbiased=False, keepdim=True) + 1e-5)
        assert rel_l2(x, n * g1.double() + b1.double()) < 2e-6


def _attention_ref(qkv, frames, tokens, heads):
    D = heads * 64
    q, k, v = qkv.double().view(frames, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)
    att = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    return (att @ v).permute(0, 2, 1, 3).reshape(frames, tokens, D)


@pytest.mark.parametrize("tokens,heads,frames", [(197, 12, 3), (257, 16, 2), (17, 2, 5), (50, 1, 1)])
@pytest.mark.parametrize("mode", ["simt_fp32", "simt_bf16", "tc_bf16"])
def test_attention(tokens, heads, frames, mode):
    ops = _ops()
    dtype = torch.float32 if mode == "simt_fp32" else torch.bfloat16
    g = torch.Generator().manual_seed(tokens)
    qkv = (torch.randn(frames, tokens, 3 * heads * 64, generator=g) * 1.5).to(DEV, dtype)
    out = torch.full((frames, tokens, heads * 64), float("nan"), device=DEV, dtype=dtype)
    impl = ops.IMPL_SIMT if mode.startswith("simt") else ops.IMPL_AUTO
    _run(ops.attention(qkv, out, frames, tokens, heads, impl=impl))
    exp = _attention_ref(qkv, frames, tokens, heads)
    assert rel_l2(out.float(), exp) < (3e-6 if dtype == torch.float32 else 6e-3)


@pytest.mark.parametrize("tokens,heads,frames", [(197, 12, 64), (257, 16, 40), (16, 1, 3), (33, 2, 200), (100, 3, 7), (128, 2, 90), (129, 2, 90),
                                                 (225, 4, 50), (241, 4, 50), (273, 4, 50), (288, 2, 80)])
def test_attention_tc_layouts(tokens, heads, frames):
    """Every TMEM plan of the tensor-core kernel (two slots + dedicated O, O in the other slot, odd key on the CUDA cores, one slot),
    several items per CTA so that the slot / ring phases wrap; rows with a dominant key (exponent range) and NaN-free output."""
    ops = _ops()
    g = torch.Generator().manual_seed(tokens * 7 + heads)
    qkv = (torch.randn(frames, tokens, 3 * heads * 64, generator=g) * 1.5)
    qkv[:, tokens // 2, :heads * 64] *= 6.0                    # a query row with a very peaked distribution
    qkv = qkv.to(DEV, torch.bfloat16)
    out = torch.full((frames, tokens, heads * 64), float("nan"), device=DEV, dtype=torch.bfloat16)
    call = ops.attention(qkv, out, frames, tokens, heads)
    _run(call)
    exp = _attention_ref(qkv, frames, tokens, heads)
    assert bool(torch.isfinite(out.float()).all())
    assert rel_l2(out.float(), exp) < 6e-3
    first = out.clone()
    _run(call)
    assert torch.equal(first, out)                             # replay is bit-identical


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("keys,batch", [(197, 9), (8, 3), (257, 2)])
def test_cross_attention(dtype, keys, batch):
    ops = _ops()
    heads, C = 6, 384
    g = torch.Generator().manual_seed(keys)
    q = torch.randn(batch, C, generator=g).to(DEV, dtype)
    kv = torch.randn(batch, keys, 2 * C, generator=g).to(DEV, dtype)
    out = torch.empty(batch, C, device=DEV, dtype=dtype)
    _run(ops.cross_attention(q, kv, out, batch, keys, heads))
    qh = q.double().view(batch, heads, 1, 64)
    k = kv.double()[..., :C].view(batch, keys, heads, 64).transpose(1, 2)
    v = kv.double()[..., C:].view(batch, keys, heads, 64).transpose(1, 2)
    exp = (torch.softmax(qh @ k.transpose(-1, -2) / 8.0, -1) @ v).reshape(batch, C)
    assert rel_l2(out.float(), exp) < (3e-6 if dtype == torch.float32 else 6e-3)


@pytest.mark.parametrize("p,res,pitch", [(16, 64, 768), (14, 56, 592), (16, 224, 768)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_patchify(p, res, pitch, dtype):
    ops = _ops()
    from oracle import dist_oracle
    clips, T, step = 2, 6, 2
    g = torch.Generator().manual_seed(p)
    video = torch.randn(clips, 3, T, res, res, generator=g).to(DEV)
    n_sel = T // step
    P = (res // p) ** 2
    out = torch.full((clips * n_sel * P, pitch), float("nan"), device=DEV, dtype=dtype)
    _run(ops.patchify(video, out, clips, T, res, res, p, 0, step, n_sel, pitch))
    frames = video[:, :, ::step].permute(0, 2, 1, 3, 4).reshape(-1, 3, res, res)
    exp = dist_oracle.patchify(frames.cpu(), p).reshape(-1, 3 * p * p)
    assert torch.equal(out[:, :3 * p * p].float().cpu(), exp.to(dtype).float())     # a pure permutation (+ cast)
    assert (out[:, 3 * p * p:] == 0).all()


def test_small_kernels():
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    dst = torch.randn(10, 5, 16, generator=g).to(DEV)
    table = torch.randn(4, 16, generator=g).to(DEV)
    before = dst.clone()
    _run(ops.rows_bcast(dst, 5 * 16, 10, 16, table, 4, True))
    exp = before.clone()
    exp[:, 0] += table[torch.arange(10) % 4]
    assert torch.allclose(dst, exp)
    _run(ops.rows_bcast(dst, 5 * 16, 10, 16, table, 1, False))
    exp[:, 0] = table[0]
    assert torch.allclose(dst, exp)
    src = torch.randn(3 * 4, 7, 32, generator=g).to(DEV)        # [(b, i), token, col]
    out = torch.empty(3, 32, device=DEV)
    _run(ops.mean_rows(src, 7 * 32, 4, 3, 32, out))
    assert torch.allclose(out, src[:, 0].view(3, 4, 32).mean(1), atol=1e-6)
    emb = torch.randn(5, 64, generator=g).to(DEV)
    text = torch.randn(11, 64, generator=g).to(DEV)
    text_n = text / text.norm(dim=1, keepdim=True)
    logits, probs = torch.empty(5, 11, device=DEV), torch.empty(5, 11, device=DEV)
    _run(ops.class_head(emb, text_n, 14.2857, 5, 64, 11, logits, probs))
    exp = 14.2857 * (emb.double() / emb.double().norm(dim=1, keepdim=True)) @ text_n.double().t()
    assert rel_l2(logits, exp) < 2e-6 and rel_l2(probs, torch.softmax(exp, -1)) < 2e-6


@pytest.mark.parametrize("p,dt", [(16, torch.float32), (16, torch.bfloat16), (14, torch.float32), (14, torch.bfloat16)])
def test_patchify_u8_matches_to_tensor_normalize_patchify(p, dt):
    """uint8 frames [b, T, H, W, 3] -> the patch rows of (x / 255 - mean) / std (torchvision to_tensor + normalize, ssv2.py:137-143)."""
    ops = _ops()
    b, T, R = 2, 4, 8 * p
    g = torch.Generator().manual_seed(p)
    frames = torch.randint(0, 256, (b, T, R, R, 3), generator=g, dtype=torch.uint8).to(DEV)
    mean, std = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)
    # the reference normalises on the CPU inside the data loader (true division; CUDA torch multiplies by 1/255 instead)
    clip = frames.cpu().float().permute(0, 4, 1, 2, 3) / 255.0                                    # [b, 3, T, H, W]
    clip = ((clip - torch.tensor(mean)[None, :, None, None, None]) / torch.tensor(std)[None, :, None, None, None]).contiguous().to(DEV)
    ld = (3 * p * p + 7) // 8 * 8
    for first, step, n_sel in ((0, 1, T), (0, 2, T // 2)):
        want = torch.full((b * n_sel * 64, ld), 3.0, device=DEV, dtype=dt)
        got = torch.full((b * n_sel * 64, ld), 3.0, device=DEV, dtype=dt)
        _run(ops.patchify(clip, want, b, T, R, R, p, first, step, n_sel, ld))
        _run(ops.patchify_u8(frames, got, b, T, R, R, p, first, step, n_sel, ld, mean, std))
        assert torch.equal(got, want)                                                             # same operation order: bit-exact


@pytest.mark.parametrize("M,N,K,gelu", [(1000, 2304, 768, False), (777, 3072, 768, True), (300, 256, 1024, False), (130, 96, 128, True)])
@pytest.mark.parametrize("impl", ["tcgen05", "simt"])
def test_gemm_with_folded_layernorm(M, N, K, gelu, impl):
    """LayerNorm(x) W^T + b computed as rstd * (x_bf16 (W*gamma)^T - mean * rowsum(W*gamma)) + (W beta + b): row_stats + GEMM epilogue."""
    ops = _ops()
    g = torch.Generator().manual_seed(M + N)
    x = (1.5 * torch.randn(M, K, generator=g) + 0.3).to(DEV)
    x[:, 7] *= 20.0                                             # an outlier channel
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    gamma, beta = (1 + 0.2 * torch.randn(K, generator=g)).to(DEV), (0.1 * torch.randn(K, generator=g)).to(DEV)
    xb = x.to(torch.bfloat16)
    wf = (W * gamma[None, :]).to(torch.bfloat16).contiguous()
    wsum = wf.float().sum(dim=1).contiguous()
    bias = (W.double() @ beta.double() + b.double()).float().contiguous()
    stats = torch.empty(M, 2, device=DEV)
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    _run(ops.row_stats(xb, stats))
    _run(ops.gemm(xb, wf, N, K, bias=bias, out=out, ld_out=N, act=ops.ACT_QUICKGELU if gelu else ops.ACT_NONE, ln_stats=stats, ln_wsum=wsum,
                  impl=ops.IMPL_SIMT if impl == "simt" else ops.IMPL_AUTO))
    xd = xb.double()                                            # the statistics are those of the bf16 rows the GEMM reads
    mu, var = xd.mean(dim=1, keepdim=True), xd.var(dim=1, unbiased=False, keepdim=True)
    assert rel_l2(stats[:, 0], mu[:, 0]) < 1e-5 and rel_l2(stats[:, 1], (var[:, 0] + 1e-5).rsqrt()) < 1e-5
    ref = ((xd - mu) / (var + 1e-5).sqrt() * gamma.double() + beta.double()) @ W.double().t() + b.double()
    if gelu:
        ref = _qgelu(ref)
    assert rel_l2(out, ref) < 6e-3                              # bf16 weights (gamma folded) and bf16 output


@pytest.mark.parametrize("M,N,K", [(1000, 768, 768), (777, 768, 3072), (300, 1024, 1024), (130, 384, 128), (50432, 768, 768)])
@pytest.mark.parametrize("impl", [0, 3])               # IMPL_AUTO (CTA pairs when the grid is large enough), IMPL_TCGEN05_1CTA
def test_gemm_emits_row_statistics(M, N, K, impl):
    """distb200_gemm_desc.stat_partials + distb200_row_stats_finalize: (mean, rstd) of exactly the bf16 rows written to out2,
    next to the unchanged fp32 / bf16 outputs; a second launch into the same slots gives the same bits (plain stores, no atomics)."""
    ops = _ops()
    g = torch.Generator().manual_seed(M + N)
    a = (torch.randn(M, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(DEV)
    w = torch.randn(N, K, generator=g).to(torch.bfloat16).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = (3.0 * torch.randn(M, N, generator=g) + 1.5).to(DEV)        # a non-zero row mean: exercises the one-pass variance
    plain, plain_b = res.clone(), torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    _run(ops.gemm(a, w, N, K, bias=bias, res=plain, ld_res=N, out=plain, ld_out=N, out2=plain_b, ld_out2=N, impl=impl))
    out, out_b = res.clone(), torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    parts = torch.zeros(M, ops.STAT_SLOTS, 2, device=DEV)
    stats = torch.empty(M, 2, device=DEV)
    call = ops.gemm(a, w, N, K, bias=bias, res=out, ld_res=N, out=out, ld_out=N, out2=out_b, ld_out2=N, stat_partials=parts, impl=impl)
    _run(call)
    _run(ops.row_stats_finalize(parts, N, stats))
    assert torch.equal(out, plain) and torch.equal(out_b, plain_b)
    xd = out_b.double()
    mu, var = xd.mean(dim=1), xd.var(dim=1, unbiased=False)
    assert rel_l2(stats[:, 0], mu) < 1e-5, rel_l2(stats[:, 0], mu)
    assert rel_l2(stats[:, 1], (var + 1e-5).rsqrt()) < 2e-5, rel_l2(stats[:, 1], (var + 1e-5).rsqrt())
    first = parts.clone()
    out.copy_(res)
    _run(call)
    assert torch.equal(parts, first)


@pytest.mark.parametrize("impl", ["tcgen05", "simt"])
def test_gemm_activation_on_a_column_suffix_and_bcast_copy(impl):
    """distb200_gemm_desc.act_from / act_to: QuickGELU on a column range only (folded IntegrationNetwork: activated ffn block |
    linear temporal block out of one GEMM); rows_bcast's optional bf16 copy of the rows it produced."""
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    M, N, K, cut = 333, 480, 384, 96
    a = (torch.randn(M, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(DEV)
    w = torch.randn(N, K, generator=g).to(torch.bfloat16).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    wide = torch.zeros(M, 576, device=DEV, dtype=torch.bfloat16)
    _run(ops.gemm(a, w, N, K, bias=bias, out=wide, ld_out=576, act=ops.ACT_QUICKGELU, act_from=cut,
                  impl=ops.IMPL_SIMT if impl == "simt" else ops.IMPL_AUTO))
    ref = a.double() @ w.double().t() + bias.double()
    ref[:, cut:] = _qgelu(ref[:, cut:])
    assert rel_l2(wide[:, :cut], ref[:, :cut]) < 4e-3 and rel_l2(wide[:, cut:N], ref[:, cut:]) < 6e-3
    assert float(wide[:, N:].abs().max()) == 0.0
    lo, hi = 64, 416                                                  # ... and on an inner column range
    _run(ops.gemm(a, w, N, K, bias=bias, out=wide, ld_out=576, act=ops.ACT_QUICKGELU, act_from=lo, act_to=hi,
                  impl=ops.IMPL_SIMT if impl == "simt" else ops.IMPL_AUTO))
    ref = a.double() @ w.double().t() + bias.double()
    ref[:, lo:hi] = _qgelu(ref[:, lo:hi])
    assert rel_l2(wide[:, :lo], ref[:, :lo]) < 4e-3 and rel_l2(wide[:, lo:hi], ref[:, lo:hi]) < 6e-3 and rel_l2(wide[:, hi:N], ref[:, hi:]) < 4e-3
    dst = torch.randn(10, 5 * 16, generator=g).to(DEV)
    want = dst.clone()
    table = torch.randn(4, 16, generator=g).to(DEV)
    copy = torch.zeros(10, 5 * 16, device=DEV, dtype=torch.bfloat16)
    _run(ops.rows_bcast(dst, 5 * 16, 10, 16, table, 4, True, dst2=copy, row_stride2=5 * 16))
    want[:, :16] += table[torch.arange(10) % 4]
    assert torch.equal(dst, want) and torch.equal(copy[:, :16], want[:, :16].to(torch.bfloat16)) and float(copy[:, 16:].abs().max()) == 0.0


@pytest.mark.parametrize("impl", ["tcgen05", "simt"])
def test_gemm_places_out2_independently(impl):
    """distb200_gemm_desc.out2_gdiv / out2_cstep / out2_gstride / out2_roff: the bf16 copy of output row (gi, r) lands at row
    (gi / g) * gstride + roff + r, column offset (gi % g) * cstep of a wider buffer (TemporalNet's (1,3,3) convolution feeding
    the K-concatenated operand), while `out` keeps its own rows."""
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    groups, rpg, N, K, gdiv, gstride, roff, pitch, col0 = 6, 49, 96, 96, 2, 50, 1, 512, 128
    a = (torch.randn(groups * rpg, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(DEV)
    w = torch.randn(N, K, generator=g).to(torch.bfloat16).to(DEV)
    out = torch.zeros(groups * rpg, N, device=DEV)
    wide = torch.zeros((groups // gdiv) * gstride, pitch, device=DEV, dtype=torch.bfloat16)
    _run(ops.gemm(a, w, N, K, a_dim=(K, rpg, groups, 1), a_stride=(1, K, rpg * K, groups * rpg * K), groups=groups, rows_per_group=rpg,
                  out=out, ld_out=N, out2=wide[:, col0:], ld_out2=pitch, out2_gdiv=gdiv, out2_cstep=N, out2_gstride=gstride, out2_roff=roff,
                  impl=ops.IMPL_SIMT if impl == "simt" else ops.IMPL_AUTO))
    ref = (a.double() @ w.double().t()).cpu()
    assert rel_l2(out, ref) < 1e-5
    want = torch.zeros_like(wide, dtype=torch.float64).cpu()
    for gi in range(groups):
        r0 = (gi // gdiv) * gstride + roff
        c0 = col0 + (gi % gdiv) * N
        want[r0:r0 + rpg, c0:c0 + N] = ref[gi * rpg:(gi + 1) * rpg]
    got = wide.double().cpu()
    assert torch.equal(got == 0, want == 0) or float((got - want).abs().max()) < 0.05      # untouched cells stay zero
    assert rel_l2(got, want) < 4e-3
