"""Kernel-level parity of the fine-tuning operators (include/distb200.h, second half) against fp64 torch definitions."""
import math

import pytest
import torch

from conftest import rel_l2
from helpers import wgrad_reference

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from dist_b200 import ops
    return ops


def _run(*calls):
    for c in calls:
        c.launch(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()


def _qgelu(x):
    return x * torch.sigmoid(1.702 * x)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n", [4096, 1003, 16000024])
def test_quickgelu_fwd_bwd(dt, n):
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    z = (2 * torch.randn(n, generator=g)).to(dt).to(DEV)
    dy = torch.randn(n, generator=g).to(dt).to(DEV)
    y32, ylp = torch.empty(n, device=DEV), torch.empty(n, device=DEV, dtype=dt)
    dz32, dzlp = torch.empty(n, device=DEV), torch.empty(n, device=DEV, dtype=dt)
    _run(ops.quickgelu(z, y32, ylp), ops.quickgelu_bwd(dy, z, dz32, dzlp))
    zz = z.double().requires_grad_(True)
    ref = _qgelu(zz)
    ref.backward(dy.double())
    assert rel_l2(y32, ref.detach()) < 1e-6
    assert rel_l2(dz32, zz.grad) < 1e-6
    tol = 1e-6 if dt == torch.float32 else 4e-3
    assert rel_l2(ylp, ref.detach()) < tol and rel_l2(dzlp, zz.grad) < tol
    # low-precision outputs only (the tensor-core training path): the one-MUFU sigmoid
    ylp.zero_(); dzlp.zero_()
    _run(ops.quickgelu(z, None, ylp), ops.quickgelu_bwd(dy, z, None, dzlp))
    assert rel_l2(ylp, ref.detach()) < tol and rel_l2(dzlp, zz.grad) < tol


def test_cast_and_group_sum():
    ops = _ops()
    x = torch.randn(3 * 4 * 5 * 8, device=DEV)
    y = torch.empty_like(x, dtype=torch.bfloat16)
    _run(ops.cast(x, y))
    assert torch.equal(y, x.to(torch.bfloat16))
    src = torch.randn(6, 3, 40, device=DEV)          # 6 groups x alpha 3 x inner 40
    dst = torch.empty(6, 40, device=DEV)
    _run(ops.group_sum(src, dst, 6, 3, 40))
    assert rel_l2(dst, src.double().sum(dim=1)) < 1e-6


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_colsum(dt):
    ops = _ops()
    src = torch.randn(12, 17, 100, device=DEV).to(dt)        # 12 frames x 17 tokens x 100 cols (pitch 100)
    out = torch.zeros(100, device=DEV)
    _run(ops.colsum(src, out, 100))                           # all rows
    assert rel_l2(out, src.double().sum(dim=(0, 1))) < 1e-5
    out = torch.ones(96, device=DEV)                          # accumulates; patch rows only, 96 of 100 columns
    _run(ops.colsum(src, out, 96, ld=100, groups=12, rows_per_group=16, gstride=17, roff=1))
    assert rel_l2(out, 1 + src.double()[:, 1:, :96].sum(dim=(0, 1))) < 1e-5
    out = torch.zeros(4, 100, device=DEV)                     # class rows per frame slot ti = frame % 4
    _run(ops.colsum(src, out, 100, groups=12, rows_per_group=1, gstride=17, roff=0, period=4))
    assert rel_l2(out, src.double()[:, 0].view(3, 4, 100).sum(dim=0)) < 1e-5
    big = torch.randn(50000, 96, device=DEV).to(dt)
    out = torch.zeros(96, device=DEV)
    _run(ops.colsum(big, out, 96))
    assert rel_l2(out, big.double().sum(dim=0)) < 1e-4


@pytest.mark.parametrize("cols,ld,rows", [(96, 96, 100352), (96, 104, 4099), (384, 384, 50432), (1536, 1536, 3001), (2048, 2048, 129), (8, 8, 5), (768, 776, 1)])
def test_colsum_dense_bf16(cols, ld, rows):
    """The contiguous-rows bf16 path (16-byte loads, cols / 8 threads per row): every width class, a row pitch wider than the columns,
    a row offset, the second output, accumulation into a non-zero buffer."""
    ops = _ops()
    g = torch.Generator().manual_seed(cols + rows)
    src = torch.randn(rows + 3, ld, generator=g).to(DEV, torch.bfloat16)
    out = torch.full((cols,), 2.0, device=DEV)
    out2 = torch.zeros(cols, device=DEV)
    _run(ops.colsum(src, out, cols, ld=ld, groups=1, rows_per_group=rows, gstride=rows, roff=3, out2=out2))
    want = src.double()[3:, :cols].sum(dim=0)
    scale = float(src.double()[3:, :cols].abs().sum(dim=0).max())
    assert float((out.double() - 2.0 - want).abs().max()) < 2e-6 * scale + 1e-6
    assert float((out2.double() - want).abs().max()) < 2e-6 * scale + 1e-6


@pytest.mark.parametrize("cols,tokens,frames", [(384, 197, 40), (96, 17, 300), (768, 50, 3)])
def test_colsum_grouped_bf16(cols, tokens, frames):
    """Patch rows only (the class-token row of every frame skipped): groups = frames, rows_per_group = tokens - 1, gstride = tokens, roff = 1."""
    ops = _ops()
    g = torch.Generator().manual_seed(cols + tokens)
    src = torch.randn(frames, tokens, cols, generator=g).to(DEV, torch.bfloat16)
    out = torch.zeros(cols, device=DEV)
    _run(ops.colsum(src, out, cols, groups=frames, rows_per_group=tokens - 1, gstride=tokens, roff=1))
    want = src.double()[:, 1:].sum(dim=(0, 1))
    scale = float(src.double()[:, 1:].abs().sum(dim=(0, 1)).max())
    assert float((out.double() - want).abs().max()) < 2e-6 * scale + 1e-6


@pytest.mark.parametrize("cols,rows", [(96, 1000), (384, 333), (768, 70), (128, 5)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_layernorm_bwd(cols, rows, dt):
    ops = _ops()
    g = torch.Generator().manual_seed(cols + rows)
    x = torch.randn(rows, cols, generator=g).to(DEV)
    x2 = torch.randn(4, cols, generator=g).to(DEV)
    g1, g2 = (1 + 0.3 * torch.randn(cols, generator=g)).to(DEV), (1 + 0.3 * torch.randn(cols, generator=g)).to(DEV)
    dy1, dy2 = torch.randn(rows, cols, generator=g).to(dt).to(DEV), torch.randn(rows, cols, generator=g).to(dt).to(DEV)
    add = torch.randn(rows, cols, generator=g).to(DEV)
    dx0 = torch.randn(rows, cols, generator=g).to(DEV)
    # reference
    xin = (x.double() + x2.double()[torch.arange(rows, device=DEV) % 4]).requires_grad_(True)
    ga, gb = g1.double().requires_grad_(True), g2.double().requires_grad_(True)
    ba, bb = torch.zeros(cols, device=DEV, dtype=torch.float64, requires_grad=True), torch.zeros(cols, device=DEV, dtype=torch.float64, requires_grad=True)
    y1 = torch.nn.functional.layer_norm(xin, (cols,), ga, ba, 1e-5)
    y2 = torch.nn.functional.layer_norm(xin, (cols,), gb, bb, 1e-5)
    (y1 * dy1.double()).sum().add((y2 * dy2.double()).sum()).backward()
    dx = dx0.clone()
    dxlp = torch.empty(rows, cols, device=DEV, dtype=torch.bfloat16)
    dg1, db1, dg2, db2 = (torch.zeros(cols, device=DEV) for _ in range(4))
    _run(ops.layernorm_bwd(x, g1, dy1, in2=x2, in2_period=4, g2=g2, dy2=dy2, add=add, dx=dx, accumulate=True, dx_lp=dxlp,
                           dg1=dg1, db1=db1, dg2=dg2, db2=db2))
    want = xin.grad + add.double() + dx0.double()
    assert rel_l2(dx, want) < 2e-5
    assert rel_l2(dxlp, want) < 4e-3
    for got, ref in ((dg1, ga.grad), (db1, ba.grad), (dg2, gb.grad), (db2, bb.grad)):
        assert rel_l2(got, ref) < 2e-5
    # single affine, overwrite
    dx = torch.full((rows, cols), 9.0, device=DEV)
    _run(ops.layernorm_bwd(x, g1, dy1, dx=dx))
    xin = x.double().requires_grad_(True)
    torch.nn.functional.layer_norm(xin, (cols,), g1.double(), None, 1e-5).mul(dy1.double()).sum().backward()
    assert rel_l2(dx, xin.grad) < 2e-5


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("batch,keys,heads", [(5, 197, 6), (3, 8, 2), (2, 33, 1)])
def test_cross_attention_bwd(dt, batch, keys, heads):
    ops = _ops()
    C = heads * 64
    g = torch.Generator().manual_seed(keys)
    q = torch.randn(batch, C, generator=g).to(dt).to(DEV)
    kv = torch.randn(batch, keys, 2 * C, generator=g).to(dt).to(DEV)
    do = torch.randn(batch, C, generator=g).to(dt).to(DEV)
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    _run(ops.cross_attention_bwd(q, kv, do, dq, dkv, batch, keys, heads))
    qq, kk = q.double().requires_grad_(True), kv.double().requires_grad_(True)
    qh = qq.view(batch, heads, 1, 64)
    kh = kk[..., :C].reshape(batch, keys, heads, 64).transpose(1, 2)
    vh = kk[..., C:].reshape(batch, keys, heads, 64).transpose(1, 2)
    out = (torch.softmax(qh @ kh.transpose(-1, -2) / 8.0, dim=-1) @ vh).reshape(batch, C)
    out.backward(do.double())
    tol = 2e-5 if dt == torch.float32 else 8e-3
    assert rel_l2(dq, qq.grad) < tol and rel_l2(dkv, kk.grad) < tol


@pytest.mark.parametrize("batch,E,C", [(4, 512, 174), (2, 64, 10), (3, 768, 400)])
def test_softce_head(batch, E, C):
    ops = _ops()
    g = torch.Generator().manual_seed(E)
    emb = torch.randn(batch, E, generator=g).to(DEV)
    text = torch.randn(C, E, generator=g).to(DEV)
    text_n = (text / text.norm(dim=1, keepdim=True)).contiguous()
    target = torch.softmax(3 * torch.randn(batch, C, generator=g), dim=-1).to(DEV)
    scale = 1 / 0.07
    logits, loss, d_emb = torch.empty(batch, C, device=DEV), torch.zeros(1, device=DEV), torch.empty(batch, E, device=DEV)
    _run(ops.softce_head(emb, text_n, scale, target, batch, E, C, logits, loss, d_emb))
    e = emb.double().requires_grad_(True)
    lg = scale * (e / e.norm(dim=1, keepdim=True)) @ text_n.double().t()
    ls = torch.sum(-target.double() * torch.log_softmax(lg, dim=-1), dim=-1).mean()
    ls.backward()
    assert rel_l2(logits, lg.detach()) < 1e-5
    assert abs(float(loss) - float(ls.detach())) < 1e-5 * abs(float(ls.detach()))
    assert rel_l2(d_emb, e.grad) < 2e-5


def test_adamw_matches_torch():
    ops = _ops()
    n = 10007
    p0 = torch.randn(n, device=DEV)
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p], lr=3.2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    q, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 5):
        g = torch.randn(n, device=DEV)
        p.grad = g.clone()
        opt.step()
        ops.adamw(q, 2 * g, m, v, n, 3.2e-4, 0.9, 0.999, 1e-8, 1e-4, step, 0.5, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert torch.allclose(q, p.detach(), rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_pack_weight(dt):
    ops = _ops()
    w = torch.randn(3, 96, 588, device=DEV)
    out = torch.full((3, 96, 592), 5.0, device=DEV, dtype=dt)
    out_t = torch.full((3, 588, 96), 5.0, device=DEV, dtype=dt)
    _run(ops.pack_weight(w, 3, 96, 588, out=out, out_t=out_t, ld_out=592, ld_out_t=96))
    assert torch.equal(out[..., :588], w.to(dt)) and float(out[..., 588:].abs().max()) == 0
    assert torch.equal(out_t, w.transpose(1, 2).to(dt))


WG = {
    "plain": dict(a_shape=(1, 1, 1000, 200), k=200, n=72, taps=[(0, 0, 0)], groups=1, rpg=1000),
    "padded_k": dict(a_shape=(1, 2, 96, 592), k=588, n=96, taps=[((k - 2) * 16, 0, 0) for k in range(5)], groups=2, rpg=96),
    "conv_t": dict(a_shape=(1, 3, 4 * 49, 96), k=96, n=96, taps=[(-49, 0, 0), (0, 0, 0), (49, 0, 0)], groups=3, rpg=4 * 49),
    "conv_s14": dict(a_shape=(5, 14, 14, 96), k=96, n=96, img_w=14, taps=[(j - 1, i - 1, 0) for i in range(3) for j in range(3)], groups=5, rpg=196),
    "t2i": dict(a_shape=(6, 2, 196, 96), k=96, n=384, taps=[(0, 0, 0), (0, 1, 0)], groups=6, rpg=196, group_dim=3, dy_gstride=197, dy_roff=1,
                dy_rows=6 * 197),
    "i2t": dict(a_shape=(1, 6, 197, 384), k=384, n=96, taps=[(1, 0, 0)], groups=6, rpg=196),
    "wide": dict(a_shape=(1, 1, 3000, 768), k=768, n=384, taps=[(0, 0, 0)], groups=1, rpg=3000),
}


@pytest.mark.parametrize("name", sorted(WG))
@pytest.mark.parametrize("mode", ["fp32", "bf16_simt", "bf16"])
def test_wgrad(name, mode):
    ops = _ops()
    sc = WG[name]
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    impl = ops.IMPL_SIMT if mode == "bf16_simt" else ops.IMPL_AUTO
    g = torch.Generator().manual_seed(3)
    a = torch.randn(*sc["a_shape"], generator=g)
    k, n, taps, groups, rpg = sc["k"], sc["n"], sc["taps"], sc["groups"], sc["rpg"]
    a[..., k:] = 7.0
    a = a.to(dt).to(DEV)
    gs, roff = sc.get("dy_gstride", rpg), sc.get("dy_roff", 0)
    dy_full = (torch.randn(sc.get("dy_rows", groups * rpg), n, generator=g) / math.sqrt(groups * rpg)).to(dt).to(DEV)
    rows = (torch.arange(groups).repeat_interleave(rpg) * gs + roff + torch.arange(rpg).repeat(groups)).to(DEV)
    a_dim = (k,) + tuple(reversed(sc["a_shape"][:3]))
    st = a.stride()
    want = wgrad_reference(a, a_dim, taps, dy_full[rows], groups, rpg, sc.get("img_w", 0), sc.get("group_dim", 2))
    dw = torch.ones(len(taps), n, k, device=DEV)                   # accumulates onto what is there
    _run(ops.wgrad(a, dy_full, dw, n, k, a_dim=a_dim, a_stride=(1, st[2], st[1], st[0]), taps=taps, img_w=sc.get("img_w", 0), groups=groups,
                   rows_per_group=rpg, group_dim=sc.get("group_dim", 2), dy_gstride=gs, dy_roff=roff, impl=impl))
    assert rel_l2(dw - 1, want) < (2e-5 if mode == "fp32" else 2e-5)      # bf16 operands are exact inputs; accumulation is fp32
