"""Fused TemporalNet block (``distb200_temporalnet``) against an fp64 restatement of its definition
(models/module_zoo/branches/dist.py:48-65 plus the nearest-upsample add of dist.py:105,231)."""
import pytest
import torch

from conftest import rel_l2
from helpers import temporalnet_reference as reference

pytestmark = pytest.mark.gpu

DEV = "cuda"


CASES = {
    # name: (C, g, T, clips, alpha, with_u, max_ctas)
    "tiny": (32, 4, 4, 2, 2, True, 0),
    "tiny_a3_long_runs": (32, 6, 6, 3, 3, True, 2),
    "tiny_a1": (32, 4, 3, 2, 1, True, 4),
    "c64": (64, 5, 4, 2, 2, False, 0),
    "b16": (96, 14, 16, 3, 2, True, 0),
    "b16_long_runs": (96, 14, 8, 5, 2, True, 6),
    "l14": (96, 16, 8, 2, 4, True, 0),
    "l14_nou": (96, 16, 5, 3, 1, False, 10),
    "g22_short_last_band": (32, 22, 3, 1, 1, True, 0),      # bands of 4,3,...: rows past a short band must never be addressed
}


@pytest.mark.parametrize("name", list(CASES))
def test_temporalnet(name):
    from dist_b200 import ops
    C, g, T, B, alpha, with_u, max_ctas = CASES[name]
    gen = torch.Generator().manual_seed(sum(map(ord, name)))
    P = g * g
    x = torch.randn(B, T, g, g, C, generator=gen) * 1.5 + 0.3
    u = (torch.randn(B, T // alpha, g, g, C, generator=gen) * 0.5).to(torch.bfloat16) if with_u else None
    gam, bet = 1 + 0.2 * torch.randn(C, generator=gen), 0.2 * torch.randn(C, generator=gen)
    w1 = torch.randn(3, C, C, generator=gen) / (3 * C) ** 0.5
    w2 = torch.randn(9, C, C, generator=gen) / (9 * C) ** 0.5 * 2
    b1, b2 = 0.3 * torch.randn(C, generator=gen), 0.3 * torch.randn(C, generator=gen)
    w1b, w2b = w1.to(torch.bfloat16), w2.to(torch.bfloat16)

    f64 = lambda t: None if t is None else t.double()
    want = reference(f64(x), f64(u), alpha, f64(gam), f64(bet), w1b.double(), f64(b1), w2b.double(), f64(b2), True)
    exact = reference(f64(x), f64(u), alpha, f64(gam), f64(bet), w1b.double(), f64(b1), w2b.double(), f64(b2), False)

    d = lambda t: None if t is None else t.to(DEV).contiguous()
    out = torch.full((B, T, P, C), float("nan"), device=DEV)
    # out2: the K-concatenated placement the engine uses - frame f lands on row (f // alpha) * (P + 1) + 1 + p, column (f % alpha) * C
    ld2 = alpha * C + 8
    out2 = torch.full((B * T // alpha * (P + 1), ld2), 7.0, device=DEV, dtype=torch.bfloat16)
    xd, ud = d(x), d(u)
    x_before = xd.clone()
    call = ops.temporalnet(xd, d(gam), d(bet), d(w1b), d(b1), d(w2b), d(b2), clips=B, frames=T, grid=g, u=ud, alpha=alpha,
                           out=out, out2=out2, ld_out2=ld2, out2_gdiv=alpha, out2_cstep=C, out2_gstride=P + 1, out2_roff=1, max_ctas=max_ctas)
    call.launch(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(xd, x_before)
    got = out.view(B, T, g, g, C)
    assert torch.isfinite(got).all()
    e_emu, e_exact = rel_l2(got, want), rel_l2(got, exact)
    worst = float((got.double().cpu() - want).abs().max())
    print("%s: rel-L2 vs bf16-emulated fp64 %.3e, vs exact fp64 %.3e, max abs %.3e" % (name, e_emu, e_exact, worst))
    assert e_emu < 3e-3 and e_exact < 1e-2, (e_emu, e_exact)
    # per-frame error: a wrong halo / tap would concentrate in a few frames or rows
    per_frame = ((got.double().cpu() - want) ** 2).sum(dim=(2, 3, 4)).sqrt() / (want ** 2).sum(dim=(2, 3, 4)).sqrt()
    assert float(per_frame.max()) < 6e-3, per_frame
    # the bf16 copy: same values, placed behind a class row, dense frame k of a group in column block k; nothing else is touched
    o2 = out2.view(B * T // alpha, P + 1, ld2)
    assert bool((o2[:, 0] == 7.0).all()) and bool((o2[:, :, alpha * C:] == 7.0).all())
    copy = o2[:, 1:, :alpha * C].reshape(B, T // alpha, P, alpha, C).permute(0, 1, 3, 2, 4).reshape(B, T, P, C)
    assert torch.equal(copy, out.to(torch.bfloat16))


def test_temporalnet_single_outputs_and_determinism():
    """out only / out2 only (the last DiST layer needs no fp32 stream) and bit-identical repeats."""
    from dist_b200 import ops
    C, g, T, B = 96, 14, 4, 2
    gen = torch.Generator().manual_seed(5)
    P = g * g
    mk = lambda *s: torch.randn(*s, generator=gen).to(DEV)
    x = mk(B, T, P, C)
    gam, bet, b1, b2 = 1 + 0.1 * mk(C), 0.1 * mk(C), 0.1 * mk(C), 0.1 * mk(C)
    w1, w2 = (mk(3, C, C) / 17).to(torch.bfloat16), (mk(9, C, C) / 29).to(torch.bfloat16)
    s = torch.cuda.current_stream().cuda_stream
    outs = []
    for rep in range(2):
        o = torch.zeros(B, T, P, C, device=DEV)
        ops.temporalnet(x, gam, bet, w1, b1, w2, b2, clips=B, frames=T, grid=g, out=o).launch(s)
        outs.append(o)
    o2 = torch.zeros(B * T * P, C, device=DEV, dtype=torch.bfloat16)
    ops.temporalnet(x, gam, bet, w1, b1, w2, b2, clips=B, frames=T, grid=g, out2=o2, ld_out2=C).launch(s)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(o2.view(B, T, P, C), outs[0].to(torch.bfloat16))


def test_temporalnet_rejects_bad_arguments():
    from dist_b200 import ops
    C, g, T, B = 48, 4, 2, 1
    z = lambda *s, dt=torch.float32: torch.zeros(*s, device=DEV, dtype=dt)
    x = z(B, T, g * g, C)
    w1, w2 = z(3, C, C, dt=torch.bfloat16), z(9, C, C, dt=torch.bfloat16)
    call = ops.temporalnet(x, z(C), z(C), w1, z(C), w2, z(C), clips=B, frames=T, grid=g, out=z(B, T, g * g, C))
    with pytest.raises(ops.DistB200Error, match="channels=48"):
        call.launch(torch.cuda.current_stream().cuda_stream)
    C = 32
    x = z(B, T, g * g, C)
    w1, w2 = z(3, C, C, dt=torch.bfloat16), z(9, C, C, dt=torch.bfloat16)
    call = ops.temporalnet(x, z(C), z(C), w1, z(C), w2, z(C), clips=B, frames=T, grid=g, out=x)
    with pytest.raises(ops.DistB200Error, match="alias"):
        call.launch(torch.cuda.current_stream().cuda_stream)
