"""Shared helpers for the parity tests: rebuild the synthetic inputs a fixture was generated from."""
import torch

from dist_b200.arch import DistArch
from dist_b200.utils import synth


def inputs_for(fix):
    arch = DistArch(**fix["arch"]).validate()
    sd = synth.synth_state_dict(arch, seed=fix["weight_seed"], init=fix["init"])
    clips = synth.synth_clips(fix["batch"], arch, seed=fix["clip_seed"], kind=fix["clip_kind"])
    text = synth.synth_text_features(arch.num_classes, arch.embed_dim, seed=fix["text_seed"])
    return arch, sd, clips, text


def check_inputs(fix, sd, clips, text, rtol=1e-6):
    """The seeded generators must reproduce the tensors the reference was run on."""
    def close(a, b):
        # [sum, sum of magnitudes]: the signed sum of a zero-mean tensor cancels to ~0, so both entries are compared on the scale of the
        # magnitude sum (the generators normalise with reductions whose summation order depends on the host's thread count: 1e-8 relative)
        scale = max(1.0, abs(b[1]))
        return all(abs(x - y) <= rtol * scale for x, y in zip(a, b))
    assert close(synth.checksum({k: v for k, v in sd.items() if k != "logit_scale"}), fix["weights_checksum"]), "synthetic weights differ from the fixture's"
    assert close(synth.checksum(clips), fix["clips_checksum"]), "synthetic clips differ from the fixture's"
    assert close(synth.checksum(text), fix["text_checksum"]), "synthetic label embeddings differ from the fixture's"


def gather_rows(a4, a_dim, tap, groups, rpg, img_w=0, group_dim=2):
    """Rows A_j(gi, r, :) of the operand addressing in include/distb200.h for one tap (zero outside the tensor), fp64."""
    dev = a4.device
    gi = torch.arange(groups, device=dev).repeat_interleave(rpg)
    r = torch.arange(rpg, device=dev).repeat(groups)
    d1, d2, d3 = tap
    if img_w > 0:
        c1, c2, c3 = r % img_w + d1, r // img_w + d2, gi + d3
    else:
        c1 = r + d1
        c2 = d2 + (gi if group_dim == 2 else 0)
        c3 = d3 + (gi if group_dim == 3 else 0)
        c2 = c2 if torch.is_tensor(c2) else torch.full_like(r, c2)
        c3 = c3 if torch.is_tensor(c3) else torch.full_like(r, c3)
    ok = (c1 >= 0) & (c1 < a_dim[1]) & (c2 >= 0) & (c2 < a_dim[2]) & (c3 >= 0) & (c3 < a_dim[3])
    rows = a4[c3.clamp(0, a_dim[3] - 1), c2.clamp(0, a_dim[2] - 1), c1.clamp(0, a_dim[1] - 1)].double()
    return rows * ok[:, None]


def gemm_reference(a4, a_dim, taps, b3, groups, rpg, img_w=0, group_dim=2):
    """fp64 restatement of the operand addressing in include/distb200.h.

    a4: logical A tensor indexed [c3, c2, c1, k] (a torch tensor of shape a_dim reversed);
    b3: [taps, n, k].  Returns acc [groups*rpg, n]."""
    K = a_dim[0]
    acc = torch.zeros(groups * rpg, b3.shape[1], dtype=torch.float64, device=a4.device)
    for j, tap in enumerate(taps):
        rows = gather_rows(a4, a_dim, tap, groups, rpg, img_w, group_dim)
        acc += rows[:, :K] @ b3[j].double()[:, :K].t()
    return acc


def wgrad_reference(a4, a_dim, taps, dy, groups, rpg, img_w=0, group_dim=2):
    """dw[j, n, k] = sum_rows dy[row, n] * A_j(row, k) in fp64; dy [groups*rpg, n] in output-row order."""
    K = a_dim[0]
    return torch.stack([dy.double().t() @ gather_rows(a4, a_dim, tap, groups, rpg, img_w, group_dim)[:, :K] for tap in taps])


def temporalnet_reference(x, u, alpha, gam, bet, w1, b1, w2, b2, emulate_bf16):
    """Definition of ``distb200_temporalnet`` (include/distb200.h) with torch's own convolutions, any float dtype.
    x [B,T,g,g,C]; u [B,T/alpha,g,g,C] or None; w1 [3,C,C] / w2 [9,C,C] as (tap, out, in).  ``emulate_bf16`` rounds the two MMA
    operands (LayerNorm output, first activation) to bf16 like the kernel.  tests/test_oracle.py ties it to the oracle's
    ``temporal_net`` (dist.py:48-65)."""
    import torch.nn.functional as F
    q = lambda v: v * torch.sigmoid(1.702 * v)
    r = (lambda t: t.to(torch.bfloat16).to(t.dtype)) if emulate_bf16 else (lambda t: t)
    xe = x if u is None else x + u.repeat_interleave(alpha, dim=1)
    C = x.shape[-1]
    y = r(F.layer_norm(xe, (C,), gam, bet, 1e-5)).permute(0, 4, 1, 2, 3)      # [B,C,T,g,g]
    k1 = w1.permute(1, 2, 0)[:, :, :, None, None]                            # [Cout,Cin,3,1,1]
    z = r(q(F.conv3d(y, k1, b1, padding=(1, 0, 0))))
    k2 = w2.reshape(3, 3, C, C).permute(2, 3, 0, 1)[:, :, None]              # [Cout,Cin,1,3,3]
    o = F.conv3d(z, k2, b2, padding=(0, 1, 1)).permute(0, 2, 3, 4, 1)
    return q(xe + o)
