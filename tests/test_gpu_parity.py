"""End-to-end parity of the CUDA path against the fixtures produced by the unmodified reference.

Bars (BASELINE.json north_star): fp32 embedding within 1e-4 relative error, bf16 within 1e-2, identical top-1 per
clip.  Top-1 is compared on the class logits; a flip is only accepted when the reference's own top-1/top-2 logit
margin is smaller than the measured logit error (a numerical tie, reported in the assertion message).
"""
import pytest
import torch

from conftest import load_golden, rel_l2
from helpers import check_inputs, inputs_for

pytestmark = pytest.mark.gpu

FP32_BAR, BF16_BAR = 1e-4, 1e-2


def _engine(fix, precision, **kw):
    from dist_b200.engine import DistEngine
    arch, sd, clips, text = inputs_for(fix)
    check_inputs(fix, sd, clips, text)
    eng = DistEngine(sd, arch, fix["batch"], device="cuda", precision=precision, text_features=text, **kw)
    return eng, clips


def _check_top1(fix, logits, tag):
    ref = fix["logits"].double()
    got = logits.double().cpu()
    err = (got - ref).abs().max(dim=1).values
    top2 = ref.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    for i in range(ref.shape[0]):
        same = int(got[i].argmax()) == int(ref[i].argmax())
        assert same or margin[i] < 2 * err[i], "%s clip %d: top-1 %d vs reference %d, margin %.3e, logit error %.3e" % (
            tag, i, int(got[i].argmax()), int(ref[i].argmax()), float(margin[i]), float(err[i]))


@pytest.mark.parametrize("name", ["tiny_ref", "tiny_scaled", "tiny_a1", "tiny_a3", "tiny_skip", "b16_8x16_ref", "b16_8x16_scaled", "b16_8x16_iid"])
def test_fp32_path_matches_reference(name):
    fix = load_golden(name)
    eng, clips = _engine(fix, "fp32")
    emb = eng.forward(clips.cuda(), use_graph=False)
    torch.cuda.synchronize()
    assert rel_l2(emb, fix["emb"]) < FP32_BAR
    assert rel_l2(eng.logits, fix["logits"]) < FP32_BAR
    assert rel_l2(eng.probs, fix["probs"]) < FP32_BAR
    _check_top1(fix, eng.logits, name + "/fp32")


@pytest.mark.parametrize("name", ["tiny_ref", "tiny_scaled", "tiny_a1", "tiny_a3", "tiny_skip", "b16_8x16_ref", "b16_8x16_scaled", "b16_8x16_iid",
                                  "b16_32x64_k400", "l14_32x64_k400"])
def test_bf16_path_matches_reference(name):
    fix = load_golden(name)
    eng, clips = _engine(fix, "bf16")
    emb = eng.forward(clips.cuda(), use_graph=False)
    torch.cuda.synchronize()
    assert torch.isfinite(emb).all()
    assert rel_l2(emb, fix["emb"]) < BF16_BAR, rel_l2(emb, fix["emb"])
    _check_top1(fix, eng.logits, name + "/bf16")


def test_bf16_simt_and_tcgen05_agree():
    """Same bf16 operands through the FFMA kernels and through the tensor-core kernels."""
    from dist_b200 import ops
    fix = load_golden("tiny_scaled")
    e1, clips = _engine(fix, "bf16", gemm_impl=ops.IMPL_SIMT, attn_impl=ops.IMPL_SIMT)
    e2, _ = _engine(fix, "bf16")
    a = e1.forward(clips.cuda(), use_graph=False).clone()
    b = e2.forward(clips.cuda(), use_graph=False).clone()
    # different accumulation order -> different bf16 roundings of the intermediates; both sit ~4.5e-3 from fp32
    assert rel_l2(a, b) < 8e-3


def test_cuda_graph_replay_is_identical_and_clips_are_independent():
    fix = load_golden("tiny_scaled")
    eng, clips = _engine(fix, "bf16")
    eager = eng.forward(clips.cuda(), use_graph=False).clone()
    eng.capture()
    replay = eng.forward(clips.cuda(), use_graph=True).clone()
    assert torch.equal(eager, replay)
    # permuting the clips of a batch permutes the embeddings (no cross-clip leakage through halos / groups)
    flipped = eng.forward(clips.flip(0).cuda(), use_graph=True).clone()
    assert torch.equal(flipped.flip(0), replay)


def test_full_batch_properties_b16():
    """At BASELINE.json's batch (32 clips) there is no oracle run; check batch-size independence instead:
    the first two clips of a 32-clip batch give the embeddings of the 2-clip fixture run."""
    fix = load_golden("b16_8x16_ref")
    from dist_b200.engine import DistEngine
    from dist_b200.utils import synth
    arch, sd, clips, text = inputs_for(fix)
    big = synth.synth_clips(32, arch, seed=99, kind="structured")
    big[:2] = clips
    eng = DistEngine(sd, arch, 32, device="cuda", precision="bf16", text_features=text)
    emb = eng.forward(big.cuda(), use_graph=False)[:2].clone()
    assert rel_l2(emb, fix["emb"]) < BF16_BAR
    small = DistEngine(sd, arch, 2, device="cuda", precision="bf16", text_features=text)
    e2 = small.forward(clips.cuda(), use_graph=False).clone()
    assert rel_l2(emb, e2) < 1e-5


def test_registry_driven_model_end_to_end():
    """configs/projects/dist YAML -> build_model -> BaseVideoModel.forward, as runs/test.py:92 calls it."""
    import os
    from conftest import ROOT
    import dist_b200.models.base  # noqa: F401
    from dist_b200.config import Config
    from dist_b200.models.base.builder import build_model
    fix = load_golden("b16_8x16_ref")
    arch, sd, clips, text = inputs_for(fix)
    cfg = Config.from_file(os.path.join(ROOT, "configs/projects/dist/ssv2/vit-b16-8+16f.yaml"), ["NUM_GPUS", "1"])
    model, _ = build_model(cfg)
    model.eval()
    enc = model.backbone.base_encoder
    enc.load_state_dict(sd, strict=True)
    preds, aux = model({"video": clips.cuda(), "texts": text.cuda()})
    assert preds.shape == (2, 174) and aux["logits_per_image"].shape == (2, 1, 174)
    assert rel_l2(aux["logits_per_image"][:, 0], fix["logits"]) < 2e-2
    assert rel_l2(preds, fix["probs"]) < 2e-2
    assert torch.allclose(preds.sum(-1), torch.ones(2, device="cuda"), atol=1e-4)
    emb = enc.forward_without_text(clips.permute(0, 2, 1, 3, 4).reshape(-1, 3, 224, 224).cuda())   # reference layout [B*T,3,H,W]
    assert emb.shape == (2, 1, 512) and rel_l2(emb[:, 0], fix["emb"]) < BF16_BAR


def test_uint8_clips_equal_normalised_float_clips():
    """Decoded uint8 frames through the fused normalise + patchify kernel give the same embedding, bit for bit, as the
    float clip torchvision's ToTensorVideo + NormalizeVideo would have produced."""
    from dist_b200.engine import DistEngine
    fix = load_golden("tiny_scaled")
    arch, sd, _, text = inputs_for(fix)
    g = torch.Generator().manual_seed(5)
    frames = torch.randint(0, 256, (2, arch.frames, arch.resolution, arch.resolution, 3), generator=g, dtype=torch.uint8)
    mean, std = torch.tensor(DistEngine.CLIP_MEAN), torch.tensor(DistEngine.CLIP_STD)
    clip = ((frames.float().permute(0, 4, 1, 2, 3) / 255.0 - mean[None, :, None, None, None]) / std[None, :, None, None, None]).contiguous()
    for precision in ("fp32", "bf16"):
        e_f = DistEngine(sd, arch, 2, precision=precision, text_features=text)
        e_u = DistEngine(sd, arch, 2, precision=precision, text_features=text, input_format="uint8")
        a = e_f.forward(clip.cuda(), use_graph=False).clone()
        b = e_u.forward(frames.cuda(), use_graph=False).clone()
        torch.cuda.synchronize()
        assert torch.equal(a, b), precision


def _reference_meter(preds, labels, clip_ids, num_videos, num_clips, num_cls, method, ks=(1, 5)):
    """The reference's TestMeter.update_stats / finalize_metrics loops (utils/meters.py:83-115,135-163) on CPU tensors."""
    video_preds = torch.zeros(num_videos, num_cls)
    video_labels = torch.zeros(num_videos).long()
    clip_count = torch.zeros(num_videos).long()
    for p, l, c in zip(preds, labels, clip_ids):
        for ind in range(p.shape[0]):
            vid = int(c[ind]) // num_clips
            video_labels[vid] = l[ind]
            if method == "sum":
                video_preds[vid] += p[ind]
            else:
                video_preds[vid] = torch.max(video_preds[vid], p[ind])
            clip_count[vid] += 1
    top = video_preds.topk(max(ks), dim=1).indices                       # utils/metrics.py topks_correct
    hits = top.eq(video_labels[:, None])
    return video_preds, video_labels, clip_count, [int(hits[:, :k].any(dim=1).sum()) for k in ks]


@pytest.mark.parametrize("method", ["sum", "max"])
def test_test_meter_matches_reference_loop(method):
    from dist_b200.meters import TestMeter
    g = torch.Generator().manual_seed(11)
    V, K, C = 37, 3, 174
    order = torch.randperm(V * K, generator=g)[: V * K - 2]              # two clips never arrive
    batches = order.split(16)
    labels_all = torch.randint(0, C, (V,), generator=g)
    meter = TestMeter(None, V, K, C, len(batches), ensemble_method=method)
    P, L, I = [], [], []
    for ids in batches:
        p = torch.softmax(2 * torch.randn(len(ids), C, generator=g), dim=-1)
        for j, cid in enumerate(ids.tolist()):                            # make the true class likely
            p[j, labels_all[cid // K]] += 0.2
        P.append(p), L.append(labels_all[ids // K]), I.append(ids)
        meter.update_stats(p.cuda(), labels_all[ids // K].cuda(), ids.cuda())
    stats = meter.finalize_metrics()
    vp, vl, cc, hits = _reference_meter(P, L, I, V, K, C, method)
    assert rel_l2(meter.video_preds, vp) < 1e-6
    assert torch.equal(meter.video_labels.cpu(), vl) and torch.equal(meter.clip_count.cpu(), cc)
    assert stats["top1_acc"] == "{:.2f}".format(hits[0] / V * 100.0) and stats["top5_acc"] == "{:.2f}".format(hits[1] / V * 100.0)


def test_perform_test_multi_view_loop():
    """runs/test.py semantics end to end on the tiny model: 3 views per video, uint8 clips, scores summed per video."""
    import dist_b200.models.base  # noqa: F401
    from dist_b200.engine import DistEngine
    from dist_b200.meters import TestMeter
    from dist_b200.runs.test import perform_test
    fix = load_golden("tiny_scaled")
    arch, sd, _, text = inputs_for(fix)
    g = torch.Generator().manual_seed(2)
    V, K, B = 4, 3, 2
    frames = torch.randint(0, 256, (V * K, arch.frames, arch.resolution, arch.resolution, 3), generator=g, dtype=torch.uint8)
    labels = torch.randint(0, arch.num_classes, (V,), generator=g)
    eng = DistEngine(sd, arch, B, precision="fp32", text_features=text, input_format="uint8")

    class Model(torch.nn.Module):
        def forward(self, x):
            eng.forward(x["video"], use_graph=False)
            return eng.probs.clone(), None

    loader = [({"video": frames[i:i + B]}, {"supervised": labels[torch.arange(i, i + B) // K]}, torch.arange(i, i + B), {}) for i in range(0, V * K, B)]
    meter = TestMeter(None, V, K, arch.num_classes, len(loader))
    stats = perform_test(loader, Model(), meter, None)
    want = torch.zeros(V, arch.num_classes)
    for i in range(0, V * K, B):
        eng.forward(frames[i:i + B].cuda(), use_graph=False)
        for j in range(B):
            want[(i + j) // K] += eng.probs[j].cpu()
    assert rel_l2(meter.video_preds, want) < 1e-6
    assert bool((meter.clip_count == K).all())
    top1 = float((want.argmax(dim=1) == labels).float().mean()) * 100
    assert stats["top1_acc"] == "{:.2f}".format(top1)


def test_stand_alone_vit_and_dist_modules():
    """The module-level surface of the reference: VisionTransformer.forward fills others["mid_feat"]["img"][l] (clip.py:263-300,177)
    and DiSTNetwork.forward(others) consumes those taps plus others["images"] (dist.py:222-247) - each on its own, against
    the fixture the reference produced (B/16 8+16f)."""
    import os
    from conftest import ROOT
    import dist_b200.models.base  # noqa: F401
    from dist_b200.config import Config
    from dist_b200.models.base.builder import build_model
    fix = load_golden("b16_8x16_ref")
    arch, sd, clips, text = inputs_for(fix)
    cfg = Config.from_file(os.path.join(ROOT, "configs/projects/dist/ssv2/vit-b16-8+16f.yaml"), ["NUM_GPUS", "1"])
    model, _ = build_model(cfg)
    enc = model.backbone.base_encoder
    enc.load_state_dict(sd, strict=True)
    frames = clips.permute(0, 2, 1, 3, 4).reshape(-1, 3, 224, 224).cuda()                  # backbone.py:233
    others = {"mid_feat": {"img": {}}}
    cls_x, x_logits, patches, others = enc.visual(frames, others)
    b, t = 2, arch.sparse_frames
    assert cls_x.shape == (b * t, arch.embed_dim) and patches.shape == (b * t, arch.patches, arch.width)
    rs, cs = fix["sample_stride"]
    for l in range(arch.layers):
        tap = others["mid_feat"]["img"][l]                                                   # [N, b*t, D]
        assert tap.shape == (arch.tokens, b * t, arch.width)
        ref = fix["parts"]["tap.%d" % l]                                                     # frame-major, sub-sampled
        assert rel_l2(tap.permute(1, 0, 2)[..., ::rs, ::cs], ref) < 2e-2, l
    others["images"] = frames
    emb, others = enc.dist_net(others)
    assert emb.shape == (b, arch.embed_dim) and rel_l2(emb, fix["emb"]) < BF16_BAR


def test_stand_alone_attention_block_protocol():
    """ATTEN_BLOCK_REGISTRY["ResidualAttentionBlockMid"](d_model, n_head, None, cfg=, layer_id=).forward((x, others)) writes the tap
    and matches the restated block (clip.py:170-178) on the fixture's weights."""
    from dist_b200.models.base.clip import ATTEN_BLOCK_REGISTRY
    from oracle import dist_oracle
    fix = load_golden("tiny_scaled")
    arch, sd, _, _ = inputs_for(fix)
    blk = ATTEN_BLOCK_REGISTRY.get("ResidualAttentionBlockMid")(arch.width, arch.heads, None, cfg=None, layer_id=1)
    pre = "visual.transformer.resblocks.1."
    blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, strict=True)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(arch.tokens, 6, arch.width, generator=g)                                  # [N, b*t, D]
    others = {"mid_feat": {"img": {}}}
    y, others = blk.cuda()((x.cuda(), others))
    sd64 = {k: v.double() for k, v in sd.items()}
    h = x.double().permute(1, 0, 2)
    t = dist_oracle._ln(h, sd64[pre + "ln_1.weight"], sd64[pre + "ln_1.bias"])
    h = h + dist_oracle._mha(t, t, sd64, pre[:-1] + ".attn", arch.heads)
    t = dist_oracle._ln(h, sd64[pre + "ln_2.weight"], sd64[pre + "ln_2.bias"])
    h = h + dist_oracle._lin(dist_oracle._qgelu(dist_oracle._lin(t, sd64, pre + "mlp.c_fc")), sd64, pre + "mlp.c_proj")
    assert y.shape == x.shape and rel_l2(y.permute(1, 0, 2), h) < 1e-2
    assert torch.equal(others["mid_feat"]["img"][1], y)


def test_two_branch_graph_is_race_free_b16():
    """Optionally the captured graph runs the ViT and the DiST chains as parallel branches (engine.plan_branches).  At the full B/16
    geometry every kernel fills the GPU, so a missing edge would show as a changed embedding: replays must be bit-identical
    to the single-stream eager run, with fresh clips each time (the tap buffers are reused across replays)."""
    from dist_b200.arch import DistArch
    from dist_b200.engine import DistEngine
    from dist_b200.utils import synth
    arch = DistArch().validate()
    sd = synth.synth_state_dict(arch, seed=0, init="scaled")
    eng = DistEngine(sd, arch, 8, device="cuda", precision="bf16", text_features=synth.synth_text_features(arch.num_classes, arch.embed_dim))
    batches = [synth.synth_clips(8, arch, seed=s, kind="structured").cuda() for s in (1, 2, 3)]
    eager = [eng.forward(v, use_graph=False).clone() for v in batches]
    eng.capture(branches=2)
    assert eng.graph_branches == 2
    for rep in range(3):
        for v, want in zip(batches, eager):
            got = eng.forward(v, use_graph=True).clone()
            assert torch.equal(got, want), "replay %d differs from the eager run" % rep
