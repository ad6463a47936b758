"""Fine-tuning step of the DiST path on the distb200 CUDA library.

What the reference does per iteration (``runs/train.py:97-112``): forward with the CLIP ViT frozen under ``no_grad``
(``models/base/clip.py:485-487``) and the DiST branches in autograd, ``SoftTargetCrossEntropy`` on the raw cosine
logits (``models/utils/losses.py:20-31``, ``base_blocks.py:579-585``), ``loss.backward()``, DDP gradient averaging
(``models/base/builder.py:72``) and ``torch.optim.AdamW`` over the ``dist_net`` tensors with the parameter groups of
``models/utils/optimizer.py:138-190``.

Here the same step is a static plan of prepared C calls:

  forward    the frozen ViT exactly as at inference (``DistEngine._plan_vit_layer``), every tap kept; the DiST layers
             with their pre-activations and GEMM operands kept per layer (the inference plan overwrites them)
  backward   hand-derived, in reverse layer order; every dense gradient is a GEMM: input gradients through
             ``distb200_gemm`` with the transposed operand copies and negated tap offsets, weight gradients through
             ``distb200_gemm_wgrad``; LayerNorm / QuickGELU / cross-attention / loss have their own kernels
  update     one flat fp32 master buffer (decayed tensors first), flat gradient / moment buffers, a single
             all-reduce of the gradient buffer over the process group (``utils/distributed.py:41-57`` semantics),
             ``distb200_adamw`` on the two decay classes, then ``distb200_pack_weight`` refreshes the GEMM operands.

Tensors that never receive a gradient (the integration->temporal branch of the last selected layer, whose output
nothing reads) are left untouched, as ``torch.optim.AdamW`` leaves parameters with ``grad is None``.
"""

import torch

from . import ops
from .arch import DistArch
from .engine import DistEngine, _pad8


# ---- master-weight layouts ---------------------------------------------------------------------------------------

def _pack_conv(name, w):
    """Reference conv weight -> [taps, N, K] (the layout of the GEMM B operand); returns (packed, unpack_fn)."""
    shp = tuple(w.shape)
    if name.endswith("temporal_stem.weight"):                      # [Ct, 3, kt, ps, ps]
        ct, c, kt, ps, _ = shp
        return (w.permute(2, 0, 1, 3, 4).reshape(kt, ct, c * ps * ps),
                lambda p: p.reshape(kt, ct, c, ps, ps).permute(1, 2, 0, 3, 4))
    if name.endswith("temporal_net.c_fc2.weight"):                 # [Ct, Ch, 1, 3, 3] -> tap = i*3 + j
        n, k = shp[0], shp[1]
        return (w[:, :, 0].permute(2, 3, 0, 1).reshape(9, n, k),
                lambda p: p.reshape(3, 3, n, k).permute(2, 3, 0, 1).unsqueeze(2))
    # [N, K, kt, 1, 1] (temporal kernels, alpha-strided fusion conv, 1x1x1 convs)
    n, k, kt = shp[0], shp[1], shp[2]
    return (w[:, :, :, 0, 0].permute(2, 0, 1), lambda p: p.permute(1, 2, 0).reshape(n, k, kt, 1, 1))


class Mat:
    """One weight matrix (or per-tap stack of matrices): master / gradient views and the two operand copies."""

    def __init__(self, w, g, batch, n, k, device, adt):
        self.w, self.g, self.batch, self.n, self.k = w, g, batch, n, k
        self.kp, self.np_ = _pad8(k), _pad8(n)
        self.f = torch.zeros(batch, n, self.kp, device=device, dtype=adt)        # forward operand  [tap][n][k]
        self.t = torch.zeros(batch, k, self.np_, device=device, dtype=adt)       # backward operand [tap][k][n]

    def pack_call(self, name):
        return ops.pack_weight(self.w, self.batch, self.n, self.k, out=self.f, out_t=self.t, ld_out=self.kp, ld_out_t=self.np_,
                               name="pack." + name)

    def rows(self, r0, r1):
        """View of output rows [r0, r1) of a single matrix (q / kv halves of a packed attention in-projection)."""
        assert self.batch == 1
        m = Mat.__new__(Mat)
        m.w, m.g, m.batch, m.n, m.k, m.kp, m.np_ = self.w[:, r0:r1], self.g[:, r0:r1], 1, r1 - r0, self.k, self.kp, self.np_
        m.f = self.f[:, r0:r1]
        m.t = self.t[:, :, r0:r1]
        return m


class Vec:
    def __init__(self, p, g):
        self.p, self.g = p, g

    def part(self, a, b):
        return Vec(self.p[a:b], self.g[a:b])


class ParamTable:
    """Flat fp32 master / gradient / Adam-moment buffers of the ``dist_net.*`` tensors, decayed tensors first."""

    def __init__(self, sd, arch: DistArch, device, adt, weight_decay):
        names = sorted(k for k in sd if k.startswith("dist_net."))
        i_last = len(arch.selected_layers) - 1
        self.unused = sorted("dist_net.integration2temporal_nets.%d.linear_fuse.%s" % (i_last, s) for s in ("weight", "bias"))
        packed, self.unpack = {}, {}
        for k in names:
            w = sd[k].detach().float().cpu()
            if w.dim() == 5:
                packed[k], self.unpack[k] = _pack_conv(k, w)
            else:
                packed[k], self.unpack[k] = w, (lambda p, shp=tuple(w.shape): p.reshape(shp))
            packed[k] = packed[k].contiguous()

        from .models.utils.optimizer import weight_decay_of

        def decay_of(k):                                            # the group rule of optimizer.py:150-166
            return weight_decay_of(k, sd[k].dim(), weight_decay)

        used = [k for k in names if k not in self.unused]
        from .models.utils.optimizer import decay_class
        decayed = [k for k in used if decay_class(k, sd[k].dim()) == "normal"]
        plain = [k for k in used if decay_class(k, sd[k].dim()) != "normal"]
        order = decayed + plain + list(self.unused)
        self.offset, off = {}, 0
        for group, attr in ((decayed, "n_decay"), (plain, "n_used"), (self.unused, "total")):
            for k in group:
                self.offset[k] = off
                off += (packed[k].numel() + 3) // 4 * 4            # 16-byte aligned tensors
            setattr(self, attr, off)
        self.weight_decay = weight_decay
        flat = torch.zeros(self.total, dtype=torch.float32)
        for k in order:
            flat[self.offset[k]:self.offset[k] + packed[k].numel()] = packed[k].reshape(-1)
        self.p = flat.to(device)
        self.g = torch.zeros_like(self.p)
        self.m = torch.zeros_like(self.p)
        self.v = torch.zeros_like(self.p)
        self.shape = {k: tuple(packed[k].shape) for k in order}
        self.names = order
        self.device, self.adt = device, adt
        self.mats = {}

    def _view(self, buf, k):
        n = 1
        for s in self.shape[k]:
            n *= s
        return buf[self.offset[k]:self.offset[k] + n].view(self.shape[k])

    def mat(self, k):
        if k not in self.mats:
            shp = self.shape[k]
            batch, n, kk = (1,) * (3 - len(shp)) + shp if len(shp) <= 3 else shp
            self.mats[k] = Mat(self._view(self.p, k).view(batch, n, kk), self._view(self.g, k).view(batch, n, kk), batch, n, kk,
                               self.device, self.adt)
        return self.mats[k]

    def vec(self, k, shape=None):
        p, g = self._view(self.p, k), self._view(self.g, k)
        if shape is not None:
            p, g = p.view(shape), g.view(shape)
        return Vec(p, g)

    def export(self, buf=None):
        """Reference-layout tensors (CPU fp32) of the masters (or of another flat buffer, e.g. the gradients)."""
        buf = self.p if buf is None else buf
        return {k: self.unpack[k](self._view(buf, k).detach().float().cpu()).contiguous() for k in self.names}


def reduce_gradients(pt, group=None):
    """SUM all-reduce of the used part of the flat gradient buffer (one collective for all ``dist_net`` tensors; the
    reference's DDP reduces buckets over every parameter of the model, ``models/base/builder.py:69-74``).  Returns the
    factor the update applies to the summed gradient (1 / world size = DDP's mean)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world == 1:
        return 1.0
    dist.all_reduce(pt.g[:pt.n_used], group=group)
    return 1.0 / world


# ---- the engine ----------------------------------------------------------------------------------------------------

class TrainEngine(DistEngine):
    """Planned forward + backward + AdamW for a fixed number of clips per step."""

    def __init__(self, state_dict, arch: DistArch, batch, device="cuda", precision="bf16", text_features=None, weight_decay=1e-4,
                 betas=(0.9, 0.999), eps=1e-8, process_group=None, gemm_impl=ops.IMPL_AUTO, attn_impl=ops.IMPL_AUTO):
        assert text_features is not None, "the training head needs the label embeddings"
        self.weight_decay, self.betas, self.eps = float(weight_decay), betas, float(eps)
        self.process_group = process_group
        self.step_count = 0
        self._sd = state_dict
        super().__init__(state_dict, arch, batch, device=device, precision=precision, text_features=text_features,
                         gemm_impl=gemm_impl, attn_impl=attn_impl)
        del self._sd
        self.fwd_graph = None
        self.repack()
        torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------------------------------
    def _alloc(self):
        super()._alloc()
        a, b, dev, adt = self.arch, self.batch, self.device, self.adt
        self.pt = ParamTable(self._sd, a, dev, adt, self.weight_decay)
        F, N, P, D, T, t = b * a.sparse_frames, a.tokens, a.patches, a.width, a.frames, a.sparse_frames
        Ci, Ct, Ch, Cm, Ih = a.integration_dim, a.temporal_dim, a.temporal_hidden, a.integration_temporal_hidden, a.integration_hidden
        Mv, Mt = F * N, b * T * P
        nl = len(a.selected_layers)
        f32 = torch.float32
        z = lambda *s, dtype=adt: torch.zeros(*s, device=dev, dtype=dtype)
        self.target = z(b, self.logits.shape[1], dtype=f32)
        self.loss = z(1, dtype=f32)
        self.d_emb = z(b, a.embed_dim, dtype=f32)
        self.d_emb_a = z(b, a.embed_dim)
        # ---- saved per DiST layer
        L = self.saved = []
        for i in range(nl):
            L.append(dict(
                tap=z(Mv, D), xT_in=z(Mt, Ct, dtype=f32), xln=z(Mt, Ct), z1=z(Mt, Ch), y1=z(Mt, Ch), z2=z(Mt, Ct, dtype=f32),
                xT_out=z(Mt, Ct, dtype=f32), xT_a=z(Mt, Ct), mid_a=z(Mv, Ci), upd=z(Mv, Ci, dtype=f32), a1=z(Mv, Ci), a2=z(Mv, Ci),
                zf=z(Mv, Ih), hf=z(Mv, Ih), tf1=z(Mv, Cm), zt=z(Mv, Cm), ht=z(Mv, Cm)))
        # ---- saved per ada-pooling layer
        A = self.saved_ada = []
        for j in range(a.ada_layers):
            A.append(dict(
                ln_cur=z(Mv, Ci), kv_s=z(Mv, 2 * Ci), sp_in=z(F, Ci, dtype=f32), sp_lnq=z(F, Ci), q_s=z(F, Ci), o_s=z(F, Ci),
                sp_mid=z(F, Ci, dtype=f32), sp_lno=z(F, Ci), z_ms=z(F, 4 * Ci), h_ms=z(F, 4 * Ci), sp_out=z(F, Ci, dtype=f32),
                fr_ln=z(F, Ci), kv_t=z(F, 2 * Ci), top_in=z(b, Ci, dtype=f32), top_lnq=z(b, Ci), q_t=z(b, Ci), o_t=z(b, Ci),
                top_mid=z(b, Ci, dtype=f32), top_lno=z(b, Ci), z_mt=z(b, 4 * Ci), h_mt=z(b, 4 * Ci), top_out=z(b, Ci, dtype=f32)))
        # ---- gradient scratch (shared by all layers)
        self.GA, self.GB = z(Mv, Ci, dtype=f32), z(Mv, Ci, dtype=f32)       # grad of res_i / of upd_i (ping-pong)
        self.GR_a, self.GU_a = z(Mv, Ci), z(Mv, Ci)
        self.GX, self.GX_a = z(Mt, Ct, dtype=f32), z(Mt, Ct)
        self.GZ2, self.GZ2_a = z(Mt, Ct, dtype=f32), z(Mt, Ct)
        self.g_hf, self.g_zf, self.g_ht, self.g_zt, self.g_tf1 = z(Mv, Ih), z(Mv, Ih), z(Mv, Cm), z(Mv, Cm), z(Mv, Cm)
        self.g_a1, self.g_a2 = z(Mv, Ci), z(Mv, Ci)
        self.d_u = z(F * P, Ct)
        self.g_y1, self.g_z1, self.g_xln = z(Mt, Ch), z(Mt, Ch), z(Mt, Ct)
        # head gradients
        self.G_top, self.G_top_a, self.G_sp, self.G_sp_a, self.D_fr = z(b, Ci, dtype=f32), z(b, Ci), z(F, Ci, dtype=f32), z(F, Ci), z(F, Ci, dtype=f32)
        self.g_b4, self.g_b4z, self.g_b1, self.g_bq = z(b, 4 * Ci), z(b, 4 * Ci), z(b, Ci), z(b, Ci)
        self.g_f4, self.g_f4z, self.g_f1, self.g_fq = z(F, 4 * Ci), z(F, 4 * Ci), z(F, Ci), z(F, Ci)
        self.g_kvt, self.g_kvs, self.g_lncur = z(F, 2 * Ci), z(Mv, 2 * Ci), z(Mv, Ci)
        self.d_zln = z(b, Ci)

    # ------------------------------------------------------------------------------------------
    # small planning helpers (forward list = self.calls, backward list = self.bwd)
    def _fwd_lin(self, x, m: Mat, bias, out, *, res=None, out2=None, tap=0, name="linear", **kw):
        """out = x @ W^T + bias (+ res); W = forward operand of ``m``"""
        self._gemm(x, m.f[tap], m.n, m.k, ldb=m.kp, bias=None if bias is None else bias.p, res=res, ld_res=m.n, out=out,
                   ld_out=kw.pop("ld_out", m.n), out2=out2, ld_out2=m.n, name=name, **kw)

    def _bwd_lin(self, dy_lp, x_lp, m: Mat, bias, dy_f32=None, *, bias2=None, dx=None, dx_res=None, dx2=None, ld_x=None, ld_dy=None, name="linear"):
        """weight / bias gradients of a plain linear and (optionally) dx = dy @ W (+ dx_res), rows = dy rows"""
        B = self.bwd.append
        rows = dy_lp.shape[0]
        B(ops.wgrad(x_lp, dy_lp, m.g[0], m.n, m.k, a_dim=(m.k, rows, 1, 1), a_stride=(1, ld_x or x_lp.stride(0), 0, 0), rows_per_group=rows,
                    ld_dy=ld_dy or dy_lp.stride(0), ld_dw=m.k, impl=self.gemm_impl, name="bwd." + name + ".wgrad"))
        if bias is not None:
            # the low-precision copy is read (half the bytes of the fp32 gradient; the sum itself is fp32)
            B(ops.colsum(dy_lp, bias.g, m.n, ld=ld_dy or dy_lp.stride(0), rows_per_group=rows, out2=None if bias2 is None else bias2.g,
                         name="bwd." + name + ".bgrad"))
        if dx is not None or dx2 is not None:
            if dx is None:
                dx, dx2 = dx2, None
            self._bgemm(dy_lp, m.t[0], m.k, m.n, a_dim=(m.n, rows, 1, 1), a_stride=(1, ld_dy or dy_lp.stride(0), 0, 0), rows_per_group=rows,
                        ldb=m.np_, res=dx_res, ld_res=m.k, out=dx, ld_out=m.k, out2=dx2, ld_out2=m.k, name="bwd." + name + ".dgrad")

    def _bgemm(self, *args, **kw):
        kw.setdefault("impl", self.gemm_impl)
        self.bwd.append(ops.gemm(*args, **kw))

    # ------------------------------------------------------------------------------------------
    def _plan(self):
        a, b, w, pt = self.arch, self.batch, self.w, self.pt
        t, T, N, P, D, g = a.sparse_frames, a.frames, a.tokens, a.patches, a.width, a.grid
        F, Ci, Ct, al, Ch, Cm, Ih = b * t, a.integration_dim, a.temporal_dim, a.alpha, a.temporal_hidden, a.integration_temporal_hidden, a.integration_hidden
        Mv, Mt = F * N, b * T * P
        add = self.calls.append
        self.bwd = []
        B = self.bwd.append
        half = a.t_kernel // 2
        sel = list(a.selected_layers)
        nl = len(sel)
        assert all(x < y for x, y in zip(sel, sel[1:]))
        S = self.saved
        V = pt.vec

        # =========================== forward ===========================
        self.xT = S[0]["xT_in"]                       # the stem writes the first layer's input
        self._plan_inputs_train()
        for l in range(a.layers):
            self._plan_vit_layer(l, S[sel.index(l)]["tap"] if l in sel else None)
            if l == sel[-1]:
                add(ops.mean_rows(self.h, N * D, t, b, D, self.clsmean, name="dist.cls_mean"))
        conv_t_taps = [((k - half) * P, 0, 0) for k in range(a.t_kernel)]
        conv_s_taps = [(j - 1, ii - 1, 0) for ii in range(3) for j in range(3)]
        tconv_taps = [((k - half) * N, 0, 0) for k in range(a.t_kernel)]
        neg = lambda taps: [(-x, -y, -zz) for (x, y, zz) in taps]
        P_ = {}                                       # per-layer parameter handles
        for i in range(nl):
            s = S[i]
            tn, it = "dist_net.temporal_nets.%d." % i, "dist_net.integration_nets.%d." % i
            t2i, i2t = "dist_net.temporal2integration_nets.%d." % i, "dist_net.integration2temporal_nets.%d." % i
            p = P_[i] = dict(
                tn_ln=(V(tn + "ln.weight"), V(tn + "ln.bias")),
                w1=pt.mat(tn + "temporal_net.c_fc1.weight"), b1=V(tn + "temporal_net.c_fc1.bias"),
                w2=pt.mat(tn + "temporal_net.c_fc2.weight"), b2=V(tn + "temporal_net.c_fc2.bias"),
                win=pt.mat("dist_net.input_linears.%d.weight" % i), bin=V("dist_net.input_linears.%d.bias" % i),
                wi2t=pt.mat(i2t + "linear_fuse.weight"), bi2t=V(i2t + "linear_fuse.bias"),
                wt2i=pt.mat(t2i + "linear_fuse.weight"), bt2i=V(t2i + "linear_fuse.bias"), cls=V(t2i + "cls_token", (t, Ci)),
                ln=(V(it + "ln.weight"), V(it + "ln.bias")), ln_t=(V(it + "ln_temporal.weight"), V(it + "ln_temporal.bias")),
                wfc=pt.mat(it + "ffn.c_fc.weight"), bfc=V(it + "ffn.c_fc.bias"),
                wpr=pt.mat(it + "ffn.c_proj.weight"), bpr=V(it + "ffn.c_proj.bias"),
                wtf1=pt.mat(it + "temporal_ffn.c_fc1.weight"), btf1=V(it + "temporal_ffn.c_fc1.bias"),
                wtf2=pt.mat(it + "temporal_ffn.c_fc2.weight"), btf2=V(it + "temporal_ffn.c_fc2.bias"),
                wtpr=pt.mat(it + "temporal_ffn.c_proj.weight"), btpr=V(it + "temporal_ffn.c_proj.bias"))
            last = i == nl - 1
            # ---- TemporalNet (dist.py:48-65): pre-activations kept
            self._ln(s["xT_in"], (p["tn_ln"][0].p, p["tn_ln"][1].p), s["xln"], name="dist.tn.ln")
            self._gemm(s["xln"], p["w1"].f, Ch, Ct, a_dim=(Ct, T * P, b, 1), a_stride=(1, Ct, T * P * Ct, Mt * Ct), taps=conv_t_taps,
                       b_tap_stride=Ch * p["w1"].kp, ldb=p["w1"].kp, groups=b, rows_per_group=T * P, bias=p["b1"].p, out=s["z1"], ld_out=Ch,
                       name="dist.tn.conv_t")
            add(ops.quickgelu(s["z1"], None, s["y1"], name="dist.tn.act1"))
            self._gemm(s["y1"], p["w2"].f, Ct, Ch, a_dim=(Ch, g, g, b * T), a_stride=(1, Ch, g * Ch, P * Ch), img_w=g, taps=conv_s_taps,
                       b_tap_stride=Ct * p["w2"].kp, ldb=p["w2"].kp, groups=b * T, rows_per_group=P, bias=p["b2"].p, res=s["xT_in"], ld_res=Ct,
                       res_gstride=P, out=s["z2"], ld_out=Ct, name="dist.tn.conv_s")
            add(ops.quickgelu(s["z2"], s["xT_out"], s["xT_a"], name="dist.tn.act2"))
            # ---- input linear (+ previous integration output) (dist.py:229); mid lives in the layer's `upd` buffer
            self._fwd_lin(s["tap"], p["win"], p["bin"], s["upd"], res=self.res if i > 0 else None, out2=s["mid_a"], name="dist.input_linear")
            # ---- temporal -> integration on the pre-fusion stream (dist.py:68-86,232)
            self._gemm(s["xT_a"], p["wt2i"].f, Ci, Ct, a_dim=(Ct, P, al, F), a_stride=(1, Ct, P * Ct, al * P * Ct), group_dim=3,
                       taps=[(0, k, 0) for k in range(al)], b_tap_stride=Ci * p["wt2i"].kp, ldb=p["wt2i"].kp, groups=F, rows_per_group=P,
                       bias=p["bt2i"].p, res=s["upd"], ld_res=Ci, res_gstride=N, res_roff=1, out=s["upd"], ld_out=Ci, out_gstride=N, out_roff=1,
                       name="dist.t2i")
            add(ops.rows_bcast(s["upd"], N * Ci, F, Ci, p["cls"].p, t, True, name="dist.t2i.cls"))
            # ---- integration -> temporal (dist.py:90-105,231); dead for the last layer (nothing reads its output)
            if not last:
                self._gemm(s["mid_a"], p["wi2t"].f, Ct, Ci, a_dim=(Ci, N, F, 1), a_stride=(1, Ci, N * Ci, Mv * Ci), taps=[(1, 0, 0)], groups=F,
                           rows_per_group=P, ldb=p["wi2t"].kp, bias=p["bi2t"].p, res=s["xT_out"], ld_res=Ct, res_gstride=al * P, res_rep_stride=P,
                           out=S[i + 1]["xT_in"], ld_out=Ct, out_gstride=al * P, out_rep=al, out_rep_stride=P, name="dist.i2t")
            # ---- IntegrationNetwork (dist.py:16-45)
            add(ops.layernorm(s["upd"], p["ln"][0].p, p["ln"][1].p, s["a1"], g2=p["ln_t"][0].p, b2=p["ln_t"][1].p, y2=s["a2"], name="dist.int.ln"))
            self._fwd_lin(s["a1"], p["wfc"], p["bfc"], s["zf"], name="dist.int.ffn_fc")
            add(ops.quickgelu(s["zf"], None, s["hf"], name="dist.int.act_f"))
            self._fwd_lin(s["a2"], p["wtf1"], p["btf1"], s["tf1"], name="dist.int.t_fc1")
            self._gemm(s["tf1"], p["wtf2"].f, Cm, Cm, a_dim=(Cm, t * N, b, 1), a_stride=(1, Cm, t * N * Cm, Mv * Cm), taps=tconv_taps,
                       b_tap_stride=Cm * p["wtf2"].kp, ldb=p["wtf2"].kp, groups=b, rows_per_group=t * N, bias=p["btf2"].p, out=s["zt"], ld_out=Cm,
                       name="dist.int.t_conv")
            add(ops.quickgelu(s["zt"], None, s["ht"], name="dist.int.act_t"))
            self._fwd_lin(s["hf"], p["wpr"], p["bpr"], self.res, name="dist.int.ffn_proj")
            self._fwd_lin(s["ht"], p["wtpr"], p["btpr"], self.res, res=self.res, name="dist.int.t_proj")
        self._plan_head_train()

        # =========================== backward ===========================
        self._plan_head_bwd()                           # leaves d loss / d cur in self.GA
        GR, GU = self.GA, self.GB
        GR_a, GU_a = self.GR_a, self.GU_a
        for i in reversed(range(nl)):
            s, p = S[i], P_[i]
            last = i == nl - 1
            nm = "dist%d." % i
            if last:
                B(ops.cast(GR, GR_a, name="bwd." + nm + "cast_gres"))      # later layers inherit the copy written with GU
            # ---- projections of the IntegrationNetwork: res = hf W_p^T + b_p + ht W_tp^T + b_tp (both biases see the same sum)
            self._bwd_lin(GR_a, s["hf"], p["wpr"], p["bpr"], bias2=p["btpr"], dx2=self.g_hf, name=nm + "int.ffn_proj")
            self._bwd_lin(GR_a, s["ht"], p["wtpr"], None, dx2=self.g_ht, name=nm + "int.t_proj")
            B(ops.quickgelu_bwd(self.g_hf, s["zf"], None, self.g_zf, name="bwd." + nm + "int.act_f"))
            B(ops.quickgelu_bwd(self.g_ht, s["zt"], None, self.g_zt, name="bwd." + nm + "int.act_t"))
            self._bwd_lin(self.g_zf, s["a1"], p["wfc"], p["bfc"], dx2=self.g_a1, name=nm + "int.ffn_fc")
            # (kt,1,1) conv along the sparse-frame axis
            B(ops.wgrad(s["tf1"], self.g_zt, p["wtf2"].g, Cm, Cm, a_dim=(Cm, t * N, b, 1), a_stride=(1, Cm, t * N * Cm, Mv * Cm), taps=tconv_taps,
                        groups=b, rows_per_group=t * N, impl=self.gemm_impl, name="bwd." + nm + "int.t_conv.wgrad"))
            B(ops.colsum(self.g_zt, p["btf2"].g, Cm, name="bwd." + nm + "int.t_conv.bgrad"))
            self._bgemm(self.g_zt, p["wtf2"].t, Cm, Cm, a_dim=(Cm, t * N, b, 1), a_stride=(1, Cm, t * N * Cm, Mv * Cm), taps=neg(tconv_taps),
                        b_tap_stride=Cm * p["wtf2"].np_, ldb=p["wtf2"].np_, groups=b, rows_per_group=t * N, out=self.g_tf1, ld_out=Cm,
                        name="bwd." + nm + "int.t_conv.dgrad")
            self._bwd_lin(self.g_tf1, s["a2"], p["wtf1"], p["btf1"], dx2=self.g_a2, name=nm + "int.t_fc1")
            # ---- LayerNorm pair -> gradient of upd (plus the direct path cur = res + upd of the last layer)
            B(ops.layernorm_bwd(s["upd"], p["ln"][0].p, self.g_a1, g2=p["ln_t"][0].p, dy2=self.g_a2, add=GR if last else None, dx=GU,
                                dx_lp=GU_a, dg1=p["ln"][0].g, db1=p["ln"][1].g, dg2=p["ln_t"][0].g, db2=p["ln_t"][1].g,
                                name="bwd." + nm + "int.ln"))
            B(ops.colsum(GU, p["cls"].g, Ci, groups=F, rows_per_group=1, gstride=N, roff=0, period=t, name="bwd." + nm + "t2i.cls"))
            # ---- integration -> temporal, part 1: collect the alpha dense frames of every sparse frame (needs the untouched GX)
            if not last:
                B(ops.group_sum(self.GX, self.d_u, F, al, P * Ct, name="bwd." + nm + "i2t.frame_sum"))
            # ---- temporal -> integration: d_v = GU[:, 1:]
            B(ops.wgrad(s["xT_a"], GU_a, p["wt2i"].g, Ci, Ct, a_dim=(Ct, P, al, F), a_stride=(1, Ct, P * Ct, al * P * Ct), group_dim=3,
                        taps=[(0, k, 0) for k in range(al)], groups=F, rows_per_group=P, dy_gstride=N, dy_roff=1, impl=self.gemm_impl,
                        name="bwd." + nm + "t2i.wgrad"))
            B(ops.colsum(GU_a, p["bt2i"].g, Ci, groups=F, rows_per_group=P, gstride=N, roff=1, name="bwd." + nm + "t2i.bgrad"))
            for k in range(al):
                self._bgemm(GU_a, p["wt2i"].t[k], Ct, Ci, a_dim=(Ci, N, F, 1), a_stride=(1, Ci, N * Ci, Mv * Ci), taps=[(1, 0, 0)], groups=F,
                            rows_per_group=P, ldb=p["wt2i"].np_, res=None if last else self.GX, ld_res=Ct, res_gstride=al * P, res_roff=k * P,
                            out=self.GX, ld_out=Ct, out_gstride=al * P, out_roff=k * P, name="bwd." + nm + "t2i.dgrad%d" % k)
            # ---- integration -> temporal, part 2
            if not last:
                B(ops.wgrad(s["mid_a"], self.d_u, p["wi2t"].g[0], Ct, Ci, a_dim=(Ci, N, F, 1), a_stride=(1, Ci, N * Ci, Mv * Ci), taps=[(1, 0, 0)],
                            groups=F, rows_per_group=P, impl=self.gemm_impl, name="bwd." + nm + "i2t.wgrad"))
                B(ops.colsum(self.d_u, p["bi2t"].g, Ct, name="bwd." + nm + "i2t.bgrad"))
                self._bgemm(self.d_u, p["wi2t"].t[0], Ci, Ct, a_dim=(Ct, P, F, 1), a_stride=(1, Ct, P * Ct, F * P * Ct), groups=F, rows_per_group=P,
                            ldb=p["wi2t"].np_, res=GU, ld_res=Ci, res_gstride=N, res_roff=1, out=GU, ld_out=Ci, out_gstride=N, out_roff=1,
                            out2=GU_a, ld_out2=Ci, name="bwd." + nm + "i2t.dgrad")
            # ---- input linear: only parameters (the taps come from the frozen ViT)
            self._bwd_lin(GU_a, s["tap"], p["win"], p["bin"], name=nm + "input_linear")
            # ---- TemporalNet: GX = gradient of the post-activation stream
            B(ops.quickgelu_bwd(self.GX, s["z2"], self.GZ2, self.GZ2_a, name="bwd." + nm + "tn.act2"))
            B(ops.wgrad(s["y1"], self.GZ2_a, p["w2"].g, Ct, Ch, a_dim=(Ch, g, g, b * T), a_stride=(1, Ch, g * Ch, P * Ch), img_w=g, taps=conv_s_taps,
                        groups=b * T, rows_per_group=P, impl=self.gemm_impl, name="bwd." + nm + "tn.conv_s.wgrad"))
            B(ops.colsum(self.GZ2_a, p["b2"].g, Ct, name="bwd." + nm + "tn.conv_s.bgrad"))
            self._bgemm(self.GZ2_a, p["w2"].t, Ch, Ct, a_dim=(Ct, g, g, b * T), a_stride=(1, Ct, g * Ct, P * Ct), img_w=g, taps=neg(conv_s_taps),
                        b_tap_stride=Ch * p["w2"].np_, ldb=p["w2"].np_, groups=b * T, rows_per_group=P, out=self.g_y1, ld_out=Ch,
                        name="bwd." + nm + "tn.conv_s.dgrad")
            B(ops.quickgelu_bwd(self.g_y1, s["z1"], None, self.g_z1, name="bwd." + nm + "tn.act1"))
            B(ops.wgrad(s["xln"], self.g_z1, p["w1"].g, Ch, Ct, a_dim=(Ct, T * P, b, 1), a_stride=(1, Ct, T * P * Ct, Mt * Ct), taps=conv_t_taps,
                        groups=b, rows_per_group=T * P, impl=self.gemm_impl, name="bwd." + nm + "tn.conv_t.wgrad"))
            B(ops.colsum(self.g_z1, p["b1"].g, Ch, name="bwd." + nm + "tn.conv_t.bgrad"))
            self._bgemm(self.g_z1, p["w1"].t, Ct, Ch, a_dim=(Ch, T * P, b, 1), a_stride=(1, Ch, T * P * Ch, Mt * Ch), taps=neg(conv_t_taps),
                        b_tap_stride=Ct * p["w1"].np_, ldb=p["w1"].np_, groups=b, rows_per_group=T * P, out=self.g_xln, ld_out=Ct,
                        name="bwd." + nm + "tn.conv_t.dgrad")
            B(ops.layernorm_bwd(s["xT_in"], p["tn_ln"][0].p, self.g_xln, add=self.GZ2, dx=self.GX, dx_lp=self.GX_a if i == 0 else None,
                                dg1=p["tn_ln"][0].g, db1=p["tn_ln"][1].g, name="bwd." + nm + "tn.ln"))
            GR, GU = GU, GR                               # the gradient of this layer's mid is the gradient of res_{i-1}
            GR_a, GU_a = GU_a, GR_a
        # ---- temporal stem (parameters only)
        stem = pt.mat("dist_net.temporal_stem.weight")
        a_dim, a_stride, taps = self._stem_operand()
        ks = 3 * a.s_patch * a.s_patch
        B(ops.wgrad(self.patches_d, self.GX_a, stem.g, Ct, ks, a_dim=a_dim, a_stride=a_stride, taps=taps, groups=b, rows_per_group=T * P,
                    ld_dw=ks, impl=self.gemm_impl, name="bwd.stem.wgrad"))
        B(ops.colsum(self.GX_a, V("dist_net.temporal_stem.bias").g, Ct, name="bwd.stem.bgrad"))

        # =========================== operand refresh ===========================
        self.pack_calls = [m.pack_call(k) for k, m in sorted(pt.mats.items())]

    # ------------------------------------------------------------------------------------------
    def _plan_inputs_train(self):
        """As ``DistEngine._plan_inputs`` with the stem reading its master-derived operand."""
        stem, bias = self.pt.mat("dist_net.temporal_stem.weight"), self.pt.vec("dist_net.temporal_stem.bias")
        assert stem.kp == self.w.kps
        self.w.stem_w, self.w.stem_b = stem.f, bias.p
        self._plan_inputs()

    def _plan_head_train(self):
        a, b, w, pt = self.arch, self.batch, self.w, self.pt
        t, N, Ci, E = a.sparse_frames, a.tokens, a.integration_dim, a.embed_dim
        F, Mv, H = b * t, b * t * N, a.integration_heads
        add = self.calls.append
        V = pt.vec
        last = self.saved[-1]
        self.HP = []
        sp_prev, top_prev = None, None
        for j in range(a.ada_layers):
            pre = "dist_net.adapooling_nets.%d." % j
            A = self.saved_ada[j]
            hp = {"pos": V(pre + "positional_embedding", (t, Ci))}
            for tag, which, ln_out, mlp in (("sp", "spatial_transformer", "ln_out_spat_cls_token", "output_map_spatial_cls_token"),
                                            ("tp", "temporal_transformer", "ln_out_temp_cls_token", "output_map_cls_token")):
                inw, inb = pt.mat(pre + which + ".attn.in_proj_weight"), V(pre + which + ".attn.in_proj_bias")
                hp[tag] = dict(ln=(V(pre + which + ".ln_1.weight"), V(pre + which + ".ln_1.bias")),
                               q=inw.rows(0, Ci), qb=inb.part(0, Ci), kv=inw.rows(Ci, 3 * Ci), kvb=inb.part(Ci, 3 * Ci),
                               o=pt.mat(pre + which + ".attn.out_proj.weight"), ob=V(pre + which + ".attn.out_proj.bias"),
                               ln_out=(V(pre + ln_out + ".weight"), V(pre + ln_out + ".bias")),
                               fc=pt.mat(pre + mlp + ".c_fc.weight"), fcb=V(pre + mlp + ".c_fc.bias"),
                               pr=pt.mat(pre + mlp + ".c_proj.weight"), prb=V(pre + mlp + ".c_proj.bias"))
            self.HP.append(hp)
            s, tp = hp["sp"], hp["tp"]
            lnp = lambda pair: (pair[0].p, pair[1].p)
            # ---- spatial pooling (dist.py:139-150)
            if j == 0:
                add(ops.rows_bcast(A["sp_in"], Ci, F, Ci, V("dist_net.aggregated_spatial_cls_token", (1, Ci)).p, 1, False, name="ada.init_sp"))
                add(ops.rows_bcast(A["top_in"], Ci, b, Ci, V("dist_net.aggregated_cls_token", (1, Ci)).p, 1, False, name="ada.init_top"))
            else:
                A["sp_in"], A["top_in"] = sp_prev, top_prev          # the previous layer's outputs (aliases)
            self._ln(self.res, lnp(s["ln"]), A["ln_cur"], in2=last["upd"], in2_period=Mv, name="ada.sp.ln_kv")
            self._fwd_lin(A["ln_cur"], s["kv"], s["kvb"], A["kv_s"], name="ada.sp.kv")
            self._ln(A["sp_in"], lnp(s["ln"]), A["sp_lnq"], name="ada.sp.ln_q")
            self._fwd_lin(A["sp_lnq"], s["q"], s["qb"], A["q_s"], name="ada.sp.q")
            add(ops.cross_attention(A["q_s"], A["kv_s"], A["o_s"], F, N, H, name="ada.sp.attn"))
            self._fwd_lin(A["o_s"], s["o"], s["ob"], A["sp_mid"], res=A["sp_in"], name="ada.sp.out_proj")
            self._ln(A["sp_mid"], lnp(s["ln_out"]), A["sp_lno"], name="ada.sp.ln_out")
            self._fwd_lin(A["sp_lno"], s["fc"], s["fcb"], A["z_ms"], name="ada.sp.fc")
            add(ops.quickgelu(A["z_ms"], None, A["h_ms"], name="ada.sp.act"))
            self._fwd_lin(A["h_ms"], s["pr"], s["prb"], A["sp_out"], res=A["sp_mid"], name="ada.sp.proj")
            # ---- temporal pooling (dist.py:152-162)
            self._ln(A["sp_out"], lnp(tp["ln"]), A["fr_ln"], in2=hp["pos"].p, in2_period=t, name="ada.tp.ln_kv")
            self._fwd_lin(A["fr_ln"], tp["kv"], tp["kvb"], A["kv_t"], name="ada.tp.kv")
            self._ln(A["top_in"], lnp(tp["ln"]), A["top_lnq"], name="ada.tp.ln_q")
            self._fwd_lin(A["top_lnq"], tp["q"], tp["qb"], A["q_t"], name="ada.tp.q")
            add(ops.cross_attention(A["q_t"], A["kv_t"], A["o_t"], b, t, H, name="ada.tp.attn"))
            self._fwd_lin(A["o_t"], tp["o"], tp["ob"], A["top_mid"], res=A["top_in"], name="ada.tp.out_proj")
            self._ln(A["top_mid"], lnp(tp["ln_out"]), A["top_lno"], name="ada.tp.ln_out")
            self._fwd_lin(A["top_lno"], tp["fc"], tp["fcb"], A["z_mt"], name="ada.tp.fc")
            add(ops.quickgelu(A["z_mt"], None, A["h_mt"], name="ada.tp.act"))
            self._fwd_lin(A["h_mt"], tp["pr"], tp["prb"], A["top_out"], res=A["top_mid"], name="ada.tp.proj")
            sp_prev, top_prev = A["sp_out"], A["top_out"]
        # ---- tail (dist.py:242-246) and train-mode head + loss
        self.pcls, self.pcls_b = pt.mat("dist_net.proj_spatial_cls_token.weight"), V("dist_net.proj_spatial_cls_token.bias")
        self.ln_post = (V("dist_net.ln_post.weight"), V("dist_net.ln_post.bias"))
        self.proj = pt.mat("dist_net.proj")                        # [Ci, E]: emb = z_ln @ proj  -> B operand = proj^T = .t
        self._fwd_lin(self.clsmean, self.pcls, self.pcls_b, self.zbuf, res=top_prev, name="tail.proj_spatial_cls")
        self._ln(self.zbuf, (self.ln_post[0].p, self.ln_post[1].p), self.z_ln, name="tail.ln_post")
        self._gemm(self.z_ln, self.proj.t[0], E, Ci, ldb=self.proj.np_, out=self.emb, ld_out=E, name="tail.proj")
        add(ops.softce_head(self.emb, self.text_n, w.logit_scale, self.target, b, E, self.text_n.shape[0], self.logits, self.loss, self.d_emb,
                            name="head.softce"))
        self.head_call = None

    def _plan_head_bwd(self):
        a, b, pt = self.arch, self.batch, self.pt
        t, N, Ci, E = a.sparse_frames, a.tokens, a.integration_dim, a.embed_dim
        F, Mv, H = b * t, b * t * N, a.integration_heads
        B = self.bwd.append
        V = pt.vec
        last = self.saved[-1]
        # ---- tail: emb = z_ln @ proj
        B(ops.cast(self.d_emb, self.d_emb_a, name="bwd.tail.cast"))
        B(ops.wgrad(self.d_emb_a, self.z_ln, self.proj.g[0], Ci, E, a_dim=(E, b, 1, 1), a_stride=(1, E, 0, 0), rows_per_group=b, ld_dy=Ci,
                    ld_dw=E, impl=self.gemm_impl, name="bwd.tail.proj.wgrad"))
        self._bgemm(self.d_emb_a, self.proj.f[0], Ci, E, ldb=self.proj.kp, out=self.d_zln, ld_out=Ci, name="bwd.tail.proj.dgrad")
        B(ops.layernorm_bwd(self.zbuf, self.ln_post[0].p, self.d_zln, dx=self.G_top, dx_lp=self.G_top_a, dg1=self.ln_post[0].g,
                            db1=self.ln_post[1].g, name="bwd.tail.ln_post"))
        self._bwd_lin(self.G_top_a, self.clsmean, self.pcls, self.pcls_b, self.G_top, name="tail.proj_spatial_cls")
        for j in reversed(range(a.ada_layers)):
            A, hp = self.saved_ada[j], self.HP[j]
            s, tp = hp["sp"], hp["tp"]
            nm = "ada%d." % j
            top_layer = j == a.ada_layers - 1
            # ---- temporal pooling, reverse order
            if not top_layer:
                B(ops.cast(self.G_top, self.G_top_a, name="bwd." + nm + "tp.cast"))
            self._bwd_lin(self.G_top_a, A["h_mt"], tp["pr"], tp["prb"], self.G_top, dx2=self.g_b4, name=nm + "tp.proj")
            B(ops.quickgelu_bwd(self.g_b4, A["z_mt"], None, self.g_b4z, name="bwd." + nm + "tp.act"))
            self._bwd_lin(self.g_b4z, A["top_lno"], tp["fc"], tp["fcb"], dx2=self.g_b1, name=nm + "tp.fc")
            B(ops.layernorm_bwd(A["top_mid"], tp["ln_out"][0].p, self.g_b1, add=self.G_top, dx=self.G_top, dx_lp=self.G_top_a,
                                dg1=tp["ln_out"][0].g, db1=tp["ln_out"][1].g, name="bwd." + nm + "tp.ln_out"))
            self._bwd_lin(self.G_top_a, A["o_t"], tp["o"], tp["ob"], self.G_top, dx2=self.g_b1, name=nm + "tp.out_proj")
            B(ops.cross_attention_bwd(A["q_t"], A["kv_t"], self.g_b1, self.g_bq, self.g_kvt, b, t, H, name="bwd." + nm + "tp.attn"))
            self._bwd_lin(self.g_bq, A["top_lnq"], tp["q"], tp["qb"], dx2=self.g_b1, name=nm + "tp.q")
            B(ops.layernorm_bwd(A["top_in"], tp["ln"][0].p, self.g_b1, add=self.G_top, dx=self.G_top, dg1=tp["ln"][0].g, db1=tp["ln"][1].g,
                                name="bwd." + nm + "tp.ln_q"))
            self._bwd_lin(self.g_kvt, A["fr_ln"], tp["kv"], tp["kvb"], dx2=self.g_f1, name=nm + "tp.kv")
            # frame tokens fr = sp_out + pos: the LayerNorm gradient feeds pos (summed over clips) and sp_out
            if top_layer:
                B(ops.layernorm_bwd(A["sp_out"], tp["ln"][0].p, self.g_f1, in2=hp["pos"].p, in2_period=t, dx=self.G_sp, dg1=tp["ln"][0].g,
                                    db1=tp["ln"][1].g, name="bwd." + nm + "tp.ln_kv"))
                B(ops.colsum(self.G_sp, hp["pos"].g, Ci, groups=F, rows_per_group=1, gstride=1, roff=0, period=t, name="bwd." + nm + "pos"))
            else:
                B(ops.layernorm_bwd(A["sp_out"], tp["ln"][0].p, self.g_f1, in2=hp["pos"].p, in2_period=t, dx=self.D_fr, dg1=tp["ln"][0].g,
                                    db1=tp["ln"][1].g, name="bwd." + nm + "tp.ln_kv"))
                B(ops.colsum(self.D_fr, hp["pos"].g, Ci, groups=F, rows_per_group=1, gstride=1, roff=0, period=t, name="bwd." + nm + "pos"))
                B(ops.layernorm_bwd(A["sp_out"], tp["ln"][0].p, self.g_f1, in2=hp["pos"].p, in2_period=t, dx=self.G_sp, accumulate=True,
                                    name="bwd." + nm + "tp.ln_kv.acc"))
            # ---- spatial pooling, reverse order
            B(ops.cast(self.G_sp, self.G_sp_a, name="bwd." + nm + "sp.cast"))
            self._bwd_lin(self.G_sp_a, A["h_ms"], s["pr"], s["prb"], self.G_sp, dx2=self.g_f4, name=nm + "sp.proj")
            B(ops.quickgelu_bwd(self.g_f4, A["z_ms"], None, self.g_f4z, name="bwd." + nm + "sp.act"))
            self._bwd_lin(self.g_f4z, A["sp_lno"], s["fc"], s["fcb"], dx2=self.g_f1, name=nm + "sp.fc")
            B(ops.layernorm_bwd(A["sp_mid"], s["ln_out"][0].p, self.g_f1, add=self.G_sp, dx=self.G_sp, dx_lp=self.G_sp_a, dg1=s["ln_out"][0].g,
                                db1=s["ln_out"][1].g, name="bwd." + nm + "sp.ln_out"))
            self._bwd_lin(self.G_sp_a, A["o_s"], s["o"], s["ob"], self.G_sp, dx2=self.g_f1, name=nm + "sp.out_proj")
            B(ops.cross_attention_bwd(A["q_s"], A["kv_s"], self.g_f1, self.g_fq, self.g_kvs, F, N, H, name="bwd." + nm + "sp.attn"))
            self._bwd_lin(self.g_fq, A["sp_lnq"], s["q"], s["qb"], dx2=self.g_f1, name=nm + "sp.q")
            B(ops.layernorm_bwd(A["sp_in"], s["ln"][0].p, self.g_f1, add=self.G_sp, dx=self.G_sp, dg1=s["ln"][0].g, db1=s["ln"][1].g,
                                name="bwd." + nm + "sp.ln_q"))
            self._bwd_lin(self.g_kvs, A["ln_cur"], s["kv"], s["kvb"], dx2=self.g_lncur, name=nm + "sp.kv")
            B(ops.layernorm_bwd(self.res, s["ln"][0].p, self.g_lncur, in2=last["upd"], in2_period=Mv, dx=self.GA, accumulate=not top_layer,
                                dg1=s["ln"][0].g, db1=s["ln"][1].g, name="bwd." + nm + "sp.ln_kv"))
        B(ops.colsum(self.G_sp, V("dist_net.aggregated_spatial_cls_token", (1, Ci)).g, Ci, name="bwd.agg_sp"))
        B(ops.colsum(self.G_top, V("dist_net.aggregated_cls_token", (1, Ci)).g, Ci, name="bwd.agg_cls"))

    # ------------------------------------------------------------------------------------------
    def _launch(self, calls, stream=None):
        s = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        for c in calls:
            c.launch(s)

    def repack(self, stream=None):
        """Refresh the bf16 / fp32 GEMM operands from the masters (after construction and after every update)."""
        self._launch(self.pack_calls, stream)

    def forward_backward(self, video, target):
        """loss and all ``dist_net`` gradients for one batch; returns the loss (1-element device tensor)."""
        self.video.copy_(video, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        self.pt.g.zero_()
        self.loss.zero_()
        if self.fwd_graph is not None:
            self.fwd_graph.replay()
        else:
            self._launch(self.calls)
            self._launch(self.bwd)
        return self.loss

    def capture(self):
        """CUDA graph of forward + backward (static buffers); the update stays outside (its scalars change per step)."""
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self._launch(self.calls, side)
            self._launch(self.bwd, side)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._launch(self.calls)
            self._launch(self.bwd)
        self.fwd_graph = graph
        return graph

    def optimizer_step(self, lr):
        """Gradient averaging over the process group, AdamW on the two decay classes, operand refresh."""
        pt = self.pt
        scale = reduce_gradients(pt, self.process_group)
        self.step_count += 1
        st = torch.cuda.current_stream(self.device).cuda_stream
        b1, b2 = self.betas
        if pt.n_decay > 0:
            ops.adamw(pt.p, pt.g, pt.m, pt.v, pt.n_decay, lr, b1, b2, self.eps, self.weight_decay, self.step_count, scale, st)
        if pt.n_used > pt.n_decay:
            o = pt.n_decay
            ops.adamw(pt.p[o:], pt.g[o:], pt.m[o:], pt.v[o:], pt.n_used - o, lr, b1, b2, self.eps, 0.0, self.step_count, scale, st)
        self.repack()

    def train_step(self, video, target, lr):
        loss = self.forward_backward(video, target)
        self.optimizer_step(lr)
        return loss

    # ------------------------------------------------------------------------------------------
    def gradients(self):
        """{reference name: gradient in the reference layout} (CPU fp32), without the tensors that get none."""
        g = self.pt.export(self.pt.g)
        return {k: v for k, v in g.items() if k not in self.pt.unused}

    def state_dict(self):
        return self.pt.export()

    def flops(self):
        return sum(c.flops for c in self.calls) + sum(c.flops for c in self.bwd)
