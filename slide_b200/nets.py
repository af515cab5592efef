"""Lowering of the reference's networks to slide programs (records of include/slide_program.h).

Everything here is driven by the reference's own JSON hyper-parameters and state-dict key names
(checkpoint ABI), so a reference checkpoint lowers unchanged:

  lower_cloud_net        PointNet2CloudCondition.forward without a condition cloud
                         (pointnet2/models/pointnet2_with_pcld_condition.py:286-489): the position / feature
                         DDPM denoisers and the decoder levels' feature extractors
  _lower_mlp             Mlp_plus_t_emb (pointnet2_ops_lib/pointnet2_ops/pointnet2_modules.py:72-176)
  _lower_attention       AttentionModule (pointnet2_ops_lib/pointnet2_ops/attention.py:35-96)
  _lower_sa / _lower_fp  PointnetSAModule / PointnetKnnFPModule forward (pointnet2_modules.py:222-292, 771-873)
  _lower_feature_map     FeatureMapModule forward (pointnet2_modules.py:640-663)
  lower_decode           PointAutoencoder.decode (pointnet2/models/autoencoder.py:42-45,
                         point_upsample_decoder.py:106-190, keypoint_decoder.py:25-36)

Algebra used (exact in real arithmetic, fp32 re-association only):
  * GroupNorm statistics are produced by the GEMM that writes a tensor and applied by its consumers (XF).
  * AttentionModule's conv over cat[q_i broadcast over K, k_ij] is split into a per-point GEMM on q (np rows)
    plus a per-pair GEMM on k (np*K rows): W[q;k] = Wq q + Wk k.  GroupNorm over the concatenation still
    uses the joint statistics (the q rows enter them with weight K).
"""
import os

import numpy as np

from .program import XF, NO_XF, Builder  # noqa: F401


FACTOR_MIN_WORK = 250000  # see _Grouped.plan

# contractions at least this large get a tensor-core weight copy (the kernel applies further shape tests)
TC_MIN_K, TC_MIN_N = 32, 32


def _np(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float32)


class Params(object):
    """Prefix view over a state dict (values: torch tensors or numpy arrays)."""

    def __init__(self, sd, prefix=""):
        self.sd, self.prefix = sd, prefix

    def __getitem__(self, key):
        return _np(self.sd[self.prefix + key])

    def has(self, key):
        return (self.prefix + key) in self.sd

    def sub(self, name):
        return Params(self.sd, self.prefix + name + ".")


def _gn_dims(C):
    """MyGroupNorm(min(32, C), C): (normalised channels, channels per group)."""
    G = min(32, C)
    nnorm = C - C % G
    return nnorm, nnorm // G


def _conv(b, P, cols=None):
    """Conv2d/Conv1d 1x1 (or Linear) weight -> (W tuple for Builder.gemm, bias offset)."""
    w = P["weight"]
    w = w.reshape(w.shape[0], -1)
    if cols is not None:
        w = w[:, cols[0]:cols[1]]
    off, ldw = b.weight_matrix(w)
    bias = b.weight(P["bias"]) if P.has("bias") else -1
    if w.shape[1] >= TC_MIN_K and w.shape[0] >= TC_MIN_N:
        wp, na = b.weight_matrix_tc(w)  # tensor-core copy (TF32, pre-swizzled tiles)
        return (off, ldw, w.shape[0], w.shape[1], wp, na), bias
    return (off, ldw, w.shape[0], w.shape[1]), bias


class _Ctx(object):
    """Per-network lowering context: timestep-table / condition sources and activation name."""

    def __init__(self, b, cfg, t_src=None, cond_src=None, cond2_src=None, inline=False):
        self.b, self.cfg = b, cfg
        self.t_src = t_src        # Tensor [T, 4*t_dim] (swish'ed timestep embeddings for every t) or None
        self.cond_src = cond_src  # Tensor [B, condition_dim] or None (class embedding, or the global feature)
        self.cond2_src = cond2_src  # Tensor [B, second_condition_dim] or None
        self.inline = inline      # True: condition projections depend on run-time data -> emit them in place
        self.setup = []           # deferred setup GEMMs (emitted into the setup segment)
        self.geom = None          # list: kNN records deferred into a per-chain "geometry" segment (frozen coordinates)
        act = cfg.get("activation", "relu")
        if act != "relu":
            raise NotImplementedError("activation %r (shipped configs use relu)" % act)
        if cfg["bn_first"] or not cfg.get("bn", True) or not cfg["res_connect"]:
            raise NotImplementedError("bn_first / bn=False / res_connect=False are not lowered")

    def knn(self, q, ref, K, idx, d2=None, note=""):
        """Neighbour search of a module.  With frozen coordinates (keypoint-conditional sampling: the xyz columns of x
        never change during a chain) the record is deferred: the caller emits it once per chain instead of every step."""
        if self.geom is None:
            self.b.knn(q, ref, K, idx, d2=d2, note=note)
        else:
            self.geom.append(lambda: self.b.knn(q, ref, K, idx, d2=d2, note=note))

    def project(self, **g):
        """A per-sample / per-timestep projection feeding an additive vector."""
        if self.inline:
            self.b.gemm(g["A"], g["W"], g["out"], bias=g["bias"], note=g["note"])
        else:
            self.setup.append(g)


class _Grouped(object):
    """The grouped input of a set-abstraction / propagation / mapper module and the 1x1 convs that read it.

    Materialised form: a GROUP record builds [f_j | geometry] rows and every conv is a GEMM over the np*K pairs.
    Factored form ("conv before gather", default): a conv over a grouped row is linear in the row's parts, so ONE GEMM
    over the source points produces U = f W_f^T for all convs that read the group (first MLP conv, residual conv,
    attention key conv) and a PAIR record per conv adds the gathered U row, the 3-vector terms and the bias -- the
    grouped tensor is never built and the pair-level GEMMs disappear (16x fewer MACs for those layers at K=16)."""

    def __init__(self, ctx, mode, feats, xyz, ctr, idx, K, name, d2=None, inc_abs=True, inc_ctr=True):
        self.ctx, self.b, self.mode, self.feats, self.xyz, self.ctr, self.idx, self.K = ctx, ctx.b, mode, feats, xyz, ctr, idx, K
        self.d2, self.inc_abs, self.inc_ctr, self.name = d2, inc_abs, inc_ctr, name
        self.B, self.np = ctr.B, ctr.R
        self.R = self.np * K
        self.C = feats.C
        self.Ctot = self.C + (11 if mode == 1 else 3 + 3 * int(inc_abs) + 3 * int(inc_ctr))
        self.factored = None  # decided in plan(): depends on how wide the convs reading the group are
        self.lite = False
        self.G = None
        self.U = None
        self.slices = {}

    def _materialise(self):
        self.G = self.b.tensor(self.name + ".grouped", self.R, self.Ctot, B=self.B)
        self.b.group(self.mode, self.feats, self.C, self.xyz, self.ctr, self.idx, self.K, self.G, d2=self.d2,
                     include_abs=self.inc_abs, include_center=self.inc_ctr, note=self.name + ".group")

    def _split(self, w):
        """(N, Ctot) conv weight -> (W_f, WX, WC, wd, ww) for the factored form."""
        C = self.C
        z = np.zeros((w.shape[0], 3), dtype=np.float32)
        if self.mode == 0:
            rel = w[:, C:C + 3]
            pos = C + 3
            ab = z
            if self.inc_abs:
                ab = w[:, pos:pos + 3]
                pos += 3
            ct = w[:, pos:pos + 3] if self.inc_ctr else z
            return w[:, :C], ab + rel, ct - rel, None, None
        wd, ww = w[:, C], w[:, C + 1]
        ab, rel, ct = w[:, C + 2:C + 5], w[:, C + 5:C + 8], w[:, C + 8:C + 11]
        return w[:, :C], ab + rel, ct - rel, wd, ww

    def plan(self, convs):
        """convs: list of Params of every conv that reads the group.  Chooses the form and emits the GROUP record
        (materialised) or the shared U GEMM (factored).  A/B on B200 (feature / position DDPM, batch 256): factoring pays
        once the pair-level GEMMs it removes are large -- Ctot x (sum of conv widths) above ~2.5e5 -- below that the
        tcgen05 GEMMs over the grouped tensor are cheaper than the extra U GEMM + the gather-bound PAIR kernels (large
        gather sources: decode / encode levels)."""
        ntot = sum(int(Pc["weight"].shape[0]) for Pc in convs)
        mode = os.environ.get("SLIDE_FACTOR_GROUP", "auto")
        if getattr(self.ctx, "factor_group", None):  # resident plans need the factored form (no grouped tensor at all)
            mode = "1"
        # (round 2) with the gather source staged in shared memory (pair_smem_kernel: source = the sample's own <= 64
        # points) PAIR is cheap, and at >= 32768 pair rows factoring wins for every module of both denoisers (B200, batch
        # 256: position step 757 -> 710 us, feature step 1575 -> 1525 us); at batch 32 it loses (425 -> 445 us)
        small_src = self.feats.R <= 64 and self.B * self.R >= 32768
        # (round 2, after the warp kNN / PAIR work) large gather sources -- autoencoder and refinement levels, 256..4096
        # source points -- also gain from factoring every module: encode + decode of 512 clouds 218.0 -> 213.1 ms, the SAP
        # refinement network 10.46 -> 9.77 ms per 32 clouds (its last propagation module otherwise runs a GROUP record
        # and three GEMMs over 1 M materialised pair rows)
        large_src = self.feats.R > 64
        self.factored = (mode == "1") or (mode == "auto" and (self.Ctot * ntot >= FACTOR_MIN_WORK or small_src or large_src))
        if not self.factored:
            self._materialise()
            return
        b = self.b
        parts, off = [], 0
        for Pc in convs:
            w = Pc["weight"]
            w = w.reshape(w.shape[0], -1)
            assert w.shape[1] == self.Ctot, (w.shape, self.Ctot)
            wf, wx, wc, wd, ww = self._split(w)
            self.slices[Pc.prefix] = dict(off=off, N=w.shape[0], wx_value=np.ascontiguousarray(wx),
                                          wx=b.weight(np.ascontiguousarray(wx)),
                                          wc=b.weight(np.ascontiguousarray(wc)),
                                          wd=b.weight(wd) if wd is not None else -1,
                                          ww=b.weight(ww) if ww is not None else -1,
                                          bias=b.weight(Pc["bias"]) if Pc.has("bias") else -1)
            parts.append(wf)
            off += _align4(w.shape[0])
        wcat = np.zeros((off, self.C), dtype=np.float32)
        for Pc, wf in zip(convs, parts):
            o = self.slices[Pc.prefix]["off"]
            wcat[o:o + wf.shape[0]] = wf
        woff, ldw = b.weight_matrix(wcat)
        W = (woff, ldw, off, self.C)
        if self.C >= TC_MIN_K and off >= TC_MIN_N:
            W = W + b.weight_matrix_tc(wcat)
        self.U = b.tensor(self.name + ".U", self.feats.R, off, B=self.B)
        b.gemm(self.feats, W, self.U, note=self.name + ".U")
        # Large gather sources: fold the neighbour-coordinate term x_j wx^T into U as well (one K = 3 fp32 GEMM over the
        # source points, in place), so that the PAIR records neither load coordinates nor carry wx -- see pair_kernel<LITE>.
        self.lite = self.feats.R > 64 and os.environ.get("SLIDE_PAIR_LITE", "1") != "0"
        if self.lite:
            wxcat = np.zeros((off, 3), dtype=np.float32)
            for Pc in convs:
                sl = self.slices[Pc.prefix]
                wxcat[sl["off"]:sl["off"] + sl["N"]] = sl["wx_value"]
            xoff, xld = b.weight_matrix(wxcat)
            b.gemm(self.xyz, (xoff, xld, off, 3), self.U, resid=self.U, note=self.name + ".Ux")

    def linear(self, Pc, out, act=None, resid=None, xfr=NO_XF, stats=None, st_choff=0, note=""):
        b = self.b
        if not self.factored:
            W, bias = _conv(b, Pc)
            b.gemm(self.G, W, out, bias=bias, act=act, resid=resid, xfr=xfr, stats=stats, st_R=self.R, st_choff=st_choff,
                   note=note)
            return
        sl = self.slices[Pc.prefix]
        b.pair(self.U.cols(sl["off"], sl["N"]), self.xyz, self.ctr, self.idx, self.K, out,
               -1 if getattr(self, "lite", False) else sl["wx"], sl["wc"],
               bias=sl["bias"], d2=self.d2 if self.mode == 1 else None, wd=sl["wd"], ww=sl["ww"], act=act, resid=resid,
               xfr=xfr, stats=stats, st_choff=st_choff, note=note)


def _align4(x):
    return (x + 3) // 4 * 4


def _mlp_out_channels(P):
    """Output channels of an Mlp_plus_t_emb (its last conv)."""
    j, last = 0, "second_mlp.0"
    while P.has("rest_mlp.%d.weight" % (3 * j)):
        last = "rest_mlp.%d" % (3 * j)
        j += 1
    return int(P[last + ".weight"].shape[0])


def _lower_mlp(ctx, P, G, R, name, out=None, use_t=False, use_cond=False, res=True, xf_in=NO_XF, first_cols=None,
               first_ev=None):
    """Mlp_plus_t_emb with bn_first=False.  G: input [B*R, Cin] (xf_in = transform still to be applied to it).
    res=True (res_connect): returns the materialised output H [B*R, Cout].
    res=False (Pnet2Stage's MLPs): returns (raw, xf) -- the last conv's raw output and the transform its consumers apply.
    first_cols / first_ev: the first conv acts on cat[G, v broadcast over rows]; G's share of the weight is
    columns first_cols and the other share enters as the per-sample vector first_ev = (tensor, rows per sample)."""
    b = ctx.b
    if P.has("first_conv.weight"):
        raise NotImplementedError("first_conv")
    stages = [("first_mlp.0", "first_mlp.1"), ("second_mlp.0", "second_mlp.1")]
    j = 0
    while P.has("rest_mlp.%d.weight" % (3 * j)):
        stages.append(("rest_mlp.%d" % (3 * j), "rest_mlp.%d" % (3 * j + 1)))
        j += 1
    grouped = G if isinstance(G, _Grouped) else None
    prev, xf_prev = G, xf_in
    for si, (ck, gk) in enumerate(stages):
        if si == 0 and grouped is not None:
            N = P[ck + ".weight"].shape[0]
        else:
            W, bias = _conv(b, P.sub(ck), cols=first_cols if si == 0 else None)
            N = W[2]
        nnorm, cg = _gn_dims(N)
        assert P[gk + ".group_norm.weight"].shape[0] == nnorm
        raw = b.tensor("%s.raw%d" % (name, si), R, N, B=G.B)
        st = b.stats("%s.st%d" % (name, si), nnorm, cg, R, R * cg, B=G.B)
        if si == 0 and grouped is not None:
            grouped.linear(P.sub(ck), raw, stats=st, note="%s.conv%d" % (name, si))
        elif si == 0 and first_ev is not None:
            b.gemm(prev, W, raw, bias=bias, xfa=xf_prev, ev=first_ev[0], ev_div=first_ev[1], stats=st,
                   note="%s.conv%d" % (name, si))
        else:
            b.gemm(prev, W, raw, bias=bias, xfa=xf_prev, stats=st, note="%s.conv%d" % (name, si))
        addvec, addmode = None, 0
        if si == 0 and P.has("fc.weight"):
            assert use_t and ctx.t_src is not None, "module has a timestep projection but no timestep source"
            Wt, bt = _conv(b, P.sub("fc"))
            addvec = b.tensor("%s.ttab" % name, ctx.t_src.rows, N, B=1)
            ctx.setup.append(dict(A=ctx.t_src, W=Wt, out=addvec, bias=bt, note="%s.fc(t)" % name))
            addmode = 1
        if si == 1 and P.has("fc_condition.weight"):
            assert use_cond and ctx.cond_src is not None, "module has a condition projection but no condition"
            Wc, bc = _conv(b, P.sub("fc_condition"))
            addvec = b.tensor("%s.cvec" % name, 1, N, B=G.B)
            ctx.project(A=ctx.cond_src, W=Wc, out=addvec, bias=bc, note="%s.fc_condition" % name)
            addmode = 0
        if si == len(stages) - 1 and P.has("fc_second_condition.weight"):
            assert si >= 2 and ctx.cond2_src is not None, "second condition needs a rest_mlp stage and a source"
            Wc, bc = _conv(b, P.sub("fc_second_condition"))
            addvec = b.tensor("%s.c2vec" % name, 1, N, B=G.B)
            ctx.project(A=ctx.cond2_src, W=Wc, out=addvec, bias=bc, note="%s.fc_second_condition" % name)
            addmode = 0
        xf_prev = XF(stats=st.tensor, cg=cg, nnorm=nnorm, choff=0, gamma=b.weight(P[gk + ".group_norm.weight"]),
                     beta=b.weight(P[gk + ".group_norm.bias"]), R=R, count=R * cg, relu=True, addvec=addvec,
                     addmode=addmode)
        prev = raw
    if not res:
        return prev, xf_prev
    N = prev.C
    if out is None:
        out = b.tensor("%s.out" % name, R, N, B=G.B)
    if P.has("res_connect.weight") and grouped is not None:
        grouped.linear(P.sub("res_connect"), out, resid=prev, xfr=xf_prev, note="%s.res" % name)
    elif P.has("res_connect.weight"):
        Wr, br = _conv(b, P.sub("res_connect"))
        b.gemm(G, Wr, out, bias=br, resid=prev, xfr=xf_prev, note="%s.res" % name)
    else:
        raise NotImplementedError("identity residual (mlp_spec[0] == mlp_spec[-1])")
    return out


def _lower_attention(ctx, P, q_feat, G, H, npnt, K, out, name):
    """AttentionModule(attention_bn=True, transform_grouped_feat_out=True); counts == K ('nn' neighbours)."""
    return _attention_tail(ctx, P, _attention_keys(ctx, P, q_feat, G, npnt, K, name), H, npnt, K, out, name)


def _attention_keys(ctx, P, q_feat, G, npnt, K, name):
    """Query / key branch of AttentionModule up to the input of weight_conv.5: depends only on the query features and
    the grouped tensor, NOT on the shared MLP's output, so callers emit it on the side branch next to that MLP."""
    b = ctx.b
    att = ctx.cfg["attention_setting"]
    assert att["attention_bn"] and att["transform_grouped_feat_out"]
    B = G.B
    Wq, bq = _conv(b, P.sub("feat_conv"))
    C1, C2 = Wq[2], P["grouped_feat_conv.weight"].shape[0]
    nn1, cg1 = _gn_dims(C1 + C2)
    st1 = b.stats(name + ".st_cat", nn1, cg1, npnt * K, npnt * K * cg1, B=B)
    q = b.tensor(name + ".q", npnt, C1, B=B)
    k = b.tensor(name + ".k", npnt * K, C2, B=B)
    b.gemm(q_feat, Wq, q, bias=bq, act="relu", stats=st1, st_R=npnt, st_choff=0, st_weight=K, note=name + ".q")
    G.linear(P.sub("grouped_feat_conv"), k, act="relu", stats=st1, st_choff=C1, note=name + ".k")
    g1 = b.weight(P["weight_conv.1.group_norm.weight"])
    be1 = b.weight(P["weight_conv.1.group_norm.bias"])
    W1q, _ = _conv(b, P.sub("weight_conv.2"), cols=(0, C1))
    W1k, b1 = _conv(b, P.sub("weight_conv.2"), cols=(C1, C1 + C2))
    inter = W1q[2]
    qp = b.tensor(name + ".qp", npnt, inter, B=B)
    b.gemm(q, W1q, qp, xfa=XF(stats=st1.tensor, cg=cg1, nnorm=nn1, choff=0, gamma=g1, beta=be1, R=npnt,
                              count=npnt * K * cg1), note=name + ".w1q")
    nn2, cg2 = _gn_dims(inter)
    st2 = b.stats(name + ".st_s1", nn2, cg2, npnt * K, npnt * K * cg2, B=B)
    s1 = b.tensor(name + ".s1", npnt * K, inter, B=B)
    b.gemm(k, W1k, s1, bias=b1, act="relu", ev=qp, ev_div=K, stats=st2, st_R=npnt * K,
           xfa=XF(stats=st1.tensor, cg=cg1, nnorm=nn1, choff=C1, gamma=g1, beta=be1, R=npnt * K,
                  count=npnt * K * cg1), note=name + ".w1k")
    return s1, st2, nn2, cg2


def _attention_tail(ctx, P, keys, H, npnt, K, out, name):
    """Value projection + weight_conv.5 + soft-max over the neighbours + weighted sum."""
    b = ctx.b
    att = ctx.cfg["attention_setting"]
    B = H.B
    s1, st2, nn2, cg2 = keys
    W2, b2 = _conv(b, P.sub("weight_conv.5"))
    Co = W2[2]
    Wv, bv = _conv(b, P.sub("feat_out_conv.0"))
    v = b.tensor(name + ".v", npnt * K, Co, B=B)
    if att["last_activation"]:
        nn3, cg3 = _gn_dims(Co)
        st3 = b.stats(name + ".st_v", nn3, cg3, npnt * K, npnt * K * cg3, B=B)
        b.gemm(H, Wv, v, bias=bv, stats=st3, st_R=npnt * K, note=name + ".v")
        xfv = XF(stats=st3.tensor, cg=cg3, nnorm=nn3, choff=0, gamma=b.weight(P["feat_out_conv.1.group_norm.weight"]),
                 beta=b.weight(P["feat_out_conv.1.group_norm.bias"]), R=npnt * K, count=npnt * K * cg3, relu=True)
    else:
        b.gemm(H, Wv, v, bias=bv, note=name + ".v")
        xfv = NO_XF
    assert out.C == Co
    if os.environ.get("SLIDE_FUSE_SOFTMAX", "1") == "0":  # A/B switch for tuning: unfused score GEMM + soft-max record
        scores = b.tensor(name + ".scores", npnt * K, Co, B=B)
        b.gemm(s1, W2, scores, bias=b2,
               xfa=XF(stats=st2.tensor, cg=cg2, nnorm=nn2, choff=0, gamma=b.weight(P["weight_conv.4.group_norm.weight"]),
                      beta=b.weight(P["weight_conv.4.group_norm.bias"]), R=npnt * K, count=npnt * K * cg2),
               note=name + ".w2")
        b.softmax_wsum(scores, v, xfv, out, B * npnt, K, note=name + ".softmax")
        return out
    # scores = weight_conv.5(...) are never written: the GEMM's epilogue takes the soft-max over the K neighbours of
    # each point and reduces the (normalised, ReLU'ed) values with it (GEMM_SMK)
    b.gemm(s1, W2, out, bias=b2, resid=v, xfr=xfv, smk=K,
           xfa=XF(stats=st2.tensor, cg=cg2, nnorm=nn2, choff=0, gamma=b.weight(P["weight_conv.4.group_norm.weight"]),
                  beta=b.weight(P["weight_conv.4.group_norm.bias"]), R=npnt * K, count=npnt * K * cg2),
           note=name + ".w2+softmax")
    return out


def _sa_geometry(ctx, xyz, npoint, nsample, name):
    """The coordinate-only part of a set-abstraction level: FPS pick, centres, neighbour indices."""
    b = ctx.b
    N, B = xyz.R, xyz.B
    pick = None
    if N <= npoint:
        new_xyz, npnt = xyz, N
    else:
        npnt = npoint
        if ctx.geom is not None:
            raise NotImplementedError("frozen-coordinate geometry with a down-sampling level (FPS moves with the features' order)")
        pick = b.tensor(name + ".fps", 1, npnt, B=B, dtype="i32")
        b.fps(0, xyz, npnt, pick, note=name + ".fps")
        new_xyz = b.tensor(name + ".new_xyz", npnt, 3, B=B)
        b.gather_rows(xyz, pick, npnt, new_xyz, note=name + ".gather_xyz")
    K = min(nsample, N)
    idx = b.tensor(name + ".idx", npnt, K, B=B, dtype="i32")
    ctx.knn(new_xyz, xyz, K, idx, note=name + ".knn")
    return dict(pick=pick, new_xyz=new_xyz, npnt=npnt, K=K, idx=idx)


def _lower_sa(ctx, P, xyz, feats, npoint, nsample, name, geom=None, side_extra=None):
    """-> (new_xyz [B*np,3], features [B*np,Cout]).  geom: this level's _sa_geometry if it was emitted earlier (hoisted
    onto an earlier module's side branch); side_extra(new_xyz): extra records for this module's side branch."""
    b, cfg = ctx.b, ctx.cfg
    B = xyz.B
    if geom is None:
        geom = _sa_geometry(ctx, xyz, npoint, nsample, name)
    new_xyz, npnt, K, idx = geom["new_xyz"], geom["npnt"], geom["K"], geom["idx"]
    if geom["pick"] is None:
        q_feat = feats
    else:
        q_feat = b.tensor(name + ".qfeat", npnt, feats.C, B=B)
        b.gather_rows(feats, geom["pick"], npnt, q_feat, note=name + ".gather_feat")
    inc_abs, inc_ctr = cfg["include_abs_coordinate"], cfg.get("include_center_coordinate", False)
    assert cfg["model.use_xyz"]
    Pa = P.sub("attention_modules.0")
    G = _Grouped(ctx, 0, feats, xyz, new_xyz, idx, K, name, inc_abs=inc_abs, inc_ctr=inc_ctr)
    G.plan([P.sub("mlps.0.first_mlp.0"), P.sub("mlps.0.res_connect"), Pa.sub("grouped_feat_conv")])
    with b.side_branch():
        keys = _attention_keys(ctx, Pa, q_feat, G, npnt, K, name + ".att")
        if side_extra is not None:
            side_extra(new_xyz)
    H = _lower_mlp(ctx, P.sub("mlps.0"), G, npnt * K, name + ".mlp", use_t=True, use_cond=True)
    b.join(name + ".join")
    out = b.tensor(name + ".out", npnt, H.C, B=B)
    _attention_tail(ctx, Pa, keys, H, npnt, K, out, name + ".att")
    return new_xyz, out


def _fp_geometry(ctx, unknown, known, K, name):
    b = ctx.b
    idx = b.tensor(name + ".idx", unknown.R, K, B=unknown.B, dtype="i32")
    d2 = b.tensor(name + ".d2", unknown.R, K, B=unknown.B)
    ctx.knn(unknown, known, K, idx, d2=d2, note=name + ".knn")
    return idx, d2


def _lower_fp(ctx, P, unknown, known, unknow_feats, known_feats, K, name, out=None, side_extra=None, geom=None):
    b = ctx.b
    n, B = unknown.R, unknown.B
    idx, d2 = geom if geom is not None else _fp_geometry(ctx, unknown, known, K, name)
    Pa = P.sub("attention_module")
    G = _Grouped(ctx, 1, known_feats, known, unknown, idx, K, name, d2=d2)
    G.plan([P.sub("mlp1.first_mlp.0"), P.sub("mlp1.res_connect"), Pa.sub("grouped_feat_conv")])
    d = _mlp_out_channels(P.sub("mlp1"))
    cat = b.tensor(name + ".cat", n, d + unknow_feats.C + 3, B=B)
    with b.side_branch():
        keys = _attention_keys(ctx, Pa, unknow_feats, G, n, K, name + ".att")
        # the skip features and coordinates of mlp2's input exist before this module starts: copied on the side branch,
        # under the pair-level GEMMs of mlp1 (they were two serial records after the attention tail)
        b.copy_cols(unknow_feats, cat.cols(d, unknow_feats.C), note=name + ".cat_skip")
        b.copy_cols(unknown.cols(0, 3), cat.cols(d + unknow_feats.C, 3), note=name + ".cat_xyz")
        if side_extra is not None:
            side_extra()
    H1 = _lower_mlp(ctx, P.sub("mlp1"), G, n * K, name + ".mlp1")
    assert H1.C == d
    b.join(name + ".join")
    _attention_tail(ctx, Pa, keys, H1, n, K, cat.cols(0, d), name + ".att")
    return _lower_mlp(ctx, P.sub("mlp2"), cat, n, name + ".mlp2", out=out, use_t=True, use_cond=True)


def _lower_feature_map(ctx, P, xyz, feats, new_xyz, q_feat, nsample, name, out):
    b, cfg = ctx.b, ctx.cfg
    npnt, B = new_xyz.R, new_xyz.B
    K = min(nsample, xyz.R)
    idx = b.tensor(name + ".idx", npnt, K, B=B, dtype="i32")
    ctx.knn(new_xyz, xyz, K, idx, note=name + ".knn")
    inc_abs, inc_ctr = cfg["include_abs_coordinate"], cfg.get("include_center_coordinate", False)
    Pa = P.sub("attention_module")
    G = _Grouped(ctx, 0, feats, xyz, new_xyz, idx, K, name, inc_abs=inc_abs, inc_ctr=inc_ctr)
    G.plan([P.sub("mlp.first_mlp.0"), P.sub("mlp.res_connect"), Pa.sub("grouped_feat_conv")])
    with b.side_branch():
        keys = _attention_keys(ctx, Pa, q_feat, G, npnt, K, name + ".att")
    H = _lower_mlp(ctx, P.sub("mlp"), G, npnt * K, name + ".mlp")
    b.join(name + ".join")
    return _attention_tail(ctx, Pa, keys, H, npnt, K, out, name + ".att")


def t_embedding_source(b, P, cfg, T, name):
    """Timestep-embedding table for t = 0..T-1: swish(fc_t2(swish(fc_t1(calc_t_emb(t))))) -> [T, 4*t_dim].
    Returns (tensor, emit) where emit() appends the ops (setup segment).  calc_t_emb:
    pointnet2/models/pointnet2_ssg_sem.py:14-31; fc_t1/fc_t2: pointnet2_with_pcld_condition.py:354-359."""
    t_dim = cfg["t_dim"]
    half = t_dim // 2
    import torch
    freq = torch.exp(torch.arange(half) * -(np.log(10000) / (half - 1))).numpy()  # same ops as the reference
    freq_off = b.weight(freq)
    ts = b.tensor(name + ".ts", 1, T, B=1, ld=T)  # dense f32 [T]
    emb = b.tensor(name + ".temb0", T, t_dim, B=1)
    W1, b1 = _conv(b, P.sub("fc_t1"))
    W2, b2 = _conv(b, P.sub("fc_t2"))
    h1 = b.tensor(name + ".temb1", T, W1[2], B=1)
    h2 = b.tensor(name + ".temb2", T, W2[2], B=1)

    def emit():
        b.temb(ts, freq_off, half, emb, note=name + ".calc_t_emb")
        b.gemm(emb, W1, h1, bias=b1, act="swish", note=name + ".fc_t1")
        b.gemm(h1, W2, h2, bias=b2, act="swish", note=name + ".fc_t2")
    return ts, h2, emit


def lower_cloud_net(b, P, cfg, X, n_points, name, T=None, labels=None, out=None, factor_group=None, defer_geometry=False):
    """Lower PointNet2CloudCondition (no condition cloud).

    X: arena tensor [B*n_points, 3 + in_fea_dim] (xyz first).  T: number of timesteps (DDPM denoisers) or None.
    labels: i32 tensor [B, 1] (class condition) or None.
    Returns dict(out=Tensor [B*n_points, out_channels], emit_setup=callable, inputs={...}).
    The network ops are appended to b immediately; emit_setup() must be called inside the setup segment
    (before the network runs) -- it emits the timestep tables and condition vectors.
    """
    assert not cfg.get("include_local_feature", True) and not cfg.get("include_global_feature", False)
    arch = cfg["architecture"]
    assert arch["neighbor_definition"] == "nn" and arch.get("use_knn_FP", False)
    assert not arch.get("include_grouper", False)
    if cfg.get("use_position_encoding", False):
        raise NotImplementedError("position encoding")
    B = X.B
    inputs = {}
    setup_pre = []
    t_src = None
    if T is not None and cfg["include_t"]:
        ts, t_src, emit_t = t_embedding_source(b, P, cfg, T, name)
        inputs["ts_table"] = ts
        setup_pre.append(emit_t)
    cond_src = None
    if labels is not None and cfg["include_class_condition"]:
        emb_w = P["class_emb.weight"]
        table = b.tensor(name + ".class_emb", emb_w.shape[0], emb_w.shape[1], B=1)
        inputs["class_emb"] = (table, emb_w)
        cond_src = b.tensor(name + ".cond", 1, emb_w.shape[1], B=B)
        lab_all = labels

        def emit_c():
            # class_emb(label): a row gather from the table, all samples index the same (B=1) table
            b._emit("SLIDE_OP_GATHER_ROWS", {"GA_SRC": table.off, "GA_LDS": table.ld, "GA_N": table.R,
                                             "GA_IDX": lab_all.off, "GA_M": B, "GA_DST": cond_src.off,
                                             "GA_LDD": cond_src.ld, "GA_NCOLS": table.C, "GA_B": 1},
                    note=name + ".class_emb")
        setup_pre.append(emit_c)
    ctx = _Ctx(b, cfg, t_src, cond_src)
    ctx.factor_group = factor_group
    if defer_geometry:
        ctx.geom = []

    Fin = X.C - 3
    assert X.R == n_points
    xyz = X.cols(0, 3)
    if cfg["attach_position_to_input_feature"]:
        feats = b.tensor(name + ".feat0", n_points, Fin + 3, B=B)
        if Fin > 0:
            b.copy_cols(X.cols(3, Fin), feats.cols(0, Fin), note=name + ".in_feat")
        if ctx.geom is not None:  # frozen coordinates: the attached position columns are written once per chain
            ctx.geom.append(lambda: b.copy_cols(X.cols(0, 3), feats.cols(Fin, 3), note=name + ".in_xyz"))
        else:
            b.copy_cols(X.cols(0, 3), feats.cols(Fin, 3), note=name + ".in_xyz")
    else:
        assert Fin > 0
        feats = X.cols(3, Fin)
    l_xyz, l_feat = [xyz], [feats]
    n_fp = len(arch["decoder_feature_dim"]) - 1
    # Down-sampling networks (refinement / encoder clouds): everything that depends on coordinates only -- the FPS picks,
    # centres and neighbour searches of levels >= 1 and the propagation modules' searches -- is emitted on the side branch
    # of level 0, underneath its pair-level GEMMs, instead of serially in front of each module (FPS is one CTA per cloud).
    sa_geoms, fp_geoms = {}, {}
    hoist = ctx.geom is None and n_points > arch["npoint"][0] and os.environ.get("SLIDE_HOIST_GEOMETRY", "1") != "0"

    def hoisted(new_xyz0):
        xs = [xyz, new_xyz0]
        for i in range(1, len(arch["npoint"])):
            sa_geoms[i] = _sa_geometry(ctx, xs[i], arch["npoint"][i], arch["nsample"][i], "%s.SA%d" % (name, i))
            xs.append(sa_geoms[i]["new_xyz"])
        for i in range(-1, -(n_fp + 1), -1):
            fp_geoms[i] = _fp_geometry(ctx, xs[i - 1], xs[i], arch.get("K", 3), "%s.FP%d" % (name, n_fp + i))

    for i, (npoint, nsample) in enumerate(zip(arch["npoint"], arch["nsample"])):
        nx, nf = _lower_sa(ctx, P.sub("SA_modules.%d" % i), l_xyz[i], l_feat[i], npoint, nsample,
                           "%s.SA%d" % (name, i), geom=sa_geoms.get(i), side_extra=hoisted if (hoist and i == 0) else None)
        l_xyz.append(nx)
        l_feat.append(nf)
    transform = cfg.get("transform_output", True)
    head_in = None
    for i in range(-1, -(n_fp + 1), -1):
        dst = None
        if i == -n_fp:
            d0 = arch["decoder_feature_dim"][0]
            if transform:
                head_in = b.tensor(name + ".head_in", n_points, d0 + 3, B=B)
                dst = head_in.cols(0, d0)
            elif out is not None:
                dst = out
        extra = None
        if head_in is not None and i == -n_fp:  # the head's coordinate columns: copied under the last module's GEMMs
            extra = lambda: b.copy_cols(xyz, head_in.cols(arch["decoder_feature_dim"][0], 3), note=name + ".head_xyz")  # noqa: E731
        l_feat[i - 1] = _lower_fp(ctx, P.sub("FP_modules.%d" % (n_fp + i)), l_xyz[i - 1], l_xyz[i], l_feat[i - 1],
                                  l_feat[i], arch.get("K", 3), "%s.FP%d" % (name, n_fp + i), out=dst, side_extra=extra,
                                  geom=fp_geoms.get(i))
    result = l_feat[0]
    if transform:
        d0 = arch["decoder_feature_dim"][0]
        W0, b0 = _conv(b, P.sub("fc_lyaer.0"))
        raw = b.tensor(name + ".head_raw", n_points, W0[2], B=B)
        assert W0[2] % 32 == 0
        cg = W0[2] // 32
        st = b.stats(name + ".head_st", W0[2], cg, n_points, n_points * cg, B=B)
        b.gemm(head_in, W0, raw, bias=b0, stats=st, st_R=n_points, note=name + ".head0")
        W3, b3 = _conv(b, P.sub("fc_lyaer.3"))
        if out is None:
            out = b.tensor(name + ".eps", n_points, W3[2], B=B)
        b.gemm(raw, W3, out, bias=b3,
               xfa=XF(stats=st.tensor, cg=cg, nnorm=W0[2], choff=0, gamma=b.weight(P["fc_lyaer.1.weight"]),
                      beta=b.weight(P["fc_lyaer.1.bias"]), R=n_points, count=n_points * cg, relu=True),
               note=name + ".head3")
        result = out

    def emit_setup():
        for fn in setup_pre:
            fn()
        for g in ctx.setup:
            b.gemm(g["A"], g["W"], g["out"], bias=g["bias"], note=g["note"])
    def emit_geometry():
        for emit in ctx.geom or []:
            emit()
    return dict(out=result, emit_setup=emit_setup, emit_geometry=emit_geometry, inputs=inputs, levels=(l_xyz, l_feat))


# ---------------------------------------------------------------------------------------------------
# autoencoder encode
# ---------------------------------------------------------------------------------------------------
def _lower_pnet2stage(ctx, P, X, n_points, name):
    """Pnet2Stage.forward (pointnet2/models/pnet.py:26-40), remove_last_activation=False.  X [B*n, C] -> [B, C_out].
    cat[feature, max-pooled feature broadcast over points] is never formed: the second MLP's first conv is split
    into its per-point share (on the rows) and its global share (one row per sample, added in the epilogue)."""
    b = ctx.b
    B = X.B
    raw, xf = _lower_mlp(ctx, P.sub("mlp1"), X, n_points, name + ".mlp1", res=False)
    C1 = raw.C
    g1 = b.tensor(name + ".g1", 1, C1, B=B)
    b.colmax(raw, xf, n_points, g1, note=name + ".maxpool1")
    Wg, _ = _conv(b, P.sub("mlp2.first_mlp.0"), cols=(C1, 2 * C1))
    gp = b.tensor(name + ".gproj", 1, Wg[2], B=B)
    b.gemm(g1, Wg, gp, note=name + ".mlp2.conv0(global)")
    raw2, xf2 = _lower_mlp(ctx, P.sub("mlp2"), raw, n_points, name + ".mlp2", res=False, xf_in=xf, first_cols=(0, C1),
                           first_ev=(gp, n_points))
    out = b.tensor(name + ".global", 1, raw2.C, B=B)
    b.colmax(raw2, xf2, n_points, out, note=name + ".maxpool2")
    return out


def lower_encoder_net(b, P, cfg, X, n_points, name, labels):
    """PointNet2Encoder.forward (pointnet2/models/pointnet2_feature_extractor.py:143-218), no timestep.
    X [B*n_points, 3 + in_fea_dim].  Returns dict(out=[B*np_last, C_last], xyz=[B*np_last, 3])."""
    arch = cfg["architecture"]
    assert arch["neighbor_definition"] == "nn" and not cfg["include_t"]
    B = X.B
    class_src = None
    if labels is not None and cfg["include_class_condition"]:
        emb_w = P["class_emb.weight"]
        table = b.tensor(name + ".class_emb", emb_w.shape[0], emb_w.shape[1], B=1)
        class_src = b.tensor(name + ".cond", 1, emb_w.shape[1], B=B)
        b._emit("SLIDE_OP_GATHER_ROWS", {"GA_SRC": table.off, "GA_LDS": table.ld, "GA_N": table.R, "GA_IDX": labels.off,
                                         "GA_M": B, "GA_DST": class_src.off, "GA_LDD": class_src.ld,
                                         "GA_NCOLS": table.C, "GA_B": 1}, note=name + ".class_emb")
    Fin = X.C - 3
    assert X.R == n_points and Fin == cfg["in_fea_dim"]
    xyz = X.cols(0, 3)
    if cfg["attach_position_to_input_feature"]:
        feats = b.tensor(name + ".feat0", n_points, Fin + 3, B=B)
        if Fin > 0:
            b.copy_cols(X.cols(3, Fin), feats.cols(0, Fin), note=name + ".in_feat")
        b.copy_cols(X.cols(0, 3), feats.cols(Fin, 3), note=name + ".in_xyz")
    else:
        feats = X.cols(3, Fin)
    ctx = _Ctx(b, cfg, None, class_src, None, inline=True)
    if cfg.get("include_global_feature", False):
        if cfg.get("global_feature_remove_last_activation", True):
            raise NotImplementedError("global_feature_remove_last_activation")
        # global_pnet input = [xyz, input features] = the raw cloud rows (pointnet2_feature_extractor.py:186-193)
        g = _lower_pnet2stage(ctx, P.sub("global_pnet"), X, n_points, name + ".pnet")
        ctx.cond_src, ctx.cond2_src = g, class_src
    l_xyz, l_feat = [xyz], [feats]
    sa_geoms = {}
    hoist = n_points > arch["npoint"][0] and os.environ.get("SLIDE_HOIST_GEOMETRY", "1") != "0"

    def hoisted(new_xyz0):  # levels >= 1: FPS / centres / neighbour searches under level 0's GEMMs (see lower_cloud_net)
        cur = new_xyz0
        for i in range(1, len(arch["npoint"])):
            sa_geoms[i] = _sa_geometry(ctx, cur, arch["npoint"][i], arch["nsample"][i], "%s.SA%d" % (name, i))
            cur = sa_geoms[i]["new_xyz"]

    for i, (npoint, nsample) in enumerate(zip(arch["npoint"], arch["nsample"])):
        nx, nf = _lower_sa(ctx, P.sub("SA_modules.%d" % i), l_xyz[i], l_feat[i], npoint, nsample, "%s.SA%d" % (name, i),
                           geom=sa_geoms.get(i), side_extra=hoisted if (hoist and i == 0) else None)
        l_xyz.append(nx)
        l_feat.append(nf)
    tables = [(table, emb_w)] if class_src is not None else []
    return dict(out=l_feat[-1], xyz=l_xyz[-1], class_tables=tables)


def lower_encode(b, P, enc_cfg, kp_cfg, cloud, keypoint, labels, noises=None, name="enc"):
    """PointAutoencoder.encode (pointnet2/models/autoencoder.py:37-40; keypoint_encoder.propagate_feature with the KL
    branch, point_upsample_decoder.py:106-147).  cloud [B*N, 6], keypoint [B*16, 3]; noises = (n1 [B*16, C1],
    n2 [B*16, C2]) arena tensors holding the two posterior draws, or None for the posterior mode.
    Returns dict(out=[B*16, C1 + C2], class_tables)."""
    B = cloud.B
    enc = lower_encoder_net(b, P.sub("encoder"), enc_cfg, cloud, cloud.R, name + ".encoder", labels)
    Pk = P.sub("keypoint_encoder")
    fx = lower_encoder_net(b, Pk.sub("feature_extractor"), kp_cfg, keypoint, keypoint.R, name + ".kp_fx", labels)
    C1 = fx["out"].C // 2
    C2 = kp_cfg["feature_mapper_setting"]["out_dim"]
    latent = b.tensor(name + ".latent", keypoint.R, C1 + C2, B=B)
    b.kl(fx["out"], latent.cols(0, C1), noise=None if noises is None else noises[0], note=name + ".kl_extractor")
    ctx = _Ctx(b, kp_cfg)
    mapped = b.tensor(name + ".mapped", keypoint.R, 2 * C2, B=B)
    _lower_feature_map(ctx, Pk.sub("feature_mapper"), enc["xyz"], enc["out"], keypoint.cols(0, 3), latent.cols(0, C1),
                       kp_cfg["feature_mapper_setting"]["nsample"], name + ".fm", mapped)
    assert not ctx.setup
    b.kl(mapped, latent.cols(C1, C2), noise=None if noises is None else noises[1], note=name + ".kl_mapper")
    return dict(out=latent, class_tables=enc["class_tables"] + fx["class_tables"])


# ---------------------------------------------------------------------------------------------------
# autoencoder decode
# ---------------------------------------------------------------------------------------------------
def _lower_upsample_points(b, P, cfg, final_feature, new_xyz, start, name):
    """PointUpsampleDecoder.upsample_points (point_upsample_decoder.py:149-182).  final_feature [B*N, Cf],
    new_xyz [B*N, Cx]; start: i32 [B,1] FPS start indices (pytorch3d random_start_point).  Returns
    the upsampled cloud [B*n_out, out_dim]."""
    up = cfg["upsampling_setting"]
    if up["first_refine_coarse_points"]:
        raise NotImplementedError("first_refine_coarse_points")
    B, N = new_xyz.B, new_xyz.R
    cat = b.tensor(name + ".up_in", N, final_feature.C + new_xyz.C, B=B)
    b.copy_cols(final_feature, cat.cols(0, final_feature.C), note=name + ".up_cat_f")
    b.copy_cols(new_xyz, cat.cols(final_feature.C, new_xyz.C), note=name + ".up_cat_x")
    W, bias = _conv(b, P.sub("fc_layer"))
    disp = b.tensor(name + ".disp", N, W[2], B=B)
    b.gemm(cat, W, disp, bias=bias, note=name + ".fc_layer")
    factor = up["point_upsample_factor"]
    out_dim = cfg["out_dim"]
    in_dim = cfg.get("in_position_and_normal_dim", out_dim)
    pts = b.tensor(name + ".split", N * factor, out_dim, B=B)
    b.upsample(new_xyz, min(in_dim, out_dim), disp, pts, factor, up["output_scale_factor"], note=name + ".upsample")
    n_out = up["num_output_points"]
    assert N * factor >= n_out
    if N * factor > n_out:
        sel = b.tensor(name + ".sel", 1, n_out, B=B, dtype="i32")
        b.fps(1, pts, n_out, sel, start=start, note=name + ".fps_p3d")
        res = b.tensor(name + ".points", n_out, out_dim, B=B)
        b.gather_rows(pts, sel, n_out, res, note=name + ".masked_gather")
        return res
    return pts


def lower_decode(b, P, decoder_cfgs, keypoint, feature, labels, name="ae"):
    """PointAutoencoder.decode.  keypoint [B*16, 3], feature [B*16, 48] arena tensors, labels i32 [B,1].
    Returns dict(out=[B*2048, 6], starts=[i32 [B,1] per level], emit_setup, inputs)."""
    B = keypoint.B
    starts = [b.tensor("%s.start%d" % (name, i), 1, B, B=1, dtype="i32") for i in range(len(decoder_cfgs))]
    setups, class_tables = [], []
    new_xyz = _lower_upsample_points(b, P.sub("keypoint_encoder"), decoder_cfgs[0], feature, keypoint, starts[0],
                                     name + ".L1")
    xyzs, feats = [keypoint, new_xyz], [feature]
    for i, cfg in enumerate(decoder_cfgs[1:]):
        Pd = P.sub("decoder.decoders.%d" % i)
        lname = "%s.L%d" % (name, i + 2)
        cur = xyzs[i + 1]
        d0 = cfg["architecture"]["decoder_feature_dim"][0]
        md = cfg["feature_mapper_setting"]["out_dim"]
        final = b.tensor(lname + ".final", cur.R, d0 + md, B=B)
        net = lower_cloud_net(b, Pd.sub("feature_extractor"), cfg, cur, cur.R, lname + ".fx", T=None, labels=labels,
                              out=final.cols(0, d0))
        setups.append(net["emit_setup"])
        if "class_emb" in net["inputs"]:
            class_tables.append(net["inputs"]["class_emb"])
        ctx = _Ctx(b, cfg)
        prev_xyz = xyzs[i].cols(0, 3)
        cur_xyz = cur.cols(0, 3)
        _lower_feature_map(ctx, Pd.sub("feature_mapper"), prev_xyz, feats[i], cur_xyz, net["out"],
                           cfg["feature_mapper_setting"]["nsample"], lname + ".fm", final.cols(d0, md))
        assert not ctx.setup
        xyzs.append(_lower_upsample_points(b, Pd, cfg, final, cur, starts[i + 1], lname))
        feats.append(final)

    def emit_setup():
        for fn in setups:
            fn()
    return dict(out=xyzs[-1], starts=starts, emit_setup=emit_setup, class_tables=class_tables, levels=xyzs)
