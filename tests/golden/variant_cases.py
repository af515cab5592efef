"""Case table shared by make_golden_variants.py (REAL reference classes -> golden_variants.npz) and
tests/test_dropin_variants_cpu.py (the drop-in classes against that file): constructor arguments of the module variants the
shipped sampling configs never select -- bn_first, swish, first_conv, identity residual, no residual, no normalisation,
second condition, plain-conv attention, un-transformed values, global attention, ball-query set abstraction with pooling,
multi-scale grouping, the propagation modules' grouper.

Plain data only: nothing here imports the reference or the drop-in.  Every case is (name, class name, kwargs, input kind);
list-valued kwargs are rebuilt per call because the reference's constructors edit `mlp_spec` in place
(pointnet2_modules.py:372-377)."""
import copy

ATT = dict(use_attention_module=True, attention_bn=True, transform_grouped_feat_out=True, last_activation=True)
ATT_PLAIN = dict(use_attention_module=True, attention_bn=False, transform_grouped_feat_out=False, last_activation=False)
GATT = dict(use_global_attention_module=True, attention_bn=True, last_activation=True)

_MLP = dict(bn=True, t_dim=16, include_t=True, bias=True, res_connect=True, include_condition=True, condition_dim=9)
_SA = dict(npoint=12, radius=0.45, nsample=6, bn=True, use_xyz=True, t_dim=16, include_t=True,
           include_abs_coordinate=True, include_center_coordinate=True, bias=True, res_connect=True,
           include_condition=True, condition_dim=9, neighbor_def="radius")
_FP = dict(bn=True, t_dim=16, include_t=True, bias=True, res_connect=True, include_condition=True, condition_dim=9,
           use_xyz=True, include_abs_coordinate=True, include_center_coordinate=True)

_CASES = [
    # ---- Mlp_plus_t_emb (pointnet2_modules.py:44-176)
    ("mlp_bn_first", "Mlp_plus_t_emb", dict(_MLP, mlp_spec=[7, 12, 12, 10], bn_first=True), "mlp"),
    ("mlp_swish", "Mlp_plus_t_emb", dict(_MLP, mlp_spec=[7, 12, 12, 10], activation="swish"), "mlp"),
    ("mlp_swish_bn_first", "Mlp_plus_t_emb", dict(_MLP, mlp_spec=[7, 12, 12, 10], activation="swish", bn_first=True), "mlp"),
    ("mlp_first_conv", "Mlp_plus_t_emb", dict(_MLP, mlp_spec=[8, 12, 12, 10], first_conv=True, first_conv_in_channel=5), "mlp"),
    ("mlp_identity_res", "Mlp_plus_t_emb", dict(_MLP, mlp_spec=[10, 12, 12, 10]), "mlp"),
    ("mlp_no_res", "Mlp_plus_t_emb", dict(_MLP, mlp_spec=[7, 12, 12, 10], res_connect=False), "mlp"),
    ("mlp_no_norm_no_bias", "Mlp_plus_t_emb", dict(_MLP, mlp_spec=[7, 12, 12, 10], bn=False, bias=False), "mlp"),
    ("mlp_second_condition", "Mlp_plus_t_emb", dict(_MLP, mlp_spec=[7, 12, 12, 14, 10], include_second_condition=True,
                                                    second_condition_dim=11), "mlp"),
    ("mlp_bare", "Mlp_plus_t_emb", dict(mlp_spec=[7, 40, 36], bn=True, include_t=False, include_condition=False), "mlp"),
    # ---- AttentionModule / GlobalAttentionModule (attention.py:35-156)
    ("att_full_count", "AttentionModule", dict(C_in1=5, C_in2=14, C1=5, C2=14, C_out=20, attention_bn=True,
                                               transform_grouped_feat_out=True, last_activation=True), "att_count"),
    ("att_no_last_act", "AttentionModule", dict(C_in1=5, C_in2=14, C1=5, C2=14, C_out=20, attention_bn=True,
                                                transform_grouped_feat_out=True, last_activation=False), "att_all"),
    ("att_plain", "AttentionModule", dict(C_in1=40, C_in2=50, C1=40, C2=50, C_out=20, attention_bn=False,
                                          transform_grouped_feat_out=False, last_activation=False), "att_count"),
    ("att_plain_transform", "AttentionModule", dict(C_in1=5, C_in2=14, C1=5, C2=14, C_out=20, attention_bn=False,
                                                    transform_grouped_feat_out=True, last_activation=True), "att_all"),
    ("gatt", "GlobalAttentionModule", dict(C=12, additional_dim=3, attention_bn=True, last_activation=True), "gatt"),
    ("gatt_plain", "GlobalAttentionModule", dict(C=12, additional_dim=0, attention_bn=False, last_activation=False), "gatt"),
    # ---- set abstraction (pointnet2_modules.py:212-455)
    ("sa_radius_pool", "PointnetSAModule", dict(_SA, mlp=[6, 12, 12, 16]), "sa_pool"),
    ("sa_radius_attention", "PointnetSAModule", dict(_SA, mlp=[6, 12, 12, 16], attention_setting=ATT), "sa"),
    ("sa_radius_plain_attention", "PointnetSAModule", dict(_SA, mlp=[6, 12, 12, 16], attention_setting=ATT_PLAIN), "sa"),
    ("sa_global_attention", "PointnetSAModule", dict(_SA, mlp=[6, 12, 12, 16], attention_setting=ATT,
                                                     global_attention_setting=GATT), "sa"),
    ("sa_first_conv_swish", "PointnetSAModule", dict(_SA, mlp=[10, 12, 12, 16], first_conv=True, first_conv_in_channel=6,
                                                     activation="swish", attention_setting=ATT), "sa"),
    ("sa_bn_first_nn", "PointnetSAModule", dict(_SA, mlp=[6, 12, 12, 16], bn_first=True, neighbor_def="nn",
                                                attention_setting=ATT), "sa"),
    # (npoint=None / GroupAll abstraction is dead in the reference: its forward asserts `self.npoint is not None`, :253)
    ("sa_msg", "PointnetSAModuleMSG", dict(npoint=12, radii=[0.3, 0.6], nsamples=[4, 8], mlps=[[6, 8, 8, 10], [6, 12, 12, 14]],
                                           bn=True, use_xyz=True, t_dim=16, include_t=True, include_abs_coordinate=False,
                                           include_center_coordinate=False, bias=True, res_connect=True,
                                           include_condition=False, neighbor_def="radius"), "sa_pool"),
    # ---- propagation / feature map (pointnet2_modules.py:457-873)
    ("fp_swish_bn_first", "PointnetFPModule", dict(_FP, mlp=[6 + 7, 12, 12, 10], bn_first=True, activation="swish",
                                                   include_grouper=False), "fp"),
    ("fp_grouper_nn", "PointnetFPModule", dict(_FP, mlp=[6 + 7, 12, 12, 10], include_grouper=True, radius=0.0, nsample=5,
                                               neighbor_def="nn"), "fp"),
    ("knnfp_grouper_radius", "PointnetKnnFPModule", dict(_FP, mlp1=[7, 12, 12], mlp2=[12 + 6, 12, 10], K=4,
                                                         include_grouper=True, radius=0.5, nsample=5, neighbor_def="radius",
                                                         attention_setting=ATT), "fp"),
    ("knnfp_plain_attention", "PointnetKnnFPModule", dict(_FP, mlp1=[7, 12, 12], mlp2=[12 + 6, 12, 10], K=4,
                                                          include_grouper=False, attention_setting=ATT_PLAIN), "fp"),
    ("knnfp_first_conv_global_attention", "PointnetKnnFPModule", dict(_FP, mlp1=[9, 12, 12], mlp2=[16, 12, 10], K=4,
                                                                      first_conv=True, first_conv_in_channel1=7,
                                                                      first_conv_in_channel2=12 + 6, include_grouper=False,
                                                                      attention_setting=ATT, global_attention_setting=GATT), "fp"),
    ("knnfp_pool_bn_first", "PointnetKnnFPModule", dict(_FP, mlp1=[7, 12, 12], mlp2=[12 + 6, 12, 10], K=4, bn_first=True,
                                                        activation="swish", include_grouper=False), "fp"),
    ("fmap_radius_pool", "FeatureMapModule", dict(mlp=[6, 12, 12, 10], radius=0.5, K=6, use_xyz=True,
                                                  include_abs_coordinate=True, include_center_coordinate=True, bn=True,
                                                  bias=True, res_connect=True, neighbor_def="radius"), "fmap"),
    ("fmap_nn_attention", "FeatureMapModule", dict(mlp=[6, 12, 12, 10], radius=0.0, K=6, use_xyz=True,
                                                   include_abs_coordinate=True, include_center_coordinate=False, bn=True,
                                                   bn_first=True, bias=True, res_connect=True, neighbor_def="nn",
                                                   activation="swish", attention_setting=ATT, query_feature_dim=7), "fmap_q"),
]


def cases():
    return [(n, c, copy.deepcopy(kw), kind) for n, c, kw, kind in _CASES]


def make_inputs(kind, kw, torch, g):
    """Deterministic inputs of a case as an ordered dict name -> CPU tensor (or the string 'all')."""
    B, N, M, K = 2, 40, 12, 6
    r = lambda *s: torch.randn(*s, generator=g)
    xyz = torch.rand(B, N, 3, generator=g) - 0.5
    if kind == "mlp":
        cin = kw["first_conv_in_channel"] if kw.get("first_conv") else kw["mlp_spec"][0]
        d = {"feature": r(B, cin, M, K)}
        if kw.get("include_t", True):
            d["t_emb"] = r(B, kw.get("t_dim", 128))
        if kw.get("include_condition"):
            d["condition_emb"] = r(B, kw["condition_dim"])
        if kw.get("include_second_condition"):
            d["second_condition_emb"] = r(B, kw["second_condition_dim"])
        return d
    if kind in ("att_count", "att_all"):
        d = {"feat": r(B, kw["C_in1"], M), "grouped_feat": r(B, kw["C_in2"], M, K), "grouped_feat_out": r(B, kw["C_out"], M, K)}
        # counts from 0 (clamped to 1 by the module) to K
        d["count"] = torch.randint(0, K + 1, (B, M), generator=g) if kind == "att_count" else "all"
        return d
    if kind == "gatt":
        return {"feat": r(B, kw["C"] + kw["additional_dim"], M)}
    if kind in ("sa", "sa_pool", "sa_all"):
        d = {"xyz": xyz, "features": r(B, 6, N), "t_emb": r(B, 16)}
        if kw.get("include_condition"):
            d["condition_emb"] = r(B, 9)
        return d
    if kind == "fp":
        known = torch.rand(B, M, 3, generator=g) - 0.5
        return {"unknown": xyz, "known": known, "unknow_feats": r(B, 6, N), "known_feats": r(B, 7, M),
                "t_emb": r(B, 16), "condition_emb": r(B, 9)}
    if kind in ("fmap", "fmap_q"):
        new_xyz = torch.rand(B, M, 3, generator=g) - 0.5
        new_xyz[:, -1] += 4.0     # a query with an empty ball
        d = {"xyz": xyz, "features": r(B, 6, N), "new_xyz": new_xyz}
        if kind == "fmap_q":
            d["query_features"] = r(B, 7, M)
        return d
    raise KeyError(kind)


def call(module, kind, inp):
    """Run a case; returns a tuple of output tensors."""
    if kind == "mlp":
        return (module(inp["feature"], t_emb=inp.get("t_emb"), condition_emb=inp.get("condition_emb"),
                       second_condition_emb=inp.get("second_condition_emb")),)
    if kind in ("att_count", "att_all"):
        return (module(inp["feat"], inp["grouped_feat"], inp["grouped_feat_out"], inp["count"]),)
    if kind == "gatt":
        return (module(inp["feat"]),)
    if kind in ("sa", "sa_pool", "sa_all"):
        outs = []
        for pooling in (("max", "avg", "avg_max") if kind == "sa_pool" else ("max",)):
            outs.extend(module(inp["xyz"], inp["features"], t_emb=inp["t_emb"], condition_emb=inp.get("condition_emb"),
                               subset=True, pooling=pooling))
        return tuple(outs)
    if kind == "fp":
        return (module(inp["unknown"], inp["known"], inp["unknow_feats"], inp["known_feats"], t_emb=inp["t_emb"],
                       condition_emb=inp["condition_emb"]),)
    if kind == "fmap":
        return tuple(module(inp["xyz"], inp["features"], inp["new_xyz"], subset=False, pooling=p) for p in ("max", "avg"))
    if kind == "fmap_q":
        return (module(inp["xyz"], inp["features"], inp["new_xyz"], subset=False, features_at_new_xyz=inp["query_features"]),)
    raise KeyError(kind)
