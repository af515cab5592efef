"""The drop-in `pointnet2_ops` package keeps the reference's checkpoint ABI: modules built with the reference's
constructor arguments expose exactly the reference's state-dict keys and shapes (schema exported from the real
reference by tests/golden/make_golden.py).  Construction only -- no GPU."""
import copy

import slide_b200
from slide_b200 import weights


def _sub(schema, prefix):
    return {k[len(prefix):]: tuple(s) for k, s in schema if k.startswith(prefix)}


def test_module_state_dict_keys_match_reference_schema():
    slide_b200.install_dropin()
    from pointnet2_ops.pointnet2_modules import PointnetSAModule, PointnetKnnFPModule, FeatureMapModule
    cfg = weights.load_json("pipeline_airplane.json")
    pc = cfg["latent_ddpm"]["pointnet_config"]
    schema = weights.load_json("schema_latent_ddpm.json")
    att = pc["attention_setting"]
    # SA_modules.1 of the feature DDPM: mlp_spec [256+... ] built like models/pointnet2_ssg_sem.py:63-101
    fd = pc["architecture"]["feature_dim"]
    sa = PointnetSAModule(npoint=16, radius=0, nsample=16, mlp=[fd[1], fd[1], fd[1], fd[2]], use_xyz=True,
                          t_dim=4 * pc["t_dim"], include_t=True, include_abs_coordinate=True,
                          include_center_coordinate=True, bn_first=False, first_conv=False, first_conv_in_channel=51,
                          res_connect=True, bias=True, include_condition=True, condition_dim=128,
                          neighbor_def="nn", activation="relu", bn=True, attention_setting=copy.deepcopy(att))
    want = _sub(schema, "SA_modules.1.")
    got = {k: tuple(v.shape) for k, v in sa.state_dict().items()}
    assert got == want
    dd = pc["architecture"]["decoder_feature_dim"]
    fp = PointnetKnnFPModule(mlp1=[dd[2], dd[1], dd[1]], mlp2=[dd[1] + fd[1], dd[1], dd[1]], K=8, first_conv=False,
                             bn=True, t_dim=4 * pc["t_dim"], include_t=True, bn_first=False, res_connect=True,
                             bias=True, include_condition=True, condition_dim=128, include_grouper=False, radius=0,
                             nsample=16, use_xyz=True, include_abs_coordinate=True, include_center_coordinate=True,
                             neighbor_def="nn", activation="relu", attention_setting=copy.deepcopy(att))
    want = _sub(schema, "FP_modules.1.")
    got = {k: tuple(v.shape) for k, v in fp.state_dict().items()}
    assert got == want
    ae = weights.load_json("schema_autoencoder.json")
    d3 = cfg["autoencoder"]["decoders"][2]
    fm_att = copy.deepcopy(d3["attention_setting"])
    want = _sub(ae, "decoder.decoders.1.feature_mapper.")
    in_dim = want["mlp.first_mlp.0.weight"][1] - 9
    fm = FeatureMapModule([in_dim] + [d3["feature_mapper_setting"]["out_dim"]] * d3["feature_mapper_setting"]["mlp_depth"],
                          0, d3["feature_mapper_setting"]["nsample"], use_xyz=True, include_abs_coordinate=True,
                          include_center_coordinate=True, bn=True, bn_first=False, bias=True, res_connect=True,
                          first_conv=False, first_conv_in_channel=0, neighbor_def="nn", activation="relu",
                          attention_setting=fm_att, query_feature_dim=d3["architecture"]["decoder_feature_dim"][0])
    got = {k: tuple(v.shape) for k, v in fm.state_dict().items()}
    assert got == want


def test_public_names_exist():
    slide_b200.install_dropin()
    from pointnet2_ops import pointnet2_utils as U, pointnet2_modules as M, attention as A
    for name in ("furthest_point_sample", "gather_operation", "three_nn", "three_interpolate", "grouping_operation",
                 "ball_query", "QueryAndGroup", "GroupAll", "group_knn", "average_feature"):
        assert hasattr(U, name), name
    for name in ("PointnetSAModule", "PointnetSAModuleMSG", "PointnetFPModule", "PointnetKnnFPModule",
                 "FeatureMapModule", "Mlp_plus_t_emb", "build_shared_mlp", "MyGroupNorm", "pooling_features"):
        assert hasattr(M, name), name
    assert hasattr(A, "AttentionModule") and hasattr(A, "GlobalAttentionModule") and hasattr(A, "count_to_mask")
    from pytorch3d.ops import knn_points, knn_gather, sample_farthest_points  # noqa: F401
    from pytorch3d.ops.utils import masked_gather  # noqa: F401
    from pytorch3d.structures.pointclouds import Pointclouds  # noqa: F401


def test_ext_argument_checks_fail_like_the_reference():
    """The pybind functions reject CPU, non-contiguous and wrong-dtype tensors through CHECK_CONTIGUOUS / CHECK_IS_FLOAT /
    CHECK_IS_INT (RuntimeError, _ext-src/include/utils.h:5-25) and have no CPU path ("CPU not supported", sampling.cpp:34).
    The drop-in raises the same way, before any library call -- so this runs without a GPU."""
    import pytest
    import torch
    slide_b200.install_dropin()
    from pointnet2_ops import _ext
    f = torch.zeros(2, 4, 8)
    xyz = torch.zeros(2, 8, 3)
    i2 = torch.zeros(2, 5, dtype=torch.int32)
    i3 = torch.zeros(2, 5, 3, dtype=torch.int32)
    w = torch.zeros(2, 5, 3)
    calls = {
        "gather_points": lambda **k: _ext.gather_points(k.get("a", f), k.get("i", i2)),
        "gather_points_grad": lambda **k: _ext.gather_points_grad(k.get("a", torch.zeros(2, 4, 5)), k.get("i", i2), 8),
        "furthest_point_sampling": lambda **k: _ext.furthest_point_sampling(k.get("a", xyz), 4),
        "three_nn": lambda **k: _ext.three_nn(k.get("a", xyz), xyz),
        "three_interpolate": lambda **k: _ext.three_interpolate(k.get("a", f), k.get("i", i3), w),
        "three_interpolate_grad": lambda **k: _ext.three_interpolate_grad(k.get("a", torch.zeros(2, 4, 5)), k.get("i", i3), w, 8),
        "ball_query": lambda **k: _ext.ball_query(k.get("a", xyz), xyz, 0.2, 4),
        "group_points": lambda **k: _ext.group_points(k.get("a", f), k.get("i", i3)),
        "group_points_grad": lambda **k: _ext.group_points_grad(k.get("a", torch.zeros(2, 4, 5, 3)), k.get("i", i3), 8),
    }
    for name, call in calls.items():
        with pytest.raises(RuntimeError, match="CPU not supported|must be a CUDA tensor"):
            call()                                                     # well-formed CPU tensors: no CPU path
        with pytest.raises(RuntimeError, match="must be a float tensor"):
            call(a=torch.zeros(2, 8, 3, dtype=torch.float64))          # wrong dtype of the first argument
    for name in ("gather_points", "three_interpolate", "group_points"):
        with pytest.raises(RuntimeError, match="must be an int tensor"):
            calls[name](i=torch.zeros(2, 5, 3, dtype=torch.int64))     # the reference's kernels take int32 indices
    with pytest.raises(RuntimeError, match="must be a contiguous tensor"):
        _ext.furthest_point_sampling(torch.zeros(2, 3, 8).transpose(1, 2), 4)
