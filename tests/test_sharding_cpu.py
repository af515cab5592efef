"""Host-side sharding logic on CPU: world-size invariance of the RNG slicing and the output all-gather (gloo, 2 ranks)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slide_b200 import pipeline, weights


def _small_cfg():
    cfg = weights.load_json("pipeline_chair.json")
    cfg["position_ddpm"]["diffusion_config"]["T"] = 5  # keep the CPU draw small; slicing logic is T-independent
    return cfg


def test_noise_slices_do_not_depend_on_world_size():
    cfg = _small_cfg()
    B = 8
    labels = torch.full((B,), cfg["label"], dtype=torch.long)
    torch.manual_seed(123)
    full = pipeline.draw_host_inputs(cfg, B, 0, 1, labels)
    parts = []
    for r in range(2):
        torch.manual_seed(123)
        parts.append(pipeline.draw_host_inputs(cfg, B, r, 2, labels))
    for key in ("pos_xT", "lat_xT", "labels"):
        assert torch.equal(full[key], torch.cat([p[key] for p in parts], dim=0)), key
    for key in ("pos_noise", "starts"):
        assert torch.equal(full[key], torch.cat([p[key] for p in parts], dim=1)), key
    # reference order: x_T first, then one z per step from T-1 down to 1, nothing at t = 0
    torch.manual_seed(123)
    x_T = torch.normal(0, 1, size=(B, 16, 3))
    z = [torch.normal(0, 1, size=(B, 16, 3)) for _ in range(4)]
    assert torch.equal(full["pos_xT"], x_T)
    assert torch.equal(full["pos_noise"][4], z[0]) and torch.equal(full["pos_noise"][1], z[3])
    assert full["pos_noise"][0].abs().max() == 0
    assert int(full["starts"].max()) < 4096 and full["starts"].shape == (3, B)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = torch.full((3, 4, 6), float(rank)) + torch.arange(3).view(3, 1, 1)
    full = pipeline.all_gather_outputs(local, world)
    ok = full.shape == (3 * world, 4, 6) and all(
        torch.equal(full[3 * r:3 * r + 3], torch.full((3, 4, 6), float(r)) + torch.arange(3).view(3, 1, 1))
        for r in range(world))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_all_gather_two_ranks_gloo():
    world = 2
    port = 29500 + (os.getpid() % 500)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0] and out[1]


def _gen_worker(rank, world, port, kp_file, save_dir, out):
    """Two ranks of the generation driver (fake pipelines: no GPU here): each takes its GeneralNpzDataset slice of the
    keypoint file, writes its rank file, rank 0 merges after a barrier -- the reference's evaluate_per_rank +
    gather_generated_results flow (mesh_evaluation.py:15-186)."""
    from slide_b200 import generation
    from tests.test_generation_cpu import _FakePipe
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kp = generation.load_keypoint_file(kp_file, rank=rank, world_size=world)
    generation.generate_per_rank(_FakePipe(2), kp["points"], kp["label"], kp["category"], kp["category_name"],
                                 save_dir=save_dir, rank=rank, world_size=world)
    dist.barrier()
    if rank == 0:
        out["merged"] = generation.gather_generated_results(save_dir, world, num_points=2048)
    dist.barrier()
    dist.destroy_process_group()


def test_generation_driver_two_ranks_gloo(tmp_path):
    import numpy as np
    n = 5
    g = np.random.RandomState(1)
    kp_file = str(tmp_path / "kp.npz")
    pts = g.rand(n, 16, 3).astype(np.float32)
    np.savez(kp_file, points=pts, label=np.arange(n), category=np.array(["c%d" % i for i in range(n)]),
             category_name=np.array(["n"] * n))
    save_dir = str(tmp_path / "out")
    os.makedirs(save_dir)
    port = 29500 + ((os.getpid() + 7) % 500)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gen_worker, args=(2, port, kp_file, save_dir, out), nprocs=2, join=True)
    data = np.load(out["merged"])
    assert os.listdir(save_dir) == ["shapenet_psr_generated_data_2048_pts.npz"]
    assert np.array_equal(data["keypoint"], pts) and list(data["category"]) == ["c%d" % i for i in range(n)]
    assert np.array_equal(data["points"][:, :16], pts)  # cloud i was generated from keypoint set i, ranks in order


def test_external_keypoint_draws_follow_reference_rng_order():
    """latent_ddpm_keypoint_conditional_generation never draws position noise: its first CPU draw is the latent x_T
    (diffusion_utils/diffusion.py:373), then the decoder's FPS start indices.  skip_position reproduces exactly that
    generator sequence (the unused position draws used to advance the generator)."""
    cfg = weights.load_json("pipeline_airplane.json")
    B = 4
    labels = torch.zeros(B, dtype=torch.long)
    torch.manual_seed(123)
    d = pipeline.draw_host_inputs(cfg, B, 0, 1, labels, skip_position=True)
    torch.manual_seed(123)
    want_xT = torch.randn(B, 16, 3 + cfg["latent_ddpm"]["pointnet_config"]["in_fea_dim"])
    assert torch.equal(d["lat_xT"], want_xT)
    assert d["pos_xT"] is None and d["pos_noise"] is None
    n_in = 16
    for lvl, dcfg in enumerate(cfg["autoencoder"]["decoders"]):
        up = dcfg["upsampling_setting"]
        P = n_in * up["point_upsample_factor"]
        if P > up["num_output_points"]:
            want = torch.tensor([int(torch.randint(high=P, size=(1,)).item()) for _ in range(B)])
            assert torch.equal(d["starts"][lvl], want)
        n_in = up["num_output_points"]
