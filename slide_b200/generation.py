"""Callers of the sampling hot path and their on-disk formats.

Mirrors, for the keypoint-conditional generation flow only,
  pointnet2/mesh_evaluation.py::evaluate_per_rank          (:15-153)  -> generate_per_rank
  pointnet2/mesh_evaluation.py::gather_generated_results   (:156-186) -> gather_generated_results
  sampling_and_inference/latent_ddpm_keypoint_conditional_generation.py (:158-177) -> load_keypoint_file
with the same file names and npz keys, so that files written by either side load in the other:
  keypoint file (input) : points (B,16,3) [, label (B,), category, category_name, keypoint_feature (B,16,F),
                          keypoint_mask (B,16)]
  result file (output)  : points (B,P,3), normals (B,P,3), label (B,), category, category_name, timing (B,),
                          keypoint (B,16,3) [, keypoint_feature (B,16,F)] [, gt_points (B,N,6) when the keypoints were
                          sampled from known shapes, mesh_evaluation.py:86-98,140-142]
What changes is where the time goes: position -> feature -> decode are chained on the device (no npz round trip
between the reference's two scripts), a batch is copied to the host once, and under torch.distributed the per-rank
clouds travel through one NCCL all-gather instead of per-rank files (the file-level gather is kept for drop-in use).
"""
import os
import time

import numpy as np
import torch

RESULT_NAME = "shapenet_psr_generated_data_%d_pts%s.npz"            # mesh_evaluation.py:39
RANK_RESULT_NAME = "shapenet_psr_generated_data_%d_pts_rank_%d%s.npz"  # mesh_evaluation.py:41


def load_keypoint_file(path, rank=0, world_size=1, local_resampling=False):
    """The keypoint file of latent_ddpm_keypoint_conditional_generation.py (--keypoint_file), sliced for this rank the
    way GeneralNpzDataset does (contiguous chunks of ceil(B / world_size), npz_dataset.py:90-98)."""
    data = np.load(path, allow_pickle=True)
    B = data["points"].shape[0]
    per = -(-B // world_size)
    sl = slice(rank * per, min((rank + 1) * per, B))
    out = {"points": torch.from_numpy(np.asarray(data["points"][sl], dtype=np.float32))}
    n = out["points"].shape[0]
    out["label"] = torch.from_numpy(np.asarray(data["label"][sl]).astype(np.int64)) if "label" in data.files \
        else torch.zeros(n, dtype=torch.int64)
    for key in ("category", "category_name"):
        out[key] = [str(v) for v in data[key][sl]] if key in data.files else [""] * n
    if local_resampling:
        # latent_ddpm_keypoint_conditional_generation.py:160-165
        feat = torch.from_numpy(np.asarray(data["keypoint_feature"][sl], dtype=np.float32))
        out["keypoint_mask"] = torch.from_numpy(np.asarray(data["keypoint_mask"][sl], dtype=np.float32))
        out["complete_x0"] = torch.cat([out["points"], feat], dim=2)
    return out


def result_file(save_dir, num_points, rank=0, world_size=1, ckpt_info=""):
    if world_size == 1:
        return os.path.join(save_dir, RESULT_NAME % (num_points, ckpt_info))
    return os.path.join(save_dir, RANK_RESULT_NAME % (num_points, rank, ckpt_info))


def pack_results(clouds, label, category, category_name, timing, keypoint=None, keypoint_feature=None,
                 split_points_and_normals=True, gt_points=None):
    """The result dict of evaluate_per_rank (mesh_evaluation.py:135-150)."""
    clouds = np.asarray(clouds, dtype=np.float32)
    result = {"points": clouds, "label": np.asarray(label), "category": list(category),
              "category_name": list(category_name), "timing": np.asarray(timing)}
    if keypoint is not None:
        result["keypoint"] = np.asarray(keypoint, dtype=np.float32)
    if gt_points is not None:
        result["gt_points"] = np.asarray(gt_points, dtype=np.float32)
    if keypoint_feature is not None:
        result["keypoint_feature"] = np.asarray(keypoint_feature, dtype=np.float32)
    if split_points_and_normals and clouds.shape[2] == 6:
        result["normals"] = clouds[:, :, 3:]
        result["points"] = clouds[:, :, 0:3]
    return result


def generate_per_rank(pipe, keypoints, label, category=None, category_name=None, save_dir=None, ckpt_info="",
                      save_keypoint_feature=False, complete_x0=None, keypoint_mask=None,
                      split_points_and_normals=True, rank=0, world_size=1, gt_points=None, keypoint_noise_magnitude=0.0):
    """evaluate_per_rank for task 'latent_keypoint_conditional_generation' with external keypoints: feature DDPM +
    decode for this rank's keypoints (any count: the tail batch is padded up to the pipeline's batch and trimmed).

    pipe: this process's SlidePipeline over its own keypoint slice (world=1; built with local_resampling=True when
    complete_x0 / keypoint_mask are given); rank / world_size only select the result file name, as in the reference
    where every rank samples its slice independently.  keypoints (n,16,3), label (n,) CPU tensors.
    gt_points (n,N,6): the shapes the keypoints were sampled from (the reference's `test_external_keypoint=False` branch,
    mesh_evaluation.py:86-98; `keypoints_from_shapes` below makes such keypoints); stored under `gt_points` as the
    reference does (:140-142).  keypoint_noise_magnitude > 0: `keypoint + magnitude * torch.randn_like(keypoint)` per
    batch on the device generator before sampling (:80-82,92-94); the noised keypoints are conditioned on and saved.
    Returns the result dict; writes it to the reference's file name when save_dir is given."""
    assert pipe.world == 1, "one independent pipeline per rank (see load_keypoint_file for the slicing)"
    n = keypoints.shape[0]
    Bl = pipe.Bl
    clouds, feats, timing, used_kp = [], [], [], []
    for b0 in range(0, n, Bl):
        m = min(Bl, n - b0)
        pad = lambda t: torch.cat([t[b0:b0 + m], t[b0:b0 + 1].expand((Bl - m,) + tuple(t.shape[1:]))]) if m < Bl \
            else t[b0:b0 + m]
        start = time.time()
        # the host draws of one batch, in the reference's order (x_T, then the decoder's FPS start indices); the
        # reference never draws position-DDPM noise on this path (diffusion.py:373 is its first draw), so neither do we
        pipe.draw_host_inputs(pad(label), skip_position=True)
        pipe.stage_inputs()
        kp_dev = keypoints[b0:b0 + m].to(pipe.device)
        if keypoint_noise_magnitude > 0:
            kp_dev = kp_dev + keypoint_noise_magnitude * torch.randn_like(kp_dev)  # the batch's real rows only
        used_kp.append(kp_dev.cpu())
        if m < Bl:
            kp_dev = torch.cat([kp_dev, kp_dev[0:1].expand(Bl - m, -1, -1)])
        out = pipe.sample_resident(keypoints=kp_dev,
                                   complete_x0=None if complete_x0 is None else pad(complete_x0),
                                   keypoint_mask=None if keypoint_mask is None else pad(keypoint_mask))
        host = out[:m].cpu()
        pipe.check_device_errors()
        timing.extend([(time.time() - start) / m] * m)
        clouds.append(host.numpy())
        feats.append(pipe.keypoint_feature[:m].cpu().numpy())
    result = pack_results(np.concatenate(clouds, axis=0), label.numpy(), category or [""] * n,
                          category_name or [""] * n, timing, keypoint=torch.cat(used_kp).numpy(),
                          gt_points=None if gt_points is None else np.asarray(gt_points),
                          keypoint_feature=np.concatenate(feats, axis=0) if save_keypoint_feature else None,
                          split_points_and_normals=split_points_and_normals)
    if save_dir is not None:
        os.makedirs(save_dir, exist_ok=True)
        np.savez(result_file(save_dir, pipe.dec.out_points, rank, world_size, ckpt_info), **result)
    return result


def keypoints_from_shapes(points, num_keypoints=16, add_centroid=True):
    """The keypoints of known shapes, as evaluate_per_rank makes them when ground truth is given
    (mesh_evaluation.py:88-91 -> data_utils/points_sampling.py:156-187, random_subsample=False): farthest point sampling
    over [centroid; points] from index 0 (add_centroid=True: the centroid is the first keypoint), or over the points from
    a random start index per shape (pytorch3d's CPU-generator draws, as in the reference: random_start_point=not
    add_centroid; the 8- / 32-keypoint ablation configs set add_centroid_to_keypoints false).
    points (B,N,3) on the GPU -> (B,K,3)."""
    from .pipeline import sample_keypoints
    if add_centroid:
        return sample_keypoints(points, K=num_keypoints)
    from . import install_dropin
    install_dropin()
    from pytorch3d.ops import sample_farthest_points
    return sample_farthest_points(points.contiguous(), K=num_keypoints, random_start_point=True)[0]


def gather_generated_results(save_dir, world_size, num_points=2048, ckpt_info="", remove_rank_files=True):
    """mesh_evaluation.py:156-186: concatenate the per-rank result files key by key into the single-process file name
    and delete the rank files."""
    merged = {}
    files = []
    for rank in range(world_size):
        f = os.path.join(save_dir, RANK_RESULT_NAME % (num_points, rank, ckpt_info))
        data = np.load(f, allow_pickle=True)
        for key in data.files:
            merged.setdefault(key, []).append(data[key])
        files.append(f)
    merged = {k: np.concatenate(v, axis=0) for k, v in merged.items()}
    out = os.path.join(save_dir, RESULT_NAME % (num_points, ckpt_info))
    np.savez(out, **merged)
    if remove_rank_files:
        for f in files:
            os.remove(f)
    return out


def gather_on_device(pipe, local_out):
    """The B200-native replacement of the file-level gather: one NCCL all-gather of the (B/W, P, 6) clouds."""
    from .pipeline import all_gather_outputs
    return all_gather_outputs(local_out, pipe.world)
