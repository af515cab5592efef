/*
 * slide_sap.h -- C ABI of the SAP mesh-reconstruction stage in libslide_b200.so: what the reference
 * runs between its refinement network and marching cubes (pointnet2/dpsr_evaluation.py::visualize_per_rank :176-289 ->
 * network_output_to_dpsr_grid :46-86 -> dpsr_utils/dpsr.py::DPSR.forward :30-77).
 *
 * Conventions as in slide_b200.h: DEVICE pointers, caller-owned buffers and scratch, work enqueued on `stream`,
 * SLIDE_OK or a negative SLIDE_ERR_* code.  The refinement network itself is a slide_program (slide_program.h),
 * lowered from the reference's JSON by slide_b200/sap.py.
 */
#ifndef SLIDE_SAP_H
#define SLIDE_SAP_H

#include "slide_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* mirror_and_concat(partial, axis, num_points=[], attach_label=True, permute)[0]
 * pointnet2/data_utils/mirror_partial.py:8-23 (mirror) and :37-58.
 * cloud f32[B,N,6] (xyz, normal) -> out f32[B,2N,>=7] with row stride ldo floats: rows 0..N-1 of the un-permuted result are
 * the input with label +1, rows N..2N-1 its reflection through the plane through the centroid normal to `axis`
 * ((x - c) negated on that axis, + c; the normal component negated) with label -1; row r of `out` is un-permuted row perm[r]
 * (perm i32[2N], the reference's torch.randperm drawn by the caller; NULL = identity, the permute=False case).
 * centre_scratch: f32[B,3]. */
int slide_sap_mirror_concat(const float *cloud, int B, int N, int axis, const int *perm, float *centre_scratch, float *out,
                            int ldo, slide_stream_t stream);

/* refined_points -> DPSR coordinates                                   pointnet2/dpsr_evaluation.py:22-32, :72-76
 * pts f32[B,n,ld>=3] (first three columns used) -> out f32[B,n,3]:
 *   explicit_normalize != 0: shapenet_psr_normalize (centre of the bounding box, longest side -> 0.99)
 *   explicit_normalize == 0: pts / dataset_scale / 2
 * followed by clamp(x / 1.2 + 0.5, 0, 0.99). */
int slide_sap_unit_cube(const float *pts, int ld, int B, int n, int explicit_normalize, float dataset_scale, float *out,
                        slide_stream_t stream);

/* DPSR(res=(res,res,res), sig, scale, shift).forward(V, N)                pointnet2/dpsr_utils/dpsr.py:10-77
 * (point_rasterize dpsr_utils/utils.py:139-200, spec_gaussian_filter :65-71, fftfreqs :24-46, grid_interp :73-115).
 * V f32[B,n,ldv>=3] in [0,1), Nrm f32[B,n,ldn>=3] -> phi f32[B,res,res,res].  res: a power of two, 8..256.
 * The transforms are this library's own shared-memory FFT passes (no cuFFT): unnormalised forward, 1/res^3 on the way back
 * (torch.fft's default "backward" norm).  workspace: slide_dpsr_workspace_bytes(B, res) bytes of scratch. */
int slide_dpsr_workspace_bytes(int B, int res, size_t *bytes);
int slide_dpsr_forward(const float *V, int ldv, const float *Nrm, int ldn, int B, int n, int res, float sig, int shift,
                       int scale, float *phi, void *workspace, size_t workspace_bytes, slide_stream_t stream);

/* Iso-surface of an indicator grid: replaces measure.marching_cubes(psr_grid[i], level) in mc_from_psr
 * (pointnet2/dpsr_utils/utils.py:246-287; scikit-image's Lewiner marching cubes on the CPU).  The level set is the same; the
 * triangulation is not scikit-image's: cells are split into the six tetrahedra around their main diagonal (marching tetrahedra:
 * no case table, no ambiguous cases, watertight by construction; vertices on grid / face-diagonal / body-diagonal edges at the
 * linear crossing).  Output order is deterministic: vertex i = i-th crossing in (node, edge type) order, faces in (cell,
 * tetrahedron) order, oriented with the normal from phi < level to phi >= level.
 *   phi f32[res,res,res] (one grid), 2 <= res <= 256.
 *   slide_mc_count: fills the workspace (slide_mc_workspace_bytes(res) bytes) and writes counts i32[2] = (n_vertices, n_faces)
 *     -- a DEVICE pointer; read it back to size the outputs;
 *   slide_mc_emit: verts f32[n_vertices,3] = index coordinates * vertex_scale (mc_from_psr divides by res: pass 1/res),
 *     normals f32[n_vertices,3] or NULL = normalised numpy-style gradient of phi interpolated along the edge,
 *     faces i32[n_faces,3]. */
int slide_mc_workspace_bytes(int res, size_t *bytes);
int slide_mc_count(const float *phi, int res, float level, void *workspace, size_t workspace_bytes, int *counts,
                   slide_stream_t stream);
int slide_mc_emit(const float *phi, int res, float level, const void *workspace, float vertex_scale, float *verts,
                  float *normals, int *faces, slide_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
