/*
 * slide_b200.h -- C ABI of libslide_b200.so: SLIDE's diffusion-sampling + autoencoder-decode hot path
 * as hand-written sm_100a CUDA.
 *
 * Conventions (every entry point):
 *   - plain pointers and sizes, no torch types; all pointers are DEVICE pointers unless stated otherwise;
 *   - the caller owns every buffer, including scratch; nothing is allocated behind the caller's back
 *     except inside a slide_program (created/destroyed explicitly);
 *   - work is enqueued on `stream` (a cudaStream_t), no implicit synchronisation;
 *   - returns SLIDE_OK (0) or a negative SLIDE_ERR_* code.  The reference prints and calls exit(-1) on a
 *     launch error (pointnet2_ops/_ext-src/include/cuda_utils.h:30-39); this library never exits.
 *
 * Each function cites the reference interface it replaces (paths relative to the reference repo;
 * EXT = pointnet2_ops_lib/pointnet2_ops/_ext-src).
 */
#ifndef SLIDE_B200_H
#define SLIDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *slide_stream_t; /* cudaStream_t */

#define SLIDE_OK 0
#define SLIDE_ERR_INVALID (-1)     /* bad argument (shape, NULL pointer, unsupported size) */
#define SLIDE_ERR_CUDA (-2)        /* a CUDA runtime call or launch failed; see slide_last_cuda_error() */
#define SLIDE_ERR_UNSUPPORTED (-3) /* valid request outside what the sm_100a kernels implement */

/* Version / diagnostics */
int slide_abi_version(void);
const char *slide_last_cuda_error(void);
/* Number of kernels this library has launched since load (all entry points); used by bench.py's
 * gpu_launches.  slide_reset_launch_count() zeroes it. */
long long slide_launch_count(void);
void slide_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Point-set index ops: the nine functions of the pybind module `pointnet2_ops._ext`
 * (EXT/src/bindings.cpp:6-19).
 * ------------------------------------------------------------------------------------------------ */

/* furthest_point_sampling(points f32[B,N,3], nsamples) -> i32[B,m]      EXT/src/sampling.cpp:66-87,
 * kernel EXT/src/sampling_gpu.cu:69-229.  Start index 0; points with |p|^2 <= 1e-3 are never picked;
 * running min distance starts at 1e10; ties resolve exactly like the reference's strided scan + shared
 * memory tree for the block size the reference launcher would use (EXT/include/cuda_utils.h:15-19).
 * No scratch needed: the running distances live in registers. */
int slide_furthest_point_sampling(const float *xyz, int B, int N, int m, int *idx, slide_stream_t stream);

/* The same op for clouds of any size: N <= slide_fps_resident_max_points() (16384) runs the register-resident
 * kernel above and ignores `temp`; larger clouds keep their running distances in `temp`, a caller-provided scratch
 * f32[B,N] -- the reference's own `tmp` tensor (EXT/src/sampling.cpp:74-76; contents need not be initialised). */
int slide_furthest_point_sampling_ws(const float *xyz, int B, int N, int m, int *idx, float *temp,
                                     slide_stream_t stream);
int slide_fps_resident_max_points(void);

/* gather_points(points f32[B,C,N], idx i32[B,m]) -> f32[B,C,m]          EXT/src/sampling.cpp:15-39 */
int slide_gather_points(const float *points, const int *idx, int B, int C, int N, int m, float *out,
                        slide_stream_t stream);
/* gather_points_grad(grad_out f32[B,C,m], idx, n) -> f32[B,C,n]         EXT/src/sampling.cpp:41-65
 * grad_points must be zero-filled by the caller (the reference allocates it with torch::zeros). */
int slide_gather_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int m,
                             float *grad_points, slide_stream_t stream);

/* ball_query(new_xyz f32[B,m,3], xyz f32[B,N,3], radius, nsample) -> (idx i32[B,m,ns], counts i32[B,m])
 * EXT/src/ball_query.cpp:10-38, kernel EXT/src/ball_query_gpu.cu:9-57.  First `nsample` hits in ascending
 * point order with d^2 < r^2 (strict), padded with the first hit; rows without a hit are all 0, count 0.
 * Both outputs are fully written (no pre-zeroing needed). */
int slide_ball_query(const float *new_xyz, const float *xyz, int B, int N, int m, float radius, int nsample,
                     int *idx, int *counts, slide_stream_t stream);

/* group_points(points f32[B,C,N], idx i32[B,np,ns]) -> f32[B,C,np,ns]   EXT/src/group_points.cpp:13-40 */
int slide_group_points(const float *points, const int *idx, int B, int C, int N, int npoint, int nsample,
                       float *out, slide_stream_t stream);
/* group_points_grad(grad_out f32[B,C,np,ns], idx, n) -> f32[B,C,n]; grad_points pre-zeroed by caller. */
int slide_group_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int npoint,
                            int nsample, float *grad_points, slide_stream_t stream);

/* three_nn(unknown f32[B,n,3], known f32[B,m,3]) -> (dist2 f32[B,n,3], idx i32[B,n,3])
 * EXT/src/interpolate.cpp:15-43, kernel EXT/src/interpolate_gpu.cu:9-68.  Squared distances, strict-<
 * insertion in ascending known index; with m < 3 the missing slots are +inf / index 0 like the reference. */
int slide_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2, int *idx,
                   slide_stream_t stream);
/* three_interpolate(points f32[B,C,m], idx i32[B,n,3], weight f32[B,n,3]) -> f32[B,C,n] */
int slide_three_interpolate(const float *points, const int *idx, const float *weight, int B, int C, int m,
                            int n, float *out, slide_stream_t stream);
/* three_interpolate_grad(grad_out f32[B,C,n], idx, weight, m) -> f32[B,C,m]; pre-zeroed by caller. */
int slide_three_interpolate_grad(const float *grad_out, const int *idx, const float *weight, int B, int C,
                                 int n, int m, float *grad_points, slide_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * pytorch3d 0.7.0 ops used on the path (the reference depends on them, environment.yml:117; call sites
 * pointnet2_ops/pointnet2_utils.py:370,506-507 and models/point_upsample_decoder.py:178-180).
 * ------------------------------------------------------------------------------------------------ */

/* knn_points(p1 f32[B,P1,3], p2 f32[B,P2,3], lengths1?, lengths2?, K) -> (dists f32[B,P1,K] squared,
 * idx i64[B,P1,K]) sorted ascending, equal distances in ascending index.  lengths* may be NULL (i64[B]).
 * Slots beyond lengths2 and rows beyond lengths1 are written as 0.  D must be 3, 1 <= K <= 64. */
int slide_knn_points(const float *p1, const float *p2, int B, int P1, int P2, int D, const int64_t *lengths1,
                     const int64_t *lengths2, int K, float *dists, int64_t *idx, slide_stream_t stream);

/* sample_farthest_points(points f32[B,P,3], lengths? i64[B], K? i64[B] (or maxK for all), start_idx? i64[B])
 * -> idx i64[B,maxK], -1 padded beyond min(K_b, length_b).  Running distances start at +inf, first pick
 * is start_idx[b] (0 if NULL), every next pick is the lowest-index arg-max. */
int slide_sample_farthest_points(const float *points, int B, int P, int D, const int64_t *lengths,
                                 const int64_t *K, const int64_t *start_idx, int maxK, int64_t *idx,
                                 slide_stream_t stream);

/* As above for P > slide_fps_resident_max_points(): `temp` is a caller-provided scratch f32[B,P]. */
int slide_sample_farthest_points_ws(const float *points, int B, int P, int D, const int64_t *lengths,
                                    const int64_t *K, const int64_t *start_idx, int maxK, int64_t *idx, float *temp,
                                    slide_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused network programs: the denoiser forward (models/pointnet2_with_pcld_condition.py:286-489), the
 * DDPM steps (util.py:197-259, diffusion_utils/diffusion.py:58-95,346-404) and the autoencoder decode
 * (models/autoencoder.py:42-45) are compiled by the host side into a flat list of slide_op records that
 * this library executes as CUDA kernels on one stream (optionally captured once into a CUDA graph and
 * replayed per diffusion step).  The record layout and op kinds are in slide_program.h.
 * ------------------------------------------------------------------------------------------------ */
typedef struct slide_program slide_program;
struct slide_op;

/* Create a program.  `ops` is a HOST array that is copied.  The program owns one device arena of
 * `arena_bytes` bytes (zero-filled) and one of `weights_bytes` bytes filled from the HOST buffer
 * `weights` (may be NULL when weights_bytes == 0). */
int slide_program_create(const struct slide_op *ops, int n_ops, size_t arena_bytes, const void *weights,
                         size_t weights_bytes, slide_program **out);
void slide_program_destroy(slide_program *p);
/* Device base pointers of the two arenas (for the host side to copy inputs in / outputs out). */
void *slide_program_arena(slide_program *p);
void *slide_program_weights(slide_program *p);
/* Enqueue ops [first, first+count) on `stream`. */
int slide_program_run(slide_program *p, int first, int count, slide_stream_t stream);
/* Capture `repeat` back-to-back passes over ops [first, first+count) into one CUDA graph (slot 0..7); replay
 * launches that graph `times` times (so one replay = `repeat` diffusion steps). */
int slide_program_capture(slide_program *p, int slot, int first, int count, int repeat, slide_stream_t stream);
int slide_program_replay(slide_program *p, int slot, int times, slide_stream_t stream);
/* GEMM kernel selection: 0 = tcgen05 (TF32 operands, fp32 accumulate in TMEM) where the contraction is dense
 * and aligned, fp32 FFMA otherwise; 1 = fp32 FFMA everywhere.  Env SLIDE_GEMM_BACKEND=simt sets 1 at creation. */
int slide_program_set_gemm_backend(slide_program *p, int backend);
/* Non-zero if a tcgen05 pipeline wait ever timed out in this process (a bug guard; results are then invalid). */
int slide_tc_error(void);
/* Clear that flag (after the caller has discarded the affected results). */
void slide_tc_reset_error(void);
/* Kernel-selection knobs (SLIDE_TC_* / SLIDE_PAIR_* environment variables, for A/B runs and tests) are read once, on
 * first use; this re-reads them. */
void slide_tc_reload_tuning(void);
/* Sample-resident execution of a record range (slide_resident.h): `plan` / `rops` are HOST arrays produced by the host
 * side's compiler (slide_b200/resident.py) for ops [plan->first, plan->first + plan->count); they are copied.  From then
 * on slide_program_run / _capture over exactly that range launch ONE kernel (one thread-block cluster per sample,
 * activations resident in shared memory) instead of one kernel per record -- while the GEMM backend is 0 and resident
 * execution is enabled (slide_program_use_resident, env SLIDE_RESIDENT=0 disables at creation).  The packed weight
 * copies the plan refers to must already be part of the program's weight blob. */
struct slide_resident_plan;
struct slide_rop;
int slide_program_set_resident(slide_program *p, const struct slide_resident_plan *plan, const struct slide_rop *rops,
                               int n_rops);
int slide_program_use_resident(slide_program *p, int enable);
/* Slice of torch's CUDA normal_() stream (replaces the T full-batch torch.randn_like calls of the feature DDPM's loop,
 * pointnet2/diffusion_utils/diffusion.py:88, on a rank that owns only part of the batch).  For call s = 0..n_calls-1 of
 * normal_() on a contiguous fp32 tensor of `numel` elements -- Philox4_32_10 (seed, offset + s * offset_increment), ATen's
 * grid-stride mapping with `grid_full` = min(sm_count * (max_threads_per_sm / 256), ceil(numel / 256)) blocks of 256
 * threads and offset_increment = ((numel - 1) / (1024 * grid_full) + 1) * 4 -- writes elements
 * [slice_begin, slice_begin + slice_len) to out[row * out_call_stride + 0..slice_len), row = s (reverse = 0) or
 * n_calls - 1 - s (reverse = 1), bit-identical to what the full draw puts there.  One launch. */
int slide_philox_normal_slice(float *out, long long out_call_stride, int n_calls, int reverse, unsigned long long seed,
                              unsigned long long offset, unsigned long long offset_increment, long long numel,
                              long long slice_begin, long long slice_len, int grid_full, slide_stream_t stream);
/* Kernels launched by one pass over ops [first, first+count). */
int slide_program_launches(slide_program *p, int first, int count);

#ifdef __cplusplus
}
#endif
#endif /* SLIDE_B200_H */
