/*
 * slide_program.h -- record layout of the fused network programs executed by libslide_b200.so
 * (slide_program_create / _run / _capture / _replay in slide_b200.h).
 *
 * A program is a flat array of `slide_op` records.  The host side (slide_b200/nets.py) lowers the
 * reference's modules -- PointNet2CloudCondition.forward (pointnet2/models/pointnet2_with_pcld_condition.py:
 * 286-489), PointnetSAModule / PointnetKnnFPModule / FeatureMapModule / Mlp_plus_t_emb / AttentionModule
 * (pointnet2_ops_lib/pointnet2_ops/pointnet2_modules.py:119-176,212-292,591-663,771-873, attention.py:35-96),
 * the samplers (pointnet2/util.py:197-259, diffusion_utils/diffusion.py:58-95,346-404) and the decoder
 * (models/autoencoder.py:42-45) -- into these records; this library turns each record into one kernel.
 *
 * Memory model.  Two device arenas: ARENA (activations, indices, statistics, noise; zero-filled at creation)
 * and WEIGHTS (read-only parameters).  Every pointer field is a BYTE OFFSET into one of them, -1 = absent.
 * Fields whose name ends in _W address WEIGHTS, all others address ARENA.
 *
 * Activations are fp32 row-major matrices [rows, ld] with CHANNELS LAST: row = (sample, point[, neighbour]),
 * column = channel.  (The reference is channel-major (B,C,np,K); channels-last turns every grouping gather
 * into a contiguous row copy and every 1x1 conv into a K-major GEMM operand.)
 *
 * GroupNorm is never a kernel of its own.  The GEMM that produces a tensor accumulates per-(sample, group)
 * sum / sum-of-squares of its outputs into a statistics buffer (`ST_*` fields, fp64 atomics); whoever
 * consumes the tensor applies normalisation + affine + ReLU + the additive timestep / condition vector while
 * loading it (`XF_*` fields).  MyGroupNorm's rule "normalise the leading floor(C/G)*G channels, pass the
 * rest" (pointnet2_modules.py:24-42) is XF_NNORM; GroupNorm eps is SLIDE_GN_EPS.
 */
#ifndef SLIDE_PROGRAM_H
#define SLIDE_PROGRAM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLIDE_OP_NPARAM 72
#define SLIDE_OP_NFPARAM 4
#define SLIDE_GN_EPS 1e-5f

struct slide_op {
  int32_t kind;  /* enum slide_op_kind */
  int32_t flags; /* SLIDE_OPF_* */
  int64_t p[SLIDE_OP_NPARAM];
  float f[SLIDE_OP_NFPARAM];
};

/* Record flags.  SIDE: the record belongs to the side branch of a two-branch region (AttentionModule's key/query
 * branch runs next to the shared MLP, they only meet at the soft-max).  The executor enqueues SIDE records on a
 * second stream that first waits for everything enqueued before the region's first SIDE record; a SLIDE_OP_JOIN
 * record makes the main stream wait for the side branch.  Records are stored in an order that is also a valid
 * sequential order, so an executor may ignore the flag. */
#define SLIDE_OPF_SIDE 1

enum slide_op_kind {
  SLIDE_OP_NOP = 0,
  SLIDE_OP_STEP_BEGIN = 1,   /* zero a statistics region, step counter -= 1 */
  SLIDE_OP_KNN = 2,          /* K nearest neighbours (pytorch3d knn_points semantics), i32 indices + squared dists */
  SLIDE_OP_GROUP = 3,        /* build grouped rows [f_j | geometry] (QueryAndGroup 'nn' / group_knn) */
  SLIDE_OP_GEMM = 4,         /* C = act(xf(A) W^T + bias + addvec + xf(resid)), statistics of C */
  SLIDE_OP_SOFTMAX_WSUM = 5, /* out[i,c] = sum_k xf(V)[i,k,c] * softmax_k(S[i,k,c])  (AttentionModule tail; the
                                lowering normally fuses this into the score GEMM, see GEMM_SMK) */
  SLIDE_OP_COPY_COLS = 6,    /* dst[:, 0:n] = src[:, 0:n] */
  SLIDE_OP_DDPM_UPDATE = 7,  /* one ancestral sampling update (position or latent flavour) */
  SLIDE_OP_FPS = 8,          /* furthest point sampling (pointnet2_ops._ext or pytorch3d semantics) */
  SLIDE_OP_GATHER_ROWS = 9,  /* dst[(s,j), :] = src[(s, idx[s,j]), :] */
  SLIDE_OP_UPSAMPLE = 10,    /* point_upsample: children = coarse + displacement * scale / sqrt(factor) */
  SLIDE_OP_TEMB = 11,        /* sinusoidal timestep embedding (calc_t_emb) */
  SLIDE_OP_COLMAX = 12,      /* out[s,c] = max_r xf(X)[s*R + r, c]  (Pnet2Stage's global max-pool) */
  SLIDE_OP_KL = 13,          /* DiagonalGaussianDistribution: mode, or mean + exp(0.5*clamp(logvar)) * noise */
  SLIDE_OP_JOIN = 14,        /* main branch waits for the side branch (no kernel) */
  SLIDE_OP_PAIR = 15,        /* a 1x1 conv over grouped rows, factored through the gather ("conv before gather") */
  SLIDE_OP_KIND_COUNT
};

/* Transform-on-load block (XF): 12 consecutive params starting at a base index.
 *   y = x                                                   if STATS < 0
 *   y = (x - mean[s,g]) * rstd[s,g] * gamma[ch] + beta[ch]   if ch = CHOFF + col < NNORM, g = ch / CG
 *   y = max(y, 0)                                            if RELU
 *   y += addvec[row, col]                                    if ADDVEC >= 0
 * with s = row / R, mean/rstd from the fp64 sums at STATS ([B, NNORM/CG, 2]) and COUNT elements per group.
 * addvec row: ADDMODE 0 -> s (per sample), 1 -> the step counter (timestep table), 2 -> row 0. */
enum slide_xf_field {
  XF_STATS = 0,
  XF_CG,
  XF_NNORM,
  XF_CHOFF,
  XF_GAMMA_W,
  XF_BETA_W,
  XF_R,
  XF_COUNT,
  XF_RELU,
  XF_ADDVEC,
  XF_ADDLD,
  XF_ADDMODE,
  XF_NFIELD
};

enum slide_step_begin_field { SB_ZERO_OFF = 0, SB_ZERO_BYTES, SB_STEP };

/* idx i32 [B,P1,K] (ascending distance, ties in ascending index), d2 f32 [B,P1,K] or absent */
enum slide_knn_field { KNN_Q = 0, KNN_LDQ, KNN_P1, KNN_REF, KNN_LDR, KNN_P2, KNN_K, KNN_IDX, KNN_D2, KNN_B };

/* MODE 0 (QueryAndGroup):  [f_j (C) | x_j - c_i | x_j (if ABS) | c_i (if CENTER)]
 * MODE 1 (group_knn):      [f_j (C) | d2 | w | x_j | x_j - c_i | c_i],  w = (1/(d2+1e-8)) / sum_k(1/(d2+1e-8)) */
enum slide_group_field {
  GRP_MODE = 0, GRP_F, GRP_LDF, GRP_C, GRP_XYZ, GRP_LDX, GRP_N, GRP_CTR, GRP_LDCTR, GRP_NP, GRP_IDX, GRP_K,
  GRP_D2, GRP_OUT, GRP_LDO, GRP_ABS, GRP_CENTER, GRP_B
};

/* C[M, N] (row stride LDC) = act( xfA(A)[M, K] * W[N, K]^T + bias[N] + ev[row / EV_DIV, N] + xfR(resid)[M, N] )
 * W is fp32 row-major [N, LDW] in WEIGHTS.  ACT: 0 none, 1 relu, 2 swish.
 * Statistics of the stored value v (after ACT): sum += ST_WEIGHT * v, sumsq += ST_WEIGHT * v * v into
 * fp64 [B, ST_NNORM / ST_CG, 2] at ST_STATS, for columns with ST_CHOFF + col < ST_NNORM, s = row / ST_R. */
enum slide_gemm_field {
  GEMM_A = 0, GEMM_LDA, GEMM_M, GEMM_K, GEMM_W_W, GEMM_LDW, GEMM_N, GEMM_C, GEMM_LDC, GEMM_BIAS_W, GEMM_ACT,
  GEMM_EV, GEMM_EVLD, GEMM_EVDIV,
  GEMM_RES, GEMM_LDR,
  GEMM_ST_STATS, GEMM_ST_CG, GEMM_ST_NNORM, GEMM_ST_CHOFF, GEMM_ST_R, GEMM_ST_WEIGHT,
  GEMM_XFA, /* XF block for A */
  GEMM_XFR = GEMM_XFA + XF_NFIELD, /* XF block for resid */
  GEMM_STEP = GEMM_XFR + XF_NFIELD, /* step counter (for XF_ADDMODE 1) */
  GEMM_WP_W,  /* tensor-core copy of W (or -1): TF32-rounded, tiled [ceil(K/32)][WP_NA][8 rows][128 B], each
                 8x128 B atom in the SWIZZLE_128B pattern, zero padded -- one bulk copy per (N tile, K block) */
  GEMM_WP_NA, /* 8-row atoms per K block in that copy (N rounded up to a multiple of 256, / 8) */
  GEMM_SMK,   /* > 0: fused AttentionModule tail.  The GEMM result S = xfA(A) W^T + bias is a score tensor; soft-max
                 is taken over every group of SMK consecutive rows (the neighbours of one point) and applied to the
                 VALUE tensor given by RES / XFR:  C[g, n] = sum_k xfR(RES)[g*SMK + k, n] * softmax_k(S[g*SMK + k, n]).
                 C then has M / SMK rows; EV, ACT and ST_* must be unset. */
  GEMM_NFIELD
};

/* S, V: [ROWS*K, C]; out: [ROWS, C] (row stride LDO) */
enum slide_softmax_field {
  SM_S = 0, SM_LDS, SM_V, SM_LDV, SM_OUT, SM_LDO, SM_ROWS, SM_K, SM_C, SM_STEP,
  SM_XFV, /* XF block for V */
  SM_NFIELD = SM_XFV + XF_NFIELD
};

enum slide_copy_field { CP_SRC = 0, CP_LDS, CP_DST, CP_LDD, CP_ROWS, CP_NCOLS };

/* MODE 0 (pointnet2/util.py:240-253):  x = (x - k1*eps) / sqrt_alpha ; if t > 0: x += sigma * noise
 * MODE 1 (diffusion_utils/diffusion.py:68-92): x0 = c1*x - c2*eps ; [clamp] ;
 *         [local resampling, :76-79: x0 = x0 * mask + X0C * (1 - mask)] ; mean = pm1*x0 + pm2*x ;
 *         x = mean + (t != 0) * sig * noise
 * MODE 2 (pointnet2/util_fastdpmv2.py:436-443, FastDPM VAR / STEP samplers): x = x * a + (c * eps + sigma * noise)
 *         with TABLE_W row = [a, c, sigma] (row index = step counter, see engine.fast_position_schedule)
 * Only columns [COL0, NCOLS) of x are written (keypoint-conditional sampling keeps the xyz columns).
 * TABLE_W: f32 [T, 8] per-timestep coefficients; NOISE: f32 [T, ROWS, NCOLS] (row stride NCOLS).
 * X0C (-1 = no local resampling): f32 [ROWS, NCOLS] (row stride LDX0C), the complete x0 whose features are kept where
 * MASK f32 [ROWS] is 0 and re-sampled where it is 1. */
enum slide_ddpm_field {
  DD_MODE = 0, DD_X, DD_LDX, DD_EPS, DD_LDE, DD_NOISE, DD_ROWS, DD_NCOLS, DD_COL0, DD_TABLE_W, DD_STEP,
  DD_X0C, DD_LDX0C, DD_MASK
};

/* MODE 0: pointnet2_ops._ext (start 0, |p|^2 <= 1e-3 skipped, i32 out); MODE 1: pytorch3d (start index from
 * START i32 [B] or 0, i32 out) */
enum slide_fps_field { FPS_MODE = 0, FPS_XYZ, FPS_LDX, FPS_N, FPS_M, FPS_OUT, FPS_START, FPS_B };

enum slide_gather_field { GA_SRC = 0, GA_LDS, GA_N, GA_IDX, GA_M, GA_DST, GA_LDD, GA_NCOLS, GA_B };

/* out[(s, n*FACTOR + q), c] = coarse[(s,n), c] (0 beyond COARSE_C) + (disp[(s,n), q*F + c] * f[0]) * f[1]
 * with f[0] = 1/sqrt(FACTOR), f[1] = output scale: two fp32 multiplies then the add, in the reference's order
 * (pointnet2/models/point_upsample_module.py:18-46). */
enum slide_upsample_field {
  UP_COARSE = 0, UP_LDC, UP_COARSE_C, UP_DISP, UP_LDD, UP_OUT, UP_LDO, UP_ROWS, UP_FACTOR, UP_F
};

/* out[i, 0:half] = sin(ts[i] * freq[j]), out[i, half:2*half] = cos(...); ts f32 [ROWS], FREQ_W f32 [half] */
enum slide_temb_field { TE_TS = 0, TE_FREQ_W, TE_HALF, TE_OUT, TE_LDO, TE_ROWS };

/* A 1x1 conv W over the grouped row of pair (i, j) -- QueryAndGroup's [f_j | x_j - c_i | x_j | c_i] or group_knn's
 * [f_j | d2 | w | x_j | x_j - c_i | c_i] -- is linear in its parts, so the grouped tensor is never built:
 *   out[(s,i,k), n] = act( U[(s, j), n] + x_j . WX[n, 0:3] + c_i . WC[n, 0:3] + bias[n]
 *                          + d2_ik * WD[n] + w_ik * WW[n] + xfR(RES)[(s,i,k), n] ),      j = IDX[s,i,k]
 * where U = f W_f^T comes from ONE GEMM over the NSRC source points (not over the NP*K pairs), WX = W_abs + W_rel,
 * WC = W_ctr - W_rel (host-combined, [N,3] row-major in WEIGHTS), WD / WW the d2 / w columns (group_knn only, else -1),
 * w_ik = (1/(d2_ik+1e-8)) / sum_k(1/(d2_ik+1e-8)).  Statistics of out as for GEMM (ST_* fields). */
enum slide_pair_field {
  PR_U = 0, PR_LDU, PR_NSRC, PR_XYZ, PR_LDX, PR_CTR, PR_LDCTR, PR_NP, PR_IDX, PR_K, PR_D2,
  PR_WX_W, PR_WC_W, PR_WD_W, PR_WW_W, PR_BIAS_W, PR_N, PR_OUT, PR_LDO, PR_ACT, PR_RES, PR_LDR,
  PR_ST_STATS, PR_ST_CG, PR_ST_NNORM, PR_ST_CHOFF, PR_ST_WEIGHT, PR_B, PR_STEP,
  PR_XFR, PR_NFIELD = PR_XFR + XF_NFIELD
};

/* out f32 [B, C] (row stride LDO); X f32 [B*R, C]; XF block = transform applied to X before the max
 * (pointnet2/models/pnet.py:32-39: F.max_pool2d over the points after the shared MLP's GroupNorm + ReLU) */
enum slide_colmax_field { CM_X = 0, CM_LDX, CM_R, CM_C, CM_OUT, CM_LDO, CM_B, CM_STEP, CM_XF, CM_NFIELD = CM_XF + XF_NFIELD };

/* P f32 [ROWS, 2C] = [mean | logvar]; out[r,c] = mean[r,c] (NOISE < 0: posterior mode) or
 * mean[r,c] + exp(0.5 * clamp(logvar[r,c], -30, 20)) * noise[r,c]   (pointnet2/data_utils/distributions.py:4-17,41-42) */
enum slide_kl_field { KL_P = 0, KL_LDP, KL_C, KL_NOISE, KL_LDN, KL_OUT, KL_LDO, KL_ROWS };

#ifdef __cplusplus
}
#endif
#endif /* SLIDE_PROGRAM_H */
