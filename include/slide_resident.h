/*
 * slide_resident.h -- record layout of SAMPLE-RESIDENT plans (slide_program_set_resident in slide_b200.h).
 *
 * A resident plan executes a range of slide_op records (slide_program.h) as ONE kernel launch: every sample of the
 * batch is owned by one thread-block cluster, its activations live in shared memory from the first layer to the last,
 * GroupNorm statistics are reduced in shared memory (and across the cluster through distributed shared memory) and
 * the only HBM traffic of a step is the sample's input/output rows, its noise and the (L2-resident) weights.
 * It exists for the denoisers over 16 latent points (position DDPM: pointnet2/util.py:197-259 loop body,
 * models/pointnet2_with_pcld_condition.py:286-489), whose per-layer kernels are launch-latency bound.
 *
 * The host side (slide_b200/resident.py) compiles the slide_op records of the range into `slide_rop` records: it maps
 * every tensor that never leaves the range to a shared-memory offset (liveness-based allocation, spilling to an
 * L2-resident scratch slot when the working set does not fit), turns every transform-on-load (XF block) into an
 * in-place XFORM pass placed after the statistics it needs are complete, and splits the points of a sample over the
 * CTAs of the cluster.  The semantics of every rop are those of the slide_op it came from.
 *
 * Row conventions inside a cluster of CL CTAs (rank r): a sample has NP points; CTA r OWNS points
 * [r*NP/CL, (r+1)*NP/CL).  POINT-level tensors (one row per point) are stored with all NP rows in every CTA
 * (computed redundantly, or published to the peers when they derive from pair rows).  PAIR-level tensors (one row per
 * (point, neighbour)) are stored with the owned points' rows only: local row lr <-> point p0 + lr / K.
 * Statistics buffers are float2 {sum, sum of squares} per group in shared memory; a buffer that receives pair-level
 * contributions is PARTIAL per CTA (point-level contributions to it are masked to the owned rows) and is completed by
 * an RS_STATSX rop (cluster barrier + sum of the peers' partials).
 */
#ifndef SLIDE_RESIDENT_H
#define SLIDE_RESIDENT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLIDE_ROP_NI 44
#define SLIDE_ROP_NF 3
#define SLIDE_RES_THREADS 512
#define SLIDE_RES_WCHUNK 16   /* K columns per staged weight chunk */
#define SLIDE_RES_WPAD 4      /* + pad floats per staged weight row (row stride 20 floats = 4 mod 8: conflict-free ldmatrix) */
#define SLIDE_RES_NBLK 128    /* max output columns of one RS_GEMM (the planner splits wider layers) */
#define SLIDE_RES_WSTAGES 3   /* weight-chunk ring: stages of (SLIDE_RES_NBLK + 8) rows */

/* Everything the kernel needs is resolved on the host: all offsets are 32-bit FLOAT indices (shared memory: from the
 * start of the dynamic shared memory; arena / weight blob: from their base), -1 = absent.  The kernel is bound by
 * instruction issue, so a record is decoded with a handful of 32-bit loads -- no pointer arithmetic in 64 bits, no
 * address-space dispatch, no integer division. */
struct slide_rop {
  int32_t kind; /* enum slide_rop_kind */
  float f[SLIDE_ROP_NF];
  int32_t i[SLIDE_ROP_NI];
};

enum slide_rop_kind {
  RS_NOP = 0,
  RS_COPY = 1,    /* dst[r, 0:n] = src[r, 0:n]  (loads of the sample's inputs, stores of its outputs, concatenations) */
  RS_KNN = 2,     /* pytorch3d knn_points over <= 32 reference points */
  RS_GEMM = 3,    /* C = act(A W^T + bias + ev[point] + res), statistics; or the fused soft-max tail (SMK) */
  RS_PAIR = 4,    /* factored conv over grouped rows (SLIDE_OP_PAIR) */
  RS_XFORM = 5,   /* in-place transform-on-load: GroupNorm from statistics + ReLU + additive vector */
  RS_STATSX = 6,  /* complete a PARTIAL statistics buffer across the cluster */
  RS_CSYNC = 7,   /* cluster barrier (after rows were published to the peers) */
  RS_DDPM = 8,    /* SLIDE_OP_DDPM_UPDATE on the sample's rows */
  RS_SPILL = 9,   /* shared memory -> this CTA's scratch slot (L2) */
  RS_FILL = 10,   /* scratch slot -> shared memory */
  RS_KIND_COUNT
};

/* RS_COPY: ROWS x COLS floats, row strides SLD / DLD.  SRC_G / DST_G: 0 = shared memory (float offset), 1 = arena, this
 * sample's block (float index = OFF + sample * SSTRIDE).  OWNED: 1 = only the rows of the owned points are copied. */
enum slide_rcopy_field { RC_SRC = 0, RC_SLD, RC_SRC_G, RC_SSTRIDE, RC_DST, RC_DLD, RC_DST_G, RC_DSTRIDE, RC_ROWS, RC_COLS, RC_OWNED };

/* RS_KNN: queries Q [P1 rows, stride QLD], references REF [P2 <= 32 rows]; K <= 16; IDX int [P1, K], D2 float [P1, K]
 * (dense shared-memory tables, D2 < 0: absent).  All P1 queries are computed by every CTA. */
enum slide_rknn_field { RK_Q = 0, RK_QLD, RK_REF, RK_RLD, RK_P1, RK_P2, RK_K, RK_IDX, RK_D2 };

/* RS_GEMM.  A [M, K] shared memory (rows = all points, or the owned pair rows when PAIRROWS); W: chunked TF32 copy in
 * the weight blob at float index WCH: [NCHUNK][NPAD rows][SLIDE_RES_WCHUNK + SLIDE_RES_WPAD] floats, zero padded.
 * C [M, N] shared memory.  BIAS: weight-blob float index or -1.  ACT 0/1 (relu).
 * EV: point-level tensor added per row: row index = point of the row (PAIRROWS: p0 + (row >> RPP_SHIFT)).
 * RES: tensor of C's shape added before ACT.  Statistics: ST (shared-memory float offset of the float2[groups] buffer or
 * -1), ST_CG, ST_NNORM, ST_CHOFF, f[0] = weight; ST_OWNED: count only rows of owned points (point-level producer of a
 * PARTIAL buffer).
 * SMK > 0: fused AttentionModule tail: scores = A W^T + bias; C[point, n] = sum_k RES[row, n] * softmax_k(scores[row, n])
 * over the SMK (8 or 16) rows of each point; C is a point-level tensor, row p0 + row / SMK, published to the peers.
 * NEXT_*: the next RS_GEMM's weight copy; its first NEXT_PF chunks (2, or 1 when an RS_PAIR runs in between -- that one
 * uses ring stages 1 and 2 as scratch) are prefetched under this rop's epilogue. */
enum slide_rgemm_field {
  RG_A = 0, RG_ALD, RG_C, RG_CLD, RG_EV, RG_EVLD, RG_RES, RG_RESLD, RG_M, RG_K, RG_N, RG_PAIRROWS, RG_RPP_SHIFT, RG_WCH,
  RG_NCHUNK, RG_NPAD, RG_BIAS, RG_ACT, RG_ST, RG_ST_CG, RG_ST_NNORM, RG_ST_CHOFF, RG_ST_OWNED, RG_SMK,
  RG_NEXT_WCH, RG_NEXT_NPAD, RG_NEXT_NCHUNK, RG_NEXT_PF
};

/* RS_PAIR: out[(i,k), n] = act(U[j, n] + x_j . WX[n] + c_i . WC[n] + bias[n] + d2_ik WD[n] + w_ik WW[n] + RES[(i,k), n]),
 * j = IDX[i, k], for the owned points i.  U / XYZ: source-point tensors (all rows); CTR: point-level (all rows);
 * IDX / D2: dense [NP, K] tables (D2 < 0: QueryAndGroup form).  WX, WC ([N,3]), WD, WW, BIAS: weight-blob float indices
 * (-1 absent).  OUT / RES: pair-level (owned rows).  Statistics as for RS_GEMM. */
enum slide_rpair_field {
  RP_U = 0, RP_ULD, RP_XYZ, RP_XLD, RP_CTR, RP_CLD, RP_OUT, RP_OLD, RP_RES, RP_RLD, RP_K, RP_IDX, RP_D2, RP_WX, RP_WC, RP_WD,
  RP_WW, RP_BIAS, RP_N, RP_ACT, RP_ST, RP_ST_CG, RP_ST_NNORM, RP_ST_CHOFF
};

/* RS_XFORM: X [ROWS, C] shared memory (row stride XLD), in place.  y = (x - mean_g) * rstd_g * gamma[ch] + beta[ch] for
 * ch = CHOFF + col < NNORM (statistics buffer ST, f[0] = 1 / elements per group); y = max(y, 0) if RELU;
 * y += arena[ADD + arow * ADDLD + col] with arow = sample (ADDMODE 0), the step counter (1) or 0 (2); ADD < 0: none.
 * ST < 0: no normalisation.  GAMMA / BETA: weight-blob float indices. */
enum slide_rxform_field { RX_X = 0, RX_XLD, RX_ROWS, RX_C, RX_ST, RX_CG, RX_NNORM, RX_CHOFF, RX_GAMMA, RX_BETA, RX_RELU, RX_ADD, RX_ADDLD, RX_ADDMODE };

/* RS_STATSX: ST float offset of the CTA's partial sums, NFLOATS (2 * groups), DST float offset of the totals (a separate
 * buffer: partials are never rewritten, so one cluster barrier suffices).  Every CTA ends up with the cluster-wide sums. */
enum slide_rstatsx_field { RT_ST = 0, RT_NFLOATS, RT_DST };

/* RS_DDPM: SLIDE_OP_DDPM_UPDATE (same modes / table) on rows of the owned points.  X: the sample's x (shared-memory
 * copy, all points, stride XLD); XG: x in the arena (float index of sample 0, stride XGLD, XGSTRIDE floats per sample);
 * EPS shared memory; NOISE: arena float index of [T, BROWS, NCOLS]; TABLE weight-blob float index; X0C / MASK: arena
 * float indices (per sample strides X0CSTRIDE / MASKSTRIDE), -1 absent; f[0] = clamp. */
enum slide_rddpm_field {
  RD_X = 0, RD_XLD, RD_XG, RD_XGLD, RD_XGSTRIDE, RD_EPS, RD_ELD, RD_X0C, RD_X0CLD, RD_X0CSTRIDE, RD_MASK, RD_MASKSTRIDE,
  RD_MODE, RD_NOISE, RD_NCOLS, RD_COL0, RD_TABLE, RD_BROWS
};

/* RS_SPILL / RS_FILL: NFLOATS floats between shared memory (float offset SMEM) and float offset SCRATCH of this CTA's
 * scratch slot. */
enum slide_rspill_field { RL_SMEM = 0, RL_NFLOATS, RL_SCRATCH };

/* Plan header passed to slide_program_set_resident. */
struct slide_resident_plan {
  int32_t first, count;      /* the slide_op range this plan replaces */
  int32_t cluster;           /* CTAs per sample: 1, 2 or 4 */
  int32_t np;                /* points per sample */
  int32_t smem_floats;       /* dynamic shared memory the rops address (floats) */
  int32_t stats_off, stats_floats; /* statistics region (zeroed at the start of every sample) */
  int32_t wstage_off;        /* float offset of the weight-chunk ring (SLIDE_RES_WSTAGES stages) */
  int32_t wstage_floats;     /* floats per stage */
  int32_t scratch_bytes;     /* scratch slot per CTA (spills), 0 = none */
  int64_t step_off;          /* arena byte offset of the step counter; the kernel uses t = counter - 1 and the last CTA
                                to finish stores t (what SLIDE_OP_STEP_BEGIN does) */
  int32_t precise;           /* 1: 3xTF32 products (validation against the fp32 oracle) */
  int32_t batch;             /* samples */
  int32_t reserved[2];
};

#ifdef __cplusplus
}
#endif
#endif /* SLIDE_RESIDENT_H */
