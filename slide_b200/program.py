"""Host side of the fused network programs: an op-record builder, the arena / weight-blob allocators and the
ctypes runtime over slide_program_* (include/slide_b200.h, record layout in include/slide_program.h).

The builder is pure Python + numpy (no GPU needed), so the lowering in nets.py can be checked on CPU by
interpreting the records (tests do that with oracle/ir_exec.py); `Program` needs the CUDA library.
"""
import ctypes
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_HEADER = os.path.join(_HERE, "..", "include", "slide_program.h")


def _parse_header(path=None, const_pattern=r"SLIDE_OP_N\w+"):
    """Field indices and constants come from the C header: one source of truth for both sides."""
    src = open(_HEADER if path is None else path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(%s)\s+(\d+)" % const_pattern, src)}
    enums = {}
    values = {}
    for m in re.finditer(r"enum\s+(\w+)\s*\{(.*?)\}", src, flags=re.S):
        nxt = 0
        names = {}
        for item in m.group(2).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, expr = [s.strip() for s in item.split("=")]
                nxt = int(eval(expr, {}, values))
            else:
                name = item
            names[name] = nxt
            values[name] = nxt
            nxt += 1
        enums[m.group(1)] = names
    return consts, enums, values


CONSTS, ENUMS, V = _parse_header()
CONSTS_FLAGS = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(SLIDE_OPF_\w+)\s+(\d+)", open(_HEADER).read())}
NPARAM = CONSTS["SLIDE_OP_NPARAM"]
NFPARAM = CONSTS["SLIDE_OP_NFPARAM"]
OP_DTYPE = np.dtype([("kind", "<i4"), ("flags", "<i4"), ("p", "<i8", (NPARAM,)), ("f", "<f4", (NFPARAM,))])
KIND = ENUMS["slide_op_kind"]
KIND_NAME = {v: k for k, v in KIND.items()}
GN_EPS = 1e-5
ACT = {None: 0, "none": 0, "relu": 1, "swish": 2}


def _align(x, a):
    return (x + a - 1) // a * a


class Tensor(object):
    """A row-major matrix in the arena.  rows = B * R (R rows per sample); `off` is a byte offset."""

    def __init__(self, name, B, R, C, ld, off, dtype="f32"):
        self.name, self.B, self.R, self.C, self.ld, self.off, self.dtype = name, B, R, C, ld, off, dtype

    @property
    def rows(self):
        return self.B * self.R

    @property
    def nbytes(self):
        return self.rows * self.ld * (8 if self.dtype == "f64" else 4)

    def cols(self, c0, n=None):
        """View of columns [c0, c0+n) (same rows / ld)."""
        n = self.C - c0 if n is None else n
        assert 0 <= c0 and c0 + n <= self.ld
        t = Tensor(self.name + "[:,%d:%d]" % (c0, c0 + n), self.B, self.R, n, self.ld, self.off + 4 * c0, self.dtype)
        return t

    def rows_view(self, r0, nrows, R=None):
        """View of rows [r0, r0+nrows) treated as one sample block (B=1) unless R is given."""
        R = nrows if R is None else R
        assert nrows % R == 0
        return Tensor(self.name + "[%d:%d]" % (r0, r0 + nrows), nrows // R, R, self.C, self.ld,
                      self.off + 4 * self.ld * r0, self.dtype)


class XF(object):
    """Transform-on-load spec (see slide_program.h): GroupNorm from statistics + ReLU + additive vector."""

    def __init__(self, stats=None, cg=1, nnorm=0, choff=0, gamma=-1, beta=-1, R=1, count=1, relu=False,
                 addvec=None, addmode=0):
        self.stats, self.cg, self.nnorm, self.choff = stats, cg, nnorm, choff
        self.gamma, self.beta, self.R, self.count, self.relu = gamma, beta, R, count, relu
        self.addvec, self.addmode = addvec, addmode

    def fields(self):
        return [self.stats.off if self.stats is not None else -1, self.cg, self.nnorm, self.choff, self.gamma,
                self.beta, self.R, self.count, int(self.relu), self.addvec.off if self.addvec is not None else -1,
                self.addvec.ld if self.addvec is not None else 0, self.addmode]


NO_XF = XF()


class Stats(object):
    """fp64 [B, G, 2] (sum, sum of squares) per (sample, group) in the arena's statistics region."""

    def __init__(self, tensor, cg, nnorm, R, count):
        self.tensor, self.cg, self.nnorm, self.R, self.count = tensor, cg, nnorm, R, count


class Builder(object):
    """Accumulates tensors, weights and op records for one program."""

    def __init__(self, B, stats_capacity=None):
        self.B = B
        self.ops = []           # list of (kind, {field: value}, [floats], note)
        self.tensors = []
        # arena layout: [0,256) control words (step counter at 0) | statistics region | tensors
        self.stats_begin = 256
        self.stats_end = 256    # grows as statistics buffers are allocated
        self.stats_cap = 256 + _align(B * (1 << 17) if stats_capacity is None else stats_capacity, 256)
        self.arena_bytes = self.stats_cap
        self.wchunks = []
        self.weights_bytes = 0
        self.step = Tensor("step", 1, 1, 1, 1, 0, "i32")
        self.segments = {}      # name -> (first_op, n_ops)
        self._seg_open = None
        self._side = False      # records emitted while True carry SLIDE_OPF_SIDE

    # ---- allocation ------------------------------------------------------------------------------
    def tensor(self, name, R, C, B=None, dtype="f32", ld=None):
        B = self.B if B is None else B
        if ld is None:
            ld = C if dtype == "i32" else _align(C, 4)  # index tensors are dense, fp32 rows 16-byte aligned
        t = Tensor(name, B, R, C, ld, self.arena_bytes, dtype)
        self.arena_bytes = _align(self.arena_bytes + max(t.nbytes, 4), 256)
        self.tensors.append(t)
        return t

    def stats(self, name, nnorm, cg, R, count, B=None):
        """Statistics buffer for a tensor whose leading `nnorm` channels are normalised in groups of `cg`.
        All of them live in one region so that one STEP_BEGIN record zeroes them."""
        assert nnorm % cg == 0
        B = self.B if B is None else B
        t = Tensor(name, B, 1, 2 * (nnorm // cg), 2 * (nnorm // cg), self.stats_end, "f64")
        self.stats_end = _align(self.stats_end + t.nbytes, 256)
        assert self.stats_end <= self.stats_cap, "statistics region overflow; raise stats_capacity"
        self.tensors.append(t)
        return Stats(t, cg, nnorm, R, count)

    def weight(self, array):
        """Append an fp32 array to the weight blob; returns its byte offset (64 B aligned)."""
        a = np.ascontiguousarray(np.asarray(array, dtype=np.float32))
        off = self.weights_bytes
        self.wchunks.append((off, a))
        self.weights_bytes = _align(off + a.nbytes, 64)
        return off

    def weight_matrix(self, w):
        """(N, K) weight -> row-major [N, ldw] with ldw = K rounded up to 4 (zero padded); returns (off, ldw)."""
        w = np.asarray(w, dtype=np.float32)
        N, K = w.shape
        ldw = _align(K, 4)
        buf = np.zeros((N, ldw), dtype=np.float32)
        buf[:, :K] = w
        return self.weight(buf), ldw

    def weight_matrix_tc(self, w):
        """Tensor-core copy of an (N, K) weight: TF32-rounded (round to nearest even on the 13 dropped mantissa
        bits) and tiled exactly as the tcgen05 B operand is laid out in shared memory, so that the kernel
        fetches one (N tile, K block) with a single bulk copy.  Returns (offset, atoms per K block)."""
        w = np.asarray(w, dtype=np.float32)
        N, K = w.shape
        KB, NP = (K + 31) // 32, _align(N, 256)
        bits = np.zeros((NP, KB * 32), dtype=np.uint32)
        bits[:N, :K] = w.view(np.uint32)
        bits = ((bits + np.uint32(0xFFF) + ((bits >> np.uint32(13)) & np.uint32(1))) & np.uint32(0xFFFFE000))
        t = bits.reshape(NP // 8, 8, KB, 8, 4)          # [atom, r, kb, chunk, within]
        out = np.zeros((KB, NP // 8, 8, 8, 4), dtype=np.uint32)
        for r in range(8):
            for c in range(8):
                # logical 16-byte chunk c of row r lives at physical chunk c ^ r
                out[:, :, r, c ^ r, :] = t[:, r, :, c, :].transpose(1, 0, 2)
        return self.weight(out.view(np.float32).reshape(-1)), NP // 8

    def weights_blob(self):
        blob = np.zeros(max(self.weights_bytes, 64), dtype=np.uint8)
        for off, a in self.wchunks:
            blob[off:off + a.nbytes] = a.view(np.uint8).reshape(-1)
        return blob

    # ---- segments (ranges of ops that are run / captured together) --------------------------------
    def begin_segment(self, name):
        assert self._seg_open is None
        self._seg_open = (name, len(self.ops))

    def end_segment(self):
        name, first = self._seg_open
        self.segments[name] = (first, len(self.ops) - first)
        self._seg_open = None

    # ---- op emitters -----------------------------------------------------------------------------
    def _emit(self, kind, fields, floats=(), note=""):
        f = dict(fields)
        if self._side:
            f["__flags__"] = CONSTS_FLAGS["SLIDE_OPF_SIDE"]
        self.ops.append((KIND[kind], f, list(floats), note))

    def side_branch(self):
        """Context manager: records emitted inside run on the side branch (see SLIDE_OPF_SIDE)."""
        builder = self

        class _Side(object):
            def __enter__(self):
                assert not builder._side
                builder._side = True

            def __exit__(self, *exc):
                builder._side = False
        return _Side()

    def join(self, note="join"):
        self._emit("SLIDE_OP_JOIN", {}, note=note)

    def step_begin(self):
        """Must be emitted AFTER every statistics buffer of the program has been allocated... the byte count is
        patched in pack() so that later allocations are still covered."""
        self._emit("SLIDE_OP_STEP_BEGIN", {"SB_ZERO_OFF": self.stats_begin, "SB_ZERO_BYTES": 0,
                                           "SB_STEP": self.step.off}, note="step_begin")

    def knn(self, q, ref, K, idx, d2=None, note=""):
        assert q.B == ref.B == idx.B and idx.R == q.R and idx.C == K
        if K > ref.R:
            # pytorch3d pads the missing neighbours with index 0 / distance 0 and the reference's modules then treat them as
            # real neighbours (group_knn, pointnet2_utils.py:497-524); no shipped config gets here (K = 8 <= 16 points)
            raise NotImplementedError("%s: %d nearest neighbours among %d points (pytorch3d's zero-padded neighbours are "
                                      "not lowered)" % (note or "knn", K, ref.R))
        self._emit("SLIDE_OP_KNN", {"KNN_Q": q.off, "KNN_LDQ": q.ld, "KNN_P1": q.R, "KNN_REF": ref.off,
                                    "KNN_LDR": ref.ld, "KNN_P2": ref.R, "KNN_K": K, "KNN_IDX": idx.off,
                                    "KNN_D2": d2.off if d2 is not None else -1, "KNN_B": q.B}, note=note)

    def group(self, mode, feats, C, xyz, ctr, idx, K, out, d2=None, include_abs=True, include_center=True, note=""):
        npnt = ctr.R
        extra = 11 if mode == 1 else 3 + 3 * int(include_abs) + 3 * int(include_center)
        assert out.R == npnt * K and out.C == C + extra, (out.R, npnt, K, out.C, C, extra)
        self._emit("SLIDE_OP_GROUP", {
            "GRP_MODE": mode, "GRP_F": feats.off if feats is not None else -1,
            "GRP_LDF": feats.ld if feats is not None else 0, "GRP_C": C, "GRP_XYZ": xyz.off, "GRP_LDX": xyz.ld,
            "GRP_N": xyz.R, "GRP_CTR": ctr.off, "GRP_LDCTR": ctr.ld, "GRP_NP": npnt, "GRP_IDX": idx.off, "GRP_K": K,
            "GRP_D2": d2.off if d2 is not None else -1, "GRP_OUT": out.off, "GRP_LDO": out.ld,
            "GRP_ABS": int(include_abs), "GRP_CENTER": int(include_center), "GRP_B": out.B}, note=note)

    def gemm(self, A, W, out, bias=-1, act=None, xfa=NO_XF, ev=None, ev_div=1, resid=None, xfr=NO_XF, stats=None,
             st_R=None, st_choff=0, st_weight=1, smk=0, note=""):
        """out = act(xfa(A) W^T + bias + ev[row // ev_div] + xfr(resid)); W = (offset, ldw, N, K[, wp_off, wp_na])."""
        woff, ldw, N, K = W[:4]
        wp_off, wp_na = (W[4], W[5]) if len(W) > 4 else (-1, 0)
        assert A.C == K and out.C == N, (note, A.C, K, out.C, N)
        if smk:
            assert A.rows == out.rows * smk and resid is not None and ev is None and stats is None and act is None
        else:
            assert A.rows == out.rows, (note, A.rows, out.rows)
        f = {"GEMM_A": A.off, "GEMM_LDA": A.ld, "GEMM_M": A.rows, "GEMM_K": K, "GEMM_W_W": woff, "GEMM_LDW": ldw,
             "GEMM_N": N, "GEMM_C": out.off, "GEMM_LDC": out.ld, "GEMM_BIAS_W": bias, "GEMM_ACT": ACT[act],
             "GEMM_EV": ev.off if ev is not None else -1, "GEMM_EVLD": ev.ld if ev is not None else 0,
             "GEMM_EVDIV": ev_div, "GEMM_RES": resid.off if resid is not None else -1,
             "GEMM_LDR": resid.ld if resid is not None else 0,
             "GEMM_ST_STATS": stats.tensor.off if stats is not None else -1,
             "GEMM_ST_CG": stats.cg if stats is not None else 1, "GEMM_ST_NNORM": stats.nnorm if stats is not None else 0,
             "GEMM_ST_CHOFF": st_choff, "GEMM_ST_R": (st_R if st_R is not None else stats.R) if stats is not None else 1, "GEMM_ST_WEIGHT": st_weight,
             "GEMM_STEP": self.step.off, "GEMM_WP_W": wp_off, "GEMM_WP_NA": wp_na, "GEMM_SMK": smk}
        if ev is not None:
            assert ev.C == N and ev.rows * ev_div == A.rows
        if resid is not None:
            assert resid.C == N and resid.rows == A.rows
        for base, xf in (("GEMM_XFA", xfa), ("GEMM_XFR", xfr)):
            for i, val in enumerate(xf.fields()):
                f[(base, i)] = val
        self._emit("SLIDE_OP_GEMM", f, note=note)

    def softmax_wsum(self, S, Vt, xfv, out, rows, K, note=""):
        assert S.rows == rows * K == Vt.rows and S.C == Vt.C == out.C and out.rows == rows
        f = {"SM_S": S.off, "SM_LDS": S.ld, "SM_V": Vt.off, "SM_LDV": Vt.ld, "SM_OUT": out.off, "SM_LDO": out.ld,
             "SM_ROWS": rows, "SM_K": K, "SM_C": S.C, "SM_STEP": self.step.off}
        for i, val in enumerate(xfv.fields()):
            f[("SM_XFV", i)] = val
        self._emit("SLIDE_OP_SOFTMAX_WSUM", f, note=note)

    def copy_cols(self, src, dst, note=""):
        assert src.rows == dst.rows and src.C == dst.C
        self._emit("SLIDE_OP_COPY_COLS", {"CP_SRC": src.off, "CP_LDS": src.ld, "CP_DST": dst.off, "CP_LDD": dst.ld,
                                          "CP_ROWS": src.rows, "CP_NCOLS": src.C}, note=note)

    def ddpm_update(self, mode, x, eps, noise, table_off, col0=0, clamp=-1.0, x0c=None, mask=None, note=""):
        """x0c / mask: local resampling (diffusion.py:76-79), latent sampler only."""
        assert eps.rows == x.rows and eps.C == x.C
        assert (x0c is None) == (mask is None)
        if x0c is not None:
            assert mode == 1 and x0c.rows == x.rows and x0c.C == x.C and mask.rows == x.rows and mask.ld == 1
        self._emit("SLIDE_OP_DDPM_UPDATE", {"DD_MODE": mode, "DD_X": x.off, "DD_LDX": x.ld, "DD_EPS": eps.off,
                                            "DD_LDE": eps.ld, "DD_NOISE": noise.off, "DD_ROWS": x.rows,
                                            "DD_NCOLS": x.C, "DD_COL0": col0, "DD_TABLE_W": table_off,
                                            "DD_STEP": self.step.off,
                                            "DD_X0C": x0c.off if x0c is not None else -1,
                                            "DD_LDX0C": x0c.ld if x0c is not None else 0,
                                            "DD_MASK": mask.off if mask is not None else -1},
                   floats=[clamp], note=note)

    def fps(self, mode, xyz, m, out, start=None, note=""):
        assert out.R == 1 and out.C == m and out.dtype == "i32"
        self._emit("SLIDE_OP_FPS", {"FPS_MODE": mode, "FPS_XYZ": xyz.off, "FPS_LDX": xyz.ld, "FPS_N": xyz.R,
                                    "FPS_M": m, "FPS_OUT": out.off, "FPS_START": start.off if start is not None else -1,
                                    "FPS_B": xyz.B}, note=note)

    def gather_rows(self, src, idx, m, dst, note=""):
        assert dst.R == m and dst.C == src.C and idx.dtype == "i32"
        self._emit("SLIDE_OP_GATHER_ROWS", {"GA_SRC": src.off, "GA_LDS": src.ld, "GA_N": src.R, "GA_IDX": idx.off,
                                            "GA_M": m, "GA_DST": dst.off, "GA_LDD": dst.ld, "GA_NCOLS": src.C,
                                            "GA_B": dst.B}, note=note)

    def upsample(self, coarse, coarse_c, disp, out, factor, scale, note=""):
        Fd = out.C
        assert disp.C == factor * Fd and out.rows == coarse.rows * factor and disp.rows == coarse.rows
        self._emit("SLIDE_OP_UPSAMPLE", {"UP_COARSE": coarse.off, "UP_LDC": coarse.ld, "UP_COARSE_C": coarse_c,
                                         "UP_DISP": disp.off, "UP_LDD": disp.ld, "UP_OUT": out.off, "UP_LDO": out.ld,
                                         "UP_ROWS": coarse.rows, "UP_FACTOR": factor, "UP_F": Fd},
                   floats=[np.float32(1 / np.sqrt(factor)), np.float32(scale)], note=note)

    def pair(self, U, xyz, ctr, idx, K, out, wx, wc, bias=-1, d2=None, wd=-1, ww=-1, act=None, resid=None, xfr=NO_XF,
             stats=None, st_choff=0, st_weight=1, note=""):
        """Factored conv over grouped rows (SLIDE_OP_PAIR).  U: [B*Nsrc, N] view; out: [B*np*K, N]."""
        npnt = ctr.R
        assert out.R == npnt * K and out.C == U.C and idx.C == K and U.R == xyz.R
        f = {"PR_U": U.off, "PR_LDU": U.ld, "PR_NSRC": U.R, "PR_XYZ": xyz.off, "PR_LDX": xyz.ld, "PR_CTR": ctr.off,
             "PR_LDCTR": ctr.ld, "PR_NP": npnt, "PR_IDX": idx.off, "PR_K": K, "PR_D2": d2.off if d2 is not None else -1,
             "PR_WX_W": wx, "PR_WC_W": wc, "PR_WD_W": wd, "PR_WW_W": ww, "PR_BIAS_W": bias, "PR_N": out.C,
             "PR_OUT": out.off, "PR_LDO": out.ld, "PR_ACT": ACT[act], "PR_RES": resid.off if resid is not None else -1,
             "PR_LDR": resid.ld if resid is not None else 0,
             "PR_ST_STATS": stats.tensor.off if stats is not None else -1, "PR_ST_CG": stats.cg if stats is not None else 1,
             "PR_ST_NNORM": stats.nnorm if stats is not None else 0, "PR_ST_CHOFF": st_choff, "PR_ST_WEIGHT": st_weight,
             "PR_B": out.B, "PR_STEP": self.step.off}
        for i, val in enumerate(xfr.fields()):
            f[("PR_XFR", i)] = val
        self._emit("SLIDE_OP_PAIR", f, note=note)

    def colmax(self, X, xf, R, out, note=""):
        assert out.C == X.C and X.rows == out.rows * R
        f = {"CM_X": X.off, "CM_LDX": X.ld, "CM_R": R, "CM_C": X.C, "CM_OUT": out.off, "CM_LDO": out.ld,
             "CM_B": out.rows, "CM_STEP": self.step.off}
        for i, val in enumerate(xf.fields()):
            f[("CM_XF", i)] = val
        self._emit("SLIDE_OP_COLMAX", f, note=note)

    def kl(self, P, out, noise=None, note=""):
        C = out.C
        assert P.C == 2 * C and P.rows == out.rows
        self._emit("SLIDE_OP_KL", {"KL_P": P.off, "KL_LDP": P.ld, "KL_C": C, "KL_NOISE": noise.off if noise is not None else -1,
                                   "KL_LDN": noise.ld if noise is not None else 0, "KL_OUT": out.off, "KL_LDO": out.ld,
                                   "KL_ROWS": out.rows}, note=note)

    def temb(self, ts, freq_off, half, out, note=""):
        assert out.C == 2 * half and ts.C == out.rows
        self._emit("SLIDE_OP_TEMB", {"TE_TS": ts.off, "TE_FREQ_W": freq_off, "TE_HALF": half, "TE_OUT": out.off,
                                     "TE_LDO": out.ld, "TE_ROWS": out.rows}, note=note)

    # ---- packing ---------------------------------------------------------------------------------
    def pack(self):
        rec = np.zeros(len(self.ops), dtype=OP_DTYPE)
        for i, (kind, fields, floats, _note) in enumerate(self.ops):
            rec[i]["kind"] = kind
            if kind == KIND["SLIDE_OP_STEP_BEGIN"]:
                fields = dict(fields, SB_ZERO_BYTES=self.stats_end - self.stats_begin)
            for key, val in fields.items():
                if key == "__flags__":
                    rec[i]["flags"] = int(val)
                    continue
                idx = V[key[0]] + key[1] if isinstance(key, tuple) else V[key]
                rec[i]["p"][idx] = int(val)
            for j, fv in enumerate(floats):
                rec[i]["f"][j] = fv
        return rec


# ---------------------------------------------------------------------------------------------------
# runtime (needs libslide_b200.so and a GPU)
# ---------------------------------------------------------------------------------------------------
class Program(object):
    """A packed program resident on one GPU.  upload()/download() move tensors between torch and the arena."""

    def __init__(self, builder, device=None):
        import torch
        from . import lib as _l
        self._l = _l
        self.lib = _l.load()
        self.builder = builder
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        rec = builder.pack()
        blob = builder.weights_blob()
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.slide_program_create(rec.ctypes.data_as(ctypes.c_void_p), len(rec),
                                               ctypes.c_size_t(builder.arena_bytes),
                                               blob.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(blob.nbytes),
                                               ctypes.byref(handle))
        _l.check(rc, "slide_program_create")
        self.handle = handle
        self.arena_ptr = self.lib.slide_program_arena(handle)
        self.n_ops = len(rec)
        self._torch = torch

    def close(self):
        if getattr(self, "handle", None):
            self.lib.slide_program_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def view(self, t):
        """torch view (rows, ld) of an arena tensor -- zero-copy, via the CUDA array interface."""
        torch = self._torch
        dt = {"f32": (torch.float32, "<f4", 4), "i32": (torch.int32, "<i4", 4), "f64": (torch.float64, "<f8", 8)}[t.dtype]

        class _Raw(object):
            pass

        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (t.rows, t.ld), "typestr": dt[1],
                                        "data": (self.arena_ptr + t.off, False), "version": 2, "strides": None}
        return torch.as_tensor(raw, device=self.device)

    def raw_arena(self):
        """uint8 torch view of the whole arena (tests use it to teacher-force single records)."""
        class _Raw(object):
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (self.builder.arena_bytes,), "typestr": "|u1",
                                        "data": (self.arena_ptr, False), "version": 2, "strides": None}
        return self._torch.as_tensor(raw, device=self.device)

    def set_gemm_backend(self, name):
        """'auto' (tcgen05 where eligible) or 'simt' (fp32 FFMA everywhere)."""
        self._l.check(self.lib.slide_program_set_gemm_backend(self.handle, {"auto": 0, "simt": 1}[name]),
                      "slide_program_set_gemm_backend")

    def set_resident(self, plan):
        """Install a sample-resident plan (slide_b200.resident.Planner, built BEFORE this Program was created so that
        its packed weight copies are in the weight blob)."""
        hdr, rec = plan.pack()
        assert plan.b is self.builder
        with self._torch.cuda.device(self.device):
            self._l.check(self.lib.slide_program_set_resident(self.handle, hdr.ctypes.data_as(ctypes.c_void_p),
                                                              rec.ctypes.data_as(ctypes.c_void_p), len(rec)),
                          "slide_program_set_resident")

    def use_resident(self, enable):
        self._l.check(self.lib.slide_program_use_resident(self.handle, int(bool(enable))), "slide_program_use_resident")

    def upload(self, t, value):
        """Copy a (rows, C) torch/numpy array into arena tensor t (async on the current stream)."""
        torch = self._torch
        v = torch.as_tensor(value)
        v = v.reshape(t.rows, t.C)
        dst = self.view(t)[:, :t.C]
        dst.copy_(v.to(dst.dtype), non_blocking=True)

    def download(self, t):
        return self.view(t)[:, :t.C].clone()

    def set_step(self, value):
        self.view(self.builder.step).fill_(int(value))

    def run(self, first, count):
        with self._torch.cuda.device(self.device):
            self._l.check(self.lib.slide_program_run(self.handle, int(first), int(count), self._stream()),
                          "slide_program_run")

    def run_segment(self, name):
        self.run(*self.builder.segments[name])

    def capture(self, slot, first, count, repeat=1):
        with self._torch.cuda.device(self.device):
            self._l.check(self.lib.slide_program_capture(self.handle, int(slot), int(first), int(count), int(repeat),
                                                         self._stream()), "slide_program_capture")

    def replay(self, slot, times):
        with self._torch.cuda.device(self.device):
            self._l.check(self.lib.slide_program_replay(self.handle, int(slot), int(times), self._stream()),
                          "slide_program_replay")

    def launches(self, first, count):
        return int(self.lib.slide_program_launches(self.handle, int(first), int(count)))
