"""Compiler from a range of slide_op records to a SAMPLE-RESIDENT plan (include/slide_resident.h).

The per-record executor launches one kernel per record and keeps every tensor in HBM; for the denoisers over 16 latent
points (65 records of at most 256 rows per sample and <= 139 channels in the position DDPM) that is pure launch latency.
A resident plan runs the same records as ONE kernel: a thread-block cluster owns a sample and keeps its tensors in shared
memory.  This module is the host-side compiler (pure python + numpy, no GPU needed -- tests interpret its output with
oracle/resident_sim.py):

  * classify every arena tensor of the range: INTERNAL (never leaves the range -> shared memory), EXTERNAL per-sample
    (x, eps -> shared-memory shadow + load / store), GLOBAL (noise, timestep tables, condition vectors);
  * split the points of a sample over the CTAs of the cluster (point-level tensors replicated, pair-level tensors
    owned rows only) and decide which statistics buffers are partial per CTA;
  * turn every transform-on-load (XF block) into an in-place XFORM rop placed before its consumer, and complete partial
    statistics with a STATSX rop after their last producer;
  * split GEMMs wider than 128 columns, pack a chunked TF32 (or fp32 for PRECISE plans) copy of every weight matrix for
    the kernel's cp.async ring and chain the "next GEMM" prefetch fields;
  * allocate shared memory by liveness (first fit), spilling the tensor with the farthest next use to the CTA's scratch
    slot when the working set does not fit.

Reference semantics are those of the records (nets.py cites the reference for each).
"""
import os

import numpy as np

from . import program as P
from .program import KIND_NAME

_HDR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "slide_resident.h")
RCONST, RENUM, RV = P._parse_header(_HDR, r"SLIDE_R\w+")
NI, NF = RCONST["SLIDE_ROP_NI"], RCONST["SLIDE_ROP_NF"]
WCH, WPAD, NBLK, WSTAGES = RCONST["SLIDE_RES_WCHUNK"], RCONST["SLIDE_RES_WPAD"], RCONST["SLIDE_RES_NBLK"], RCONST["SLIDE_RES_WSTAGES"]
WLD = WCH + WPAD
ROP_DTYPE = np.dtype([("kind", "<i4"), ("f", "<f4", (NF,)), ("i", "<i4", (NI,))])
PLAN_DTYPE = np.dtype([("first", "<i4"), ("count", "<i4"), ("cluster", "<i4"), ("np", "<i4"), ("smem_floats", "<i4"),
                       ("stats_off", "<i4"), ("stats_floats", "<i4"), ("wstage_off", "<i4"), ("wstage_floats", "<i4"),
                       ("scratch_bytes", "<i4"), ("step_off", "<i8"), ("precise", "<i4"), ("batch", "<i4"),
                       ("reserved", "<i4", (2,))])
assert ROP_DTYPE.itemsize == 16 + 4 * NI and PLAN_DTYPE.itemsize == 64
RKIND = RENUM["slide_rop_kind"]
SMEM_LIMIT_FLOATS = (227 * 1024 - 2 * ROP_DTYPE.itemsize - 64) // 4
STAGE_FLOATS = (NBLK + 8) * WLD

XF_NAMES = ["XF_STATS", "XF_CG", "XF_NNORM", "XF_CHOFF", "XF_GAMMA_W", "XF_BETA_W", "XF_R", "XF_COUNT", "XF_RELU",
            "XF_ADDVEC", "XF_ADDLD", "XF_ADDMODE"]

# arena-offset fields per record kind (for the inside / outside classification)
_OFFSET_FIELDS = {
    "SLIDE_OP_KNN": ["KNN_Q", "KNN_REF", "KNN_IDX", "KNN_D2"],
    "SLIDE_OP_GROUP": ["GRP_F", "GRP_XYZ", "GRP_CTR", "GRP_IDX", "GRP_D2", "GRP_OUT"],
    "SLIDE_OP_GEMM": ["GEMM_A", "GEMM_C", "GEMM_EV", "GEMM_RES", "GEMM_ST_STATS", ("GEMM_XFA", 0), ("GEMM_XFA", 9),
                      ("GEMM_XFR", 0), ("GEMM_XFR", 9)],
    "SLIDE_OP_SOFTMAX_WSUM": ["SM_S", "SM_V", "SM_OUT", ("SM_XFV", 0), ("SM_XFV", 9)],
    "SLIDE_OP_COPY_COLS": ["CP_SRC", "CP_DST"],
    "SLIDE_OP_DDPM_UPDATE": ["DD_X", "DD_EPS", "DD_NOISE", "DD_X0C", "DD_MASK"],
    "SLIDE_OP_FPS": ["FPS_XYZ", "FPS_OUT", "FPS_START"],
    "SLIDE_OP_GATHER_ROWS": ["GA_SRC", "GA_IDX", "GA_DST"],
    "SLIDE_OP_UPSAMPLE": ["UP_COARSE", "UP_DISP", "UP_OUT"],
    "SLIDE_OP_TEMB": ["TE_TS", "TE_OUT"],
    "SLIDE_OP_COLMAX": ["CM_X", "CM_OUT", ("CM_XF", 0), ("CM_XF", 9)],
    "SLIDE_OP_KL": ["KL_P", "KL_NOISE", "KL_OUT"],
    "SLIDE_OP_PAIR": ["PR_U", "PR_XYZ", "PR_CTR", "PR_IDX", "PR_D2", "PR_OUT", "PR_RES", "PR_ST_STATS", ("PR_XFR", 0),
                      ("PR_XFR", 9)],
}


class Unsupported(Exception):
    """The range cannot be made resident (the per-record executor stays in charge)."""


def _ld_smem(C):
    """Row stride of a shared-memory tensor: >= C, = 4 (mod 8) floats (conflict-free ldmatrix rows, 16-byte aligned)."""
    ld = (C + 3) // 4 * 4
    while ld % 8 != 4:
        ld += 4
    return ld


def _log2(v):
    if v < 1 or v & (v - 1):
        raise Unsupported("%d rows per point (a power of two is required)" % v)
    return int(v).bit_length() - 1


def tf32_rna(a):
    """Round fp32 to TF32 (10-bit mantissa), nearest with ties away from zero -- what cvt.rna.tf32.f32 does."""
    bits = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    out = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).astype(np.uint32)
    return out.view(np.float32)


def pack_weight(w, precise):
    """(N, K) fp32 -> [nchunk][npad][WLD] float32 (zero padded) for the kernel's cp.async ring."""
    N, K = w.shape
    npad, nchunk = (N + 7) // 8 * 8, (K + WCH - 1) // WCH
    buf = np.zeros((nchunk, npad, WLD), dtype=np.float32)
    src = w if precise else tf32_rna(w)
    for q in range(nchunk):
        k0, k1 = q * WCH, min(K, (q + 1) * WCH)
        buf[q, :N, :k1 - k0] = src[:, k0:k1]
    return buf, npad, nchunk


class RTensor(object):
    """A tensor of the range as the resident kernel sees it."""

    def __init__(self, base, np_, npl):
        self.base = base                      # program.Tensor (arena)
        self.name = base.name
        if base.R == np_:
            self.level, self.rpp = "point", 1
            self.rows = np_
        elif base.R % np_ == 0:
            self.level, self.rpp = "pair", base.R // np_
            self.rows = npl * self.rpp
        else:
            raise Unsupported("tensor %s has %d rows per sample (points per sample: %d)" % (base.name, base.R, np_))
        self.is_int = base.dtype == "i32"
        self.ld = base.C if self.is_int else _ld_smem(base.C)
        self.floats = (self.rows * self.ld + 3) // 4 * 4
        self.external = False
        self.pinned = False
        self.first = self.last = None
        self.uses = []                        # rop indices that touch it
        self.off = None                       # shared-memory float offset (current)
        self.scratch = None


class Rop(object):
    def __init__(self, kind, note=""):
        self.kind, self.note = kind, note
        self.i = {}          # field name -> int
        self.ops = {}        # operand base field -> ("smem", RTensor, coloff) | ("arena_s", off, ld, sstride) | ...
        self.f = [0.0] * NF
        self.touch = []      # RTensors this rop reads / writes
        self.stats_sym = {}  # field name -> (stats key, "partial" | "total")


class Planner(object):
    def __init__(self, builder, first, count, cluster=2, precise=False, np_points=16):
        self.b, self.first, self.count = builder, first, count
        self.cl, self.precise, self.np = cluster, bool(precise), np_points
        if np_points % cluster:
            raise Unsupported("points per sample not divisible by the cluster size")
        self.npl = np_points // cluster
        self.tensors = sorted([t for t in builder.tensors], key=lambda t: t.off)
        self.rt = {}          # base tensor offset -> RTensor
        self.rops = []
        self.stats = {}       # arena offset of a statistics tensor -> dict(G, partial, off_p, off_t)
        self._analyse()

    # ---- tensor lookup ------------------------------------------------------------------------------------------
    def _find(self, off):
        """arena byte offset -> (base Tensor, column offset in floats)."""
        lo, hi = 0, len(self.tensors) - 1
        while lo < hi:
            mid = (lo + hi + 1) // 2
            if self.tensors[mid].off <= off:
                lo = mid
            else:
                hi = mid - 1
        t = self.tensors[lo]
        if not (t.off <= off < t.off + max(t.nbytes, 4)):
            raise Unsupported("offset %d belongs to no tensor" % off)
        rel = (off - t.off) // 4
        row, col = divmod(rel, t.ld)
        if row != 0:
            raise Unsupported("row views are not supported in resident plans (%s)" % t.name)
        return t, col

    def _offsets_of(self, op):
        kind, fields = KIND_NAME[op[0]], op[1]
        out = []
        for key in _OFFSET_FIELDS.get(kind, []):
            v = fields.get(key, -1)
            if v is not None and v >= 0:
                out.append(v)
        return out

    def _rt(self, base):
        r = self.rt.get(base.off)
        if r is None:
            r = RTensor(base, self.np, self.npl)
            self.rt[base.off] = r
        return r

    # ---- analysis ------------------------------------------------------------------------------------------------
    def _analyse(self):
        b = self.b
        inside = range(self.first, self.first + self.count)
        outside_bases = set()
        for i, op in enumerate(b.ops):
            if i in inside:
                continue
            for off in self._offsets_of(op):
                outside_bases.add(self._find(off)[0].off)
        self.outside = outside_bases
        ops = [b.ops[i] for i in inside]
        if not ops or KIND_NAME[ops[0][0]] != "SLIDE_OP_STEP_BEGIN":
            raise Unsupported("a resident range starts with STEP_BEGIN (the kernel owns the step-counter update)")
        self.step_off = ops[0][1]["SB_STEP"]
        body = ops[1:]
        # statistics buffers: who contributes (pair-level producers make them partial per CTA)
        for op in body:
            kind, f = KIND_NAME[op[0]], op[1]
            if kind == "SLIDE_OP_GEMM" and f["GEMM_ST_STATS"] >= 0:
                out_base = self._find(f["GEMM_C"])[0]
                self._stat(f["GEMM_ST_STATS"], f["GEMM_ST_NNORM"] // f["GEMM_ST_CG"], self._rt(out_base).level == "pair")
            if kind == "SLIDE_OP_PAIR" and f["PR_ST_STATS"] >= 0:
                self._stat(f["PR_ST_STATS"], f["PR_ST_NNORM"] // f["PR_ST_CG"], True)
        # consumers' transforms must agree per tensor view
        self.xf_seen = {}
        self.materialised = set()
        self.last_stat_rop = {}
        loads = {}
        for op in body:
            kind = KIND_NAME[op[0]]
            if kind in ("SLIDE_OP_JOIN", "SLIDE_OP_NOP"):
                continue
            fn = getattr(self, "_lower_" + kind[len("SLIDE_OP_"):].lower(), None)
            if fn is None:
                raise Unsupported("record kind %s has no resident form" % kind)
            fn(op[1], op[2], op[3], loads)
        self._finish(loads)

    def _stat(self, off, G, pair):
        s = self.stats.setdefault(off, dict(G=G, partial=False, key=off))
        assert s["G"] == G
        if pair and self.cl > 1:
            s["partial"] = True

    # operand helpers -----------------------------------------------------------------------------------------------
    def _opnd(self, rop, field, off, write=False, loads=None, need_align=False):
        """Bind an arena offset to a resident operand (shared memory; external tensors get a shadow that is loaded on
        first read)."""
        if off is None or off < 0:
            rop.ops[field] = None
            return None
        base, col = self._find(off)
        rt = self._rt(base)
        if base.off in self.outside:
            rt.external = True
        if not write and not getattr(rt, "written", False):
            # read before any rop of the range wrote it: an input of the range (x) -> load its rows from the arena
            if loads is None:
                raise Unsupported("tensor %s is read before it is produced" % base.name)
            rt.external = True
            loads.setdefault(base.off, rt)
        if write:
            rt.written = True
            rt.smem_written = True
        if need_align and col % 4:
            raise Unsupported("GEMM operand %s starts at column %d (not 16-byte aligned)" % (base.name, col))
        rop.ops[field] = ("smem", rt, col)
        rop.touch.append(rt)
        return rt

    def _emit(self, rop):
        self.rops.append(rop)
        return rop

    def _xform(self, fields, xbase, off, rows_level_hint, note, loads):
        """Materialise the transform `fields[(xbase, i)]` on the tensor view at `off` (once)."""
        xf = [fields.get((xbase, i), d) for i, d in enumerate([-1, 1, 0, 0, -1, -1, 1, 1, 0, -1, 0, 0])]
        stats, cg, nnorm, choff, gamma, beta, R, count, relu, addvec, addld, addmode = xf
        trivial = stats < 0 and not relu and addvec < 0
        base, col = self._find(off)
        key = (base.off, col)
        sig = None if trivial else tuple(xf)
        if key in self.xf_seen:
            if self.xf_seen[key] != sig:
                raise Unsupported("tensor %s is consumed with two different transforms" % base.name)
            return
        self.xf_seen[key] = sig
        if trivial:
            return
        rt = self._rt(base)
        if base.off in self.outside:
            raise Unsupported("transform on external tensor %s" % base.name)
        rop = Rop("RS_XFORM", note + ".xform")
        self._opnd(rop, "RX_X", off, write=True)
        rop.i.update(RX_ROWS=rt.rows, RX_C=int(fields["__C__"]), RX_CG=cg, RX_NNORM=nnorm, RX_CHOFF=choff, RX_GAMMA=gamma,
                     RX_BETA=beta, RX_RELU=int(relu), RX_ADDMODE=addmode)
        rop.f[0] = np.float32(1.0) / np.float32(count) if count > 0 else 1.0
        if stats >= 0:
            rop.stats_sym["RX_ST"] = (stats, "total")
            if stats not in self.stats:
                raise Unsupported("statistics buffer of %s is produced outside the range" % base.name)
        else:
            rop.i["RX_ST"] = -1
        if addvec >= 0:
            rop.ops["RX_ADD"] = ("arena", addvec, addld, 0)
        else:
            rop.ops["RX_ADD"] = None
        self._emit(rop)

    # lowering of each record kind ------------------------------------------------------------------------------------
    def _lower_copy_cols(self, f, fl, note, loads):
        rop = Rop("RS_COPY", note)
        src = self._opnd(rop, "RC_SRC", f["CP_SRC"], loads=loads)
        dst = self._opnd(rop, "RC_DST", f["CP_DST"], write=True)
        if src.level != "point" or dst.level != "point":
            raise Unsupported("COPY_COLS of pair-level tensors")
        rop.i.update(RC_ROWS=self.np, RC_COLS=f["CP_NCOLS"], RC_OWNED=0)
        self._emit(rop)

    def _lower_knn(self, f, fl, note, loads):
        if f["KNN_P2"] > 32 or f["KNN_K"] > 16 or f["KNN_P1"] != self.np:
            raise Unsupported("kNN shape")
        rop = Rop("RS_KNN", note)
        self._opnd(rop, "RK_Q", f["KNN_Q"], loads=loads)
        self._opnd(rop, "RK_REF", f["KNN_REF"], loads=loads)
        rop.i.update(RK_P1=f["KNN_P1"], RK_P2=f["KNN_P2"], RK_K=f["KNN_K"])
        rop.ops["RK_IDX#"] = self._table(rop, f["KNN_IDX"], True)
        rop.ops["RK_D2#"] = self._table(rop, f["KNN_D2"], True)
        self._emit(rop)

    def _table(self, rop, off, write=False):
        """[NP, K] index / distance tables are addressed by a bare shared-memory offset."""
        if off is None or off < 0:
            return None
        base, col = self._find(off)
        assert col == 0
        rt = self._rt(base)
        if base.off in self.outside:
            raise Unsupported("table %s leaves the range" % base.name)
        if rt.level != "point":
            raise Unsupported("table %s is not [points, K]" % base.name)
        rt.ld = base.C  # dense rows: the kernel indexes tables as [point * K + k]
        rt.floats = (rt.rows * rt.ld + 3) // 4 * 4
        rop.touch.append(rt)
        return rt

    def _lower_gemm(self, f, fl, note, loads):
        M, K, N = f["GEMM_M"], f["GEMM_K"], f["GEMM_N"]
        smk = f.get("GEMM_SMK", 0)
        a_base = self._find(f["GEMM_A"])[0]
        a_rt = self._rt(a_base)
        pair = a_rt.level == "pair"
        # transforms first (in place)
        fx = dict(f)
        fx["__C__"] = K
        self._xform(fx, "GEMM_XFA", f["GEMM_A"], None, note + ".A", loads)
        if f["GEMM_RES"] >= 0:
            fx["__C__"] = N
            self._xform(fx, "GEMM_XFR", f["GEMM_RES"], None, note + ".R", loads)
        if f["GEMM_ACT"] not in (0, 1):
            raise Unsupported("activation %d" % f["GEMM_ACT"])
        if smk and smk not in (8, 16):
            raise Unsupported("soft-max over %d neighbours" % smk)
        w = self._weight_matrix(f["GEMM_W_W"], f["GEMM_LDW"], N, K)
        for n0 in range(0, N, NBLK):
            nb = min(NBLK, N - n0)
            rop = Rop("RS_GEMM", note + ("" if N <= NBLK else "[%d:%d]" % (n0, n0 + nb)))
            self._opnd(rop, "RG_A", f["GEMM_A"], loads=loads, need_align=True)
            c_rt = self._opnd(rop, "RG_C", f["GEMM_C"] + 4 * n0, write=True)
            ev_rt = self._opnd(rop, "RG_EV", f["GEMM_EV"] + 4 * n0 if f["GEMM_EV"] >= 0 else -1)
            self._opnd(rop, "RG_RES", f["GEMM_RES"] + 4 * n0 if f["GEMM_RES"] >= 0 else -1)
            if ev_rt is not None and (ev_rt.level != "point" or not pair or f["GEMM_EVDIV"] != a_rt.rpp):
                raise Unsupported("ev operand shape (%s)" % note)
            rows = a_rt.rows
            if rows % 16 or (rows > 128 and rows % 128) or (rows < 128 and rows not in (16, 32, 64)):
                raise Unsupported("GEMM over %d resident rows" % rows)
            buf, npad, nchunk = pack_weight(w[n0:n0 + nb], self.precise)
            rop.i.update(RG_M=rows, RG_K=K, RG_N=nb, RG_PAIRROWS=int(pair), RG_RPP_SHIFT=_log2(a_rt.rpp),
                         RG_WCH=self.b.weight(buf.reshape(-1)), RG_NCHUNK=nchunk, RG_NPAD=npad,
                         RG_BIAS=f["GEMM_BIAS_W"] + 4 * n0 if f["GEMM_BIAS_W"] >= 0 else -1, RG_ACT=f["GEMM_ACT"],
                         RG_ST_CG=f["GEMM_ST_CG"], RG_ST_NNORM=f["GEMM_ST_NNORM"], RG_ST_CHOFF=f["GEMM_ST_CHOFF"] + n0,
                         RG_ST_OWNED=0, RG_SMK=smk, RG_NEXT_WCH=-1, RG_NEXT_NPAD=0, RG_NEXT_NCHUNK=0, RG_NEXT_PF=0)
            rop.f[0] = float(f["GEMM_ST_WEIGHT"])
            if smk:
                if c_rt.level != "point" or not pair or a_rt.rpp != smk:
                    raise Unsupported("fused soft-max shape (%s)" % note)
                c_rt.pinned = True  # written by the peers: never share its address with another tensor
                rop.publish = True
            elif (c_rt.level == "pair") != pair:
                raise Unsupported("GEMM changes the row level (%s)" % note)
            if f["GEMM_ST_STATS"] >= 0:
                s = self.stats[f["GEMM_ST_STATS"]]
                rop.stats_sym["RG_ST"] = (f["GEMM_ST_STATS"], "partial")
                if s["partial"] and not pair:
                    rop.i["RG_ST_OWNED"] = 1
                self.last_stat_rop[f["GEMM_ST_STATS"]] = len(self.rops)
            else:
                rop.i["RG_ST"] = -1
            self._emit(rop)
        if smk and self.cl > 1:
            self._emit(Rop("RS_CSYNC", note + ".publish"))

    def _weight_matrix(self, off, ldw, N, K):
        for woff, arr in self.b.wchunks:
            if woff == off:
                return np.asarray(arr, dtype=np.float32).reshape(-1)[:N * ldw].reshape(N, ldw)[:, :K]
        raise Unsupported("weight matrix at %d not found" % off)

    def _lower_pair(self, f, fl, note, loads):
        fx = dict(f)
        fx["__C__"] = f["PR_N"]
        if f["PR_RES"] >= 0:
            self._xform(fx, "PR_XFR", f["PR_RES"], None, note + ".R", loads)
        if f["PR_ACT"] not in (0, 1):
            raise Unsupported("activation")
        if f["PR_NSRC"] != self.np or f["PR_NP"] != self.np:
            raise Unsupported("PAIR over %d source / %d centre points" % (f["PR_NSRC"], f["PR_NP"]))
        rop = Rop("RS_PAIR", note)
        self._opnd(rop, "RP_U", f["PR_U"], loads=loads)
        self._opnd(rop, "RP_XYZ", f["PR_XYZ"], loads=loads)
        self._opnd(rop, "RP_CTR", f["PR_CTR"], loads=loads)
        out = self._opnd(rop, "RP_OUT", f["PR_OUT"], write=True)
        self._opnd(rop, "RP_RES", f["PR_RES"])
        if out.level != "pair" or out.rpp != f["PR_K"]:
            raise Unsupported("PAIR output shape")
        rop.ops["RP_IDX#"] = self._table(rop, f["PR_IDX"])
        rop.ops["RP_D2#"] = self._table(rop, f["PR_D2"])
        rop.i.update(RP_K=f["PR_K"], RP_WX=f["PR_WX_W"], RP_WC=f["PR_WC_W"], RP_WD=f["PR_WD_W"], RP_WW=f["PR_WW_W"],
                     RP_BIAS=f["PR_BIAS_W"], RP_N=f["PR_N"], RP_ACT=f["PR_ACT"], RP_ST_CG=f["PR_ST_CG"],
                     RP_ST_NNORM=f["PR_ST_NNORM"], RP_ST_CHOFF=f["PR_ST_CHOFF"])
        if f["PR_ST_WEIGHT"] != 1:
            raise Unsupported("weighted PAIR statistics")
        if f["PR_ST_STATS"] >= 0:
            rop.stats_sym["RP_ST"] = (f["PR_ST_STATS"], "partial")
            self.last_stat_rop[f["PR_ST_STATS"]] = len(self.rops)
        else:
            rop.i["RP_ST"] = -1
        self._emit(rop)

    def _lower_ddpm_update(self, f, fl, note, loads):
        rop = Rop("RS_DDPM", note)
        x = self._opnd(rop, "RD_X", f["DD_X"], loads=loads)
        self._opnd(rop, "RD_EPS", f["DD_EPS"], loads=loads)
        xb, col = self._find(f["DD_X"])
        assert col == 0 and x.level == "point"
        rop.ops["RD_XG"] = ("arena_s", xb.off, xb.ld, xb.R * xb.ld * 4)
        for name, key, ldkey in (("RD_X0C", "DD_X0C", "DD_LDX0C"), ("RD_MASK", "DD_MASK", None)):
            off = f.get(key, -1)
            if off >= 0:
                tb, c0 = self._find(off)
                rop.ops[name] = ("arena_s", off, tb.ld, tb.R * tb.ld * 4)
            else:
                rop.ops[name] = None
        rop.i.update(RD_MODE=f["DD_MODE"], RD_NOISE=f["DD_NOISE"], RD_NCOLS=f["DD_NCOLS"], RD_COL0=f["DD_COL0"],
                     RD_TABLE=f["DD_TABLE_W"], RD_BROWS=f["DD_ROWS"])
        rop.f[0] = float(fl[0]) if fl else -1.0
        self._emit(rop)

    # ---- post passes -------------------------------------------------------------------------------------------------
    def _finish(self, loads):
        # 1. loads of external inputs at the top, stores of external outputs at the bottom
        head, tail = [], []
        for off, rt in loads.items():
            rop = Rop("RS_COPY", "load." + rt.name)
            base = rt.base
            rop.ops["RC_SRC"] = ("arena_s", base.off, base.ld, base.R * base.ld * 4)
            rop.ops["RC_DST"] = ("smem", rt, 0)
            rop.touch.append(rt)
            if rt.level != "point":
                raise Unsupported("external pair-level input %s" % rt.name)
            rop.i.update(RC_ROWS=self.np, RC_COLS=base.C, RC_OWNED=0)
            head.append(rop)
        for rt in self.rt.values():
            if not (rt.external and getattr(rt, "smem_written", False)):
                continue
            # an external tensor produced by the range (eps of a forward-only plan): every CTA stores its points' rows
            if rt.level != "point" or rt.is_int:
                raise Unsupported("external output %s" % rt.name)
            rop = Rop("RS_COPY", "store." + rt.name)
            base = rt.base
            rop.ops["RC_SRC"] = ("smem", rt, 0)
            rop.ops["RC_DST"] = ("arena_s", base.off, base.ld, base.R * base.ld * 4)
            rop.touch.append(rt)
            rop.i.update(RC_ROWS=self.np, RC_COLS=base.C, RC_OWNED=1)
            tail.append(rop)
        # 2. STATSX after the last producer of every partial statistics buffer
        body = []
        inserts = {}
        for off, idx in self.last_stat_rop.items():
            if self.stats[off]["partial"]:
                inserts.setdefault(idx, []).append(off)
        for i, rop in enumerate(self.rops):
            body.append(rop)
            for off in inserts.get(i, []):
                x = Rop("RS_STATSX", "statsx.%d" % off)
                x.stats_sym["RT_ST"] = (off, "partial")
                x.stats_sym["RT_DST"] = (off, "total")
                x.i["RT_NFLOATS"] = 2 * self.stats[off]["G"]
                body.append(x)
        # a DDPM update must come before the stores of the shadows it reads?  (it writes x in the arena directly)
        self.rops = head + body + tail
        # 3. the "next GEMM" prefetch chain
        gi = [i for i, r in enumerate(self.rops) if r.kind == "RS_GEMM"]
        for ia, ib in zip(gi, gi[1:]):
            a, nxt = self.rops[ia], self.rops[ib]
            between_pair = any(r.kind == "RS_PAIR" for r in self.rops[ia + 1:ib])  # RS_PAIR uses ring stages 1 + 2
            a.i.update(RG_NEXT_WCH=nxt.i["RG_WCH"], RG_NEXT_NPAD=nxt.i["RG_NPAD"], RG_NEXT_NCHUNK=nxt.i["RG_NCHUNK"],
                       RG_NEXT_PF=1 if between_pair else 2)
        # 4. shared-memory layout
        self._allocate()

    def _allocate(self):
        # fixed regions: statistics (partials then totals), weight ring
        cur = 0
        for off, s in sorted(self.stats.items()):
            n = 2 * s["G"]
            s["off_p"] = cur
            cur += (n + 3) // 4 * 4
            if s["partial"]:
                s["off_t"] = cur
                cur += (n + 3) // 4 * 4
            else:
                s["off_t"] = s["off_p"]
        self.stats_off, self.stats_floats = 0, cur
        self.wstage_off = cur
        cur += WSTAGES * STAGE_FLOATS
        pool0 = cur
        budget = SMEM_LIMIT_FLOATS - 8  # a few floats of slack for ldmatrix rows that start inside the last tensor
        # liveness
        for i, rop in enumerate(self.rops):
            for rt in rop.touch:
                rt.uses.append(i)
        tensors = [rt for rt in self.rt.values() if rt.uses]
        for rt in tensors:
            rt.first, rt.last = rt.uses[0], rt.uses[-1]
            if rt.pinned:
                rt.first, rt.last = 0, len(self.rops) - 1
        # offline placement first (all lifetimes are known): largest tensors first, each at the lowest address that is
        # free over its whole lifetime.  Only if that exceeds the budget does the online allocator below run, which can
        # spill.
        placed = []
        peak = pool0
        for rt in sorted(tensors, key=lambda t: (-t.floats, t.first)):
            pos = pool0
            for o in sorted((p for p in placed if not (p.last < rt.first or rt.last < p.first)), key=lambda p: p.off):
                if o.off - pos >= rt.floats:
                    break
                pos = max(pos, o.off + o.floats)
            rt.off = pos
            placed.append(rt)
            peak = max(peak, pos + rt.floats)
        if peak <= budget:
            self.peak = peak
            for rop in self.rops:
                rop.bound = self._bind(rop)
            self.final = list(self.rops)
            self.scratch_bytes = 0
            self.smem_floats = (self.peak + 8 + 3) // 4 * 4
            return
        for rt in tensors:
            rt.off = None
        live = []           # (off, floats, rt) sorted by off
        out = []            # final rop list with spills / fills
        scratch_cur = 0
        self.peak = pool0

        def try_place(rt):
            pos = pool0
            for off, n, _ in live:
                if off - pos >= rt.floats:
                    break
                pos = max(pos, off + n)
            if pos + rt.floats > budget:
                return False
            rt.off = pos
            live.append((pos, rt.floats, rt))
            live.sort(key=lambda e: e[0])
            self.peak = max(self.peak, pos + rt.floats)
            return True

        def release(rt):
            for e in live:
                if e[2] is rt:
                    live.remove(e)
                    return

        def next_use(rt, i):
            for u in rt.uses:
                if u >= i:
                    return u
            return 1 << 30

        def place(rt, i, needed):
            nonlocal scratch_cur
            while not try_place(rt):
                cands = [e[2] for e in live if e[2] not in needed and not e[2].pinned and next_use(e[2], i) < (1 << 30)]
                dead = [e[2] for e in live if next_use(e[2], i) == (1 << 30) and not e[2].pinned]
                if dead:
                    release(dead[0])
                    continue
                if not cands:
                    raise Unsupported("shared memory exhausted at rop %d (%s): %d floats needed" % (i, self.rops[i].note, rt.floats))
                victim = max(cands, key=lambda t: next_use(t, i))
                if victim.scratch is None:
                    victim.scratch = scratch_cur
                    scratch_cur += victim.floats * 4
                sp = Rop("RS_SPILL", "spill." + victim.name)
                sp.i.update(RL_SMEM=victim.off, RL_NFLOATS=victim.floats, RL_SCRATCH=victim.scratch)
                out.append(sp)
                release(victim)
                victim.off = None
                victim.spilled = True

        # pinned tensors first (whole-kernel lifetime, fixed addresses)
        for rt in tensors:
            if rt.pinned:
                if not try_place(rt):
                    raise Unsupported("pinned tensors do not fit")
        for i, rop in enumerate(self.rops):
            needed = list(rop.touch)
            for rt in needed:
                if rt.off is None:
                    was_spilled = getattr(rt, "spilled", False)
                    place(rt, i, needed)
                    if was_spilled:
                        fl = Rop("RS_FILL", "fill." + rt.name)
                        fl.i.update(RL_SMEM=rt.off, RL_NFLOATS=rt.floats, RL_SCRATCH=rt.scratch)
                        out.append(fl)
                        rt.spilled = False
            # bind addresses NOW (a tensor may move between a spill and its fill)
            rop.bound = self._bind(rop)
            out.append(rop)
            for rt in needed:
                if rt.last == i and not rt.pinned:
                    release(rt)
                    rt.off = None
        self.final = out
        self.scratch_bytes = (scratch_cur + 255) // 256 * 256
        self.smem_floats = (self.peak + 8 + 3) // 4 * 4

    # operand base name -> (offset field, stride field, [is-global field, sample-stride field])
    _OPERANDS = {
        "RG_A": ("RG_A", "RG_ALD"), "RG_C": ("RG_C", "RG_CLD"), "RG_EV": ("RG_EV", "RG_EVLD"), "RG_RES": ("RG_RES", "RG_RESLD"),
        "RP_U": ("RP_U", "RP_ULD"), "RP_XYZ": ("RP_XYZ", "RP_XLD"), "RP_CTR": ("RP_CTR", "RP_CLD"), "RP_OUT": ("RP_OUT", "RP_OLD"),
        "RP_RES": ("RP_RES", "RP_RLD"), "RX_X": ("RX_X", "RX_XLD"), "RX_ADD": ("RX_ADD", "RX_ADDLD"),
        "RK_Q": ("RK_Q", "RK_QLD"), "RK_REF": ("RK_REF", "RK_RLD"),
        "RC_SRC": ("RC_SRC", "RC_SLD", "RC_SRC_G", "RC_SSTRIDE"), "RC_DST": ("RC_DST", "RC_DLD", "RC_DST_G", "RC_DSTRIDE"),
        "RD_X": ("RD_X", "RD_XLD"), "RD_EPS": ("RD_EPS", "RD_ELD"), "RD_XG": ("RD_XG", "RD_XGLD", None, "RD_XGSTRIDE"),
        "RD_X0C": ("RD_X0C", "RD_X0CLD", None, "RD_X0CSTRIDE"), "RD_MASK": ("RD_MASK", None, None, "RD_MASKSTRIDE"),
    }
    # byte offsets into the weight blob / the arena that the record carries as float indices
    _W_BYTES = ("RG_WCH", "RG_BIAS", "RG_NEXT_WCH", "RP_WX", "RP_WC", "RP_WD", "RP_WW", "RP_BIAS", "RX_GAMMA", "RX_BETA", "RD_TABLE")
    _A_BYTES = ("RD_NOISE",)

    def _bind(self, rop):
        """Resolve symbolic operands to the numbers of the slide_rop record (32-bit float indices, -1 = absent)."""
        vals = dict(rop.i)
        for field, o in rop.ops.items():
            if field.endswith("#"):
                vals[field[:-1]] = -1 if o is None else o.off
                continue
            spec = self._OPERANDS[field]
            off_f, ld_f = spec[0], spec[1]
            g_f = spec[2] if len(spec) > 2 else None
            ss_f = spec[3] if len(spec) > 3 else None
            off, ld, is_g, ss = -1, 0, 0, 0
            if o is None:
                pass
            elif o[0] == "smem":
                rt, col = o[1], o[2]
                off, ld = rt.off + col, rt.ld
                if len(spec) > 2 and g_f is None:
                    raise Unsupported("%s must live in the arena" % field)
            elif o[0] in ("arena_s", "arena"):
                assert o[1] % 4 == 0 and o[3] % 4 == 0
                off, ld, is_g, ss = o[1] // 4, o[2], 1, o[3] // 4
                if len(spec) == 2 and field != "RX_ADD":
                    raise Unsupported("%s must live in shared memory" % field)
            else:
                raise AssertionError(o)
            vals[off_f] = off
            if ld_f is not None:
                vals[ld_f] = ld
            if g_f is not None:
                vals[g_f] = is_g
            if ss_f is not None:
                vals[ss_f] = ss
        for field, (key, which) in rop.stats_sym.items():
            s = self.stats[key]
            vals[field] = s["off_t"] if which == "total" else s["off_p"]
        for k in self._W_BYTES + self._A_BYTES:
            if k in vals and vals[k] >= 0:
                assert vals[k] % 4 == 0
                vals[k] //= 4
        if rop.kind in ("RS_SPILL", "RS_FILL"):
            vals["RL_SCRATCH"] //= 4
        for k, v in vals.items():
            if not (-2 ** 31 <= int(v) < 2 ** 31):
                raise Unsupported("offset %s = %d does not fit 32 bits" % (k, v))
        return vals

    # ---- output --------------------------------------------------------------------------------------------------------
    def pack(self):
        rec = np.zeros(len(self.final), dtype=ROP_DTYPE)
        for i, rop in enumerate(self.final):
            rec[i]["kind"] = RKIND[rop.kind]
            vals = rop.bound if hasattr(rop, "bound") else self._bind(rop)
            for key, v in vals.items():
                rec[i]["i"][RV[key]] = int(v)
            for j, v in enumerate(rop.f):
                rec[i]["f"][j] = v
        hdr = np.zeros(1, dtype=PLAN_DTYPE)
        h = hdr[0]
        h["first"], h["count"], h["cluster"], h["np"] = self.first, self.count, self.cl, self.np
        h["smem_floats"], h["stats_off"], h["stats_floats"] = self.smem_floats, self.stats_off, self.stats_floats
        h["wstage_off"], h["wstage_floats"], h["scratch_bytes"] = self.wstage_off, STAGE_FLOATS, self.scratch_bytes
        h["step_off"], h["precise"], h["batch"] = self.step_off, int(self.precise), self.b.B
        return hdr, rec

    def summary(self):
        kinds = {}
        for r in self.final:
            kinds[r.kind] = kinds.get(r.kind, 0) + 1
        return dict(rops=len(self.final), kinds=kinds, smem_bytes=self.smem_floats * 4, scratch_bytes=self.scratch_bytes,
                    cluster=self.cl, spills=kinds.get("RS_SPILL", 0))


def plan_segment(builder, name, cluster=2, precise=False, np_points=16):
    """Compile segment `name` of `builder` into a resident plan.  MUST run before Program(builder) is created: the packed
    weight copies are appended to the builder's weight blob.  Raises Unsupported if the range cannot be made resident."""
    first, count = builder.segments[name]
    return Planner(builder, first, count, cluster=cluster, precise=precise, np_points=np_points)
