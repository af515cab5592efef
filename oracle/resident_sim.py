"""CPU interpreter of sample-resident plans (include/slide_resident.h, compiled by slide_b200/resident.py).

TEST INFRASTRUCTURE ONLY -- never imported by slide_b200/.  It executes the packed `slide_rop` records exactly as the
kernel's data flow does -- one shared-memory array per CTA of a cluster (NaN-filled, so a read of anything that was never
written or that another tensor's allocation overwrote shows up), owned rows only for pair-level tensors, partial
statistics + exchange, published rows, spills to a scratch slot -- but with plain numpy fp32 arithmetic.  Comparing its
result with oracle/ir_exec.py on the same arena checks the COMPILER (operand binding, shared-memory liveness, transform
placement, statistics ownership, weight packing) without a GPU; the GPU tests then only have to pin the kernel.
"""
import numpy as np
import torch

from slide_b200 import resident as R
from . import ops

RV = R.RV
KIND_NAME = {v: k for k, v in R.RKIND.items()}
EPS = 1e-5


class ResidentSim(object):
    def __init__(self, builder, hdr, rec):
        self.h = hdr[0]
        self.rec = rec
        self.weights = builder.weights_blob()
        self.cl = int(self.h["cluster"])
        self.np = int(self.h["np"])
        self.npl = self.np // self.cl

    # ---- operands ----------------------------------------------------------------------------------------------------
    @staticmethod
    def _sm(sm, off, ld, rows, cols, row0=0):
        """(rows x cols) window of a shared-memory array at float offset `off` (None when off < 0)."""
        if off < 0:
            return None
        idx = (np.arange(row0, row0 + rows)[:, None] * ld + np.arange(cols)[None, :])
        return _View(sm, off + idx)

    @staticmethod
    def _ar(arena, off, ld, rows, cols, row0=0):
        """(rows x cols) window of the arena at FLOAT index `off`."""
        if off < 0:
            return None
        buf = np.frombuffer(arena, dtype=np.float32)
        idx = (np.arange(row0, row0 + rows)[:, None] * ld + np.arange(cols)[None, :])
        return _View(buf, off + idx)

    def _w(self, off, n):
        return np.frombuffer(self.weights, dtype=np.float32, count=n, offset=4 * int(off))

    # ---- execution ---------------------------------------------------------------------------------------------------
    def run(self, arena):
        """arena: the uint8 array of an ir_exec.Machine (modified in place like the kernel modifies the device arena)."""
        h = self.h
        B = int(h["batch"])
        step_view = np.frombuffer(arena, dtype=np.int32, count=1, offset=int(h["step_off"]))
        t = int(step_view[0]) - 1
        self.max_smem = 0
        for s in range(B):
            sms = [np.full(int(h["smem_floats"]), np.nan, dtype=np.float32) for _ in range(self.cl)]
            scratch = [np.full(max(int(h["scratch_bytes"]) // 4, 1), np.nan, dtype=np.float32) for _ in range(self.cl)]
            for sm in sms:
                sm[int(h["stats_off"]):int(h["stats_off"]) + int(h["stats_floats"])] = 0
            for r in self.rec:
                kind = KIND_NAME[int(r["kind"])]
                fn = getattr(self, "_" + kind[3:].lower())
                if kind in ("RS_STATSX", "RS_CSYNC"):
                    fn(r, arena, sms, s, t)
                    continue
                for rank in range(self.cl):
                    fn(r, arena, sms, s, t, rank, scratch[rank])
        step_view[0] = t

    def _nop(self, *a):
        pass

    def _csync(self, *a):
        pass

    def _copy(self, r, arena, sms, s, t, rank, scratch):
        i = r["i"]
        g = lambda k: int(i[RV[k]])
        rows, cols, r0 = g("RC_ROWS"), g("RC_COLS"), 0
        if g("RC_OWNED"):
            r0, rows = rank * self.npl, self.npl
        sm = sms[rank]
        src = self._ar(arena, g("RC_SRC") + s * g("RC_SSTRIDE"), g("RC_SLD"), rows, cols, r0) if g("RC_SRC_G") else \
            self._sm(sm, g("RC_SRC"), g("RC_SLD"), rows, cols, r0)
        dst = self._ar(arena, g("RC_DST") + s * g("RC_DSTRIDE"), g("RC_DLD"), rows, cols, r0) if g("RC_DST_G") else \
            self._sm(sm, g("RC_DST"), g("RC_DLD"), rows, cols, r0)
        dst.set(src.get())

    def _knn(self, r, arena, sms, s, t, rank, scratch):
        i = r["i"]
        g = lambda k: int(i[RV[k]])
        P1, P2, K = g("RK_P1"), g("RK_P2"), g("RK_K")
        sm = sms[rank]
        q = self._sm(sm, g("RK_Q"), g("RK_QLD"), P1, 3).get()
        ref = self._sm(sm, g("RK_REF"), g("RK_RLD"), P2, 3).get()
        assert np.isfinite(q).all() and np.isfinite(ref).all()
        res = ops.knn_points(torch.from_numpy(q)[None], torch.from_numpy(ref)[None], K=K)
        io = g("RK_IDX")
        sm[io:io + P1 * K] = res.idx.reshape(-1).numpy().astype(np.int32).view(np.float32)
        do = g("RK_D2")
        if do >= 0:
            sm[do:do + P1 * K] = res.dists.reshape(-1).numpy()

    def _unpack_w(self, i, N, K):
        npad, nchunk = int(i[RV["RG_NPAD"]]), int(i[RV["RG_NCHUNK"]])
        buf = self._w(int(i[RV["RG_WCH"]]), nchunk * npad * R.WLD).reshape(nchunk, npad, R.WLD)
        assert nchunk == (K + R.WCH - 1) // R.WCH and npad == (N + 7) // 8 * 8
        w = np.concatenate([buf[q, :, :R.WCH] for q in range(nchunk)], axis=1)
        assert not w[N:].any() and not w[:, K:].any(), "weight copy padding must be zero"
        return w[:N, :K]

    def _stats_add(self, sm, st, cg, nnorm, choff, n0_cols, vals, weight):
        """vals: (rows, ncols) values of columns n0_cols (absolute local column indices of this rop)."""
        for j, n in enumerate(n0_cols):
            ch = choff + n
            if ch < nnorm:
                g = ch // cg
                v = vals[:, j].astype(np.float32)
                sm[st + 2 * g] += np.float32(weight) * v.sum(dtype=np.float32)
                sm[st + 2 * g + 1] += np.float32(weight) * (v * v).sum(dtype=np.float32)

    def _gemm(self, r, arena, sms, s, t, rank, scratch):
        i, f = r["i"], r["f"]
        g = lambda k: int(i[RV[k]])
        sm = sms[rank]
        M, K, N = g("RG_M"), g("RG_K"), g("RG_N")
        pair, rshift, smk = g("RG_PAIRROWS"), g("RG_RPP_SHIFT"), g("RG_SMK")
        p0 = rank * self.npl
        A = self._sm(sm, g("RG_A"), g("RG_ALD"), M, K).get()
        assert np.isfinite(A).all(), "GEMM A operand holds uninitialised / clobbered values"
        W = self._unpack_w(i, N, K)
        if not int(self.h["precise"]):
            A = R.tf32_rna(A)
        c = (torch.from_numpy(np.ascontiguousarray(A)) @ torch.from_numpy(np.ascontiguousarray(W)).t()).numpy()
        if g("RG_BIAS") >= 0:
            c = c + self._w(g("RG_BIAS"), N)
        point = (p0 + (np.arange(M) >> rshift)) if pair else np.arange(M)
        if smk > 0:
            assert (1 << rshift) == smk
            val = self._sm(sm, g("RG_RES"), g("RG_RESLD"), M, N).get()
            assert np.isfinite(val).all()
            w = torch.softmax(torch.from_numpy(c.astype(np.float32).reshape(M // smk, smk, N)), dim=1).numpy()
            out = (val.reshape(M // smk, smk, N) * w).sum(axis=1, dtype=np.float32)
            rows = p0 + np.arange(M // smk)
            for rk in range(self.cl):  # published to every CTA of the cluster
                self._sm(sms[rk], g("RG_C"), g("RG_CLD"), self.np, N).set_rows(rows, out)
            return
        ev = self._sm(sm, g("RG_EV"), g("RG_EVLD"), self.np, N)
        if ev is not None:
            e = ev.get()[point]
            assert np.isfinite(e).all()
            c = c + e
        res = self._sm(sm, g("RG_RES"), g("RG_RESLD"), M, N)
        if res is not None:
            rv = res.get()
            assert np.isfinite(rv).all()
            c = c + rv
        if g("RG_ACT") == 1:
            c = np.maximum(c, 0)
        c = c.astype(np.float32)
        self._sm(sm, g("RG_C"), g("RG_CLD"), M, N).set(c)
        st = g("RG_ST")
        if st >= 0:
            vals = c
            if g("RG_ST_OWNED"):
                vals = c[(point >= p0) & (point < p0 + self.npl)]
            self._stats_add(sm, st, g("RG_ST_CG"), g("RG_ST_NNORM"), g("RG_ST_CHOFF"), range(N), vals, f[0])

    def _pair(self, r, arena, sms, s, t, rank, scratch):
        i = r["i"]
        g = lambda k: int(i[RV[k]])
        sm = sms[rank]
        K, N = g("RP_K"), g("RP_N")
        p0, npl, npts = rank * self.npl, self.npl, self.np
        U = self._sm(sm, g("RP_U"), g("RP_ULD"), npts, N).get()
        X = self._sm(sm, g("RP_XYZ"), g("RP_XLD"), npts, 3).get()
        CT = self._sm(sm, g("RP_CTR"), g("RP_CLD"), npts, 3).get()
        io = g("RP_IDX")
        idx = sm[io:io + npts * K].view(np.int32).reshape(npts, K)[p0:p0 + npl]
        assert np.isfinite(U).all() and np.isfinite(X).all() and (idx >= 0).all() and (idx < npts).all()
        wx, wc = self._w(g("RP_WX"), N * 3).reshape(N, 3), self._w(g("RP_WC"), N * 3).reshape(N, 3)
        out = U[idx] + X[idx] @ wx.T + (CT[p0:p0 + npl] @ wc.T)[:, None, :]
        if g("RP_BIAS") >= 0:
            out = out + self._w(g("RP_BIAS"), N)
        do = g("RP_D2")
        if do >= 0:
            d2 = sm[do:do + npts * K].reshape(npts, K)[p0:p0 + npl][:, :, None]
            inv = (np.float32(1.0) / (d2 + np.float32(1e-8))).astype(np.float32)
            w = inv / inv.sum(axis=1, keepdims=True, dtype=np.float32)
            out = out + d2 * self._w(g("RP_WD"), N) + w * self._w(g("RP_WW"), N)
        c = out.reshape(npl * K, N).astype(np.float32)
        res = self._sm(sm, g("RP_RES"), g("RP_RLD"), npl * K, N)
        if res is not None:
            rv = res.get()
            assert np.isfinite(rv).all()
            c = c + rv
        if g("RP_ACT") == 1:
            c = np.maximum(c, 0)
        c = c.astype(np.float32)
        self._sm(sm, g("RP_OUT"), g("RP_OLD"), npl * K, N).set(c)
        st = g("RP_ST")
        if st >= 0:
            self._stats_add(sm, st, g("RP_ST_CG"), g("RP_ST_NNORM"), g("RP_ST_CHOFF"), range(N), c, 1.0)

    def _xform(self, r, arena, sms, s, t, rank, scratch):
        i, f = r["i"], r["f"]
        g = lambda k: int(i[RV[k]])
        sm = sms[rank]
        rows, C = g("RX_ROWS"), g("RX_C")
        X = self._sm(sm, g("RX_X"), g("RX_XLD"), rows, C)
        x = X.get()
        assert np.isfinite(x).all(), "XFORM input holds uninitialised / clobbered values"
        st, cg, nnorm, choff = g("RX_ST"), g("RX_CG"), g("RX_NNORM"), g("RX_CHOFF")
        y = x.copy()
        if st >= 0:
            gam, bet = self._w(g("RX_GAMMA"), nnorm), self._w(g("RX_BETA"), nnorm)
            for n in range(C):
                ch = choff + n
                if ch < nnorm:
                    gi = ch // cg
                    mean = sm[st + 2 * gi] * f[0]
                    var = max(sm[st + 2 * gi + 1] * f[0] - mean * mean, np.float32(0))
                    rstd = np.float32(1.0) / np.sqrt(np.float32(var + np.float32(EPS)))
                    a = rstd * gam[ch]
                    b = bet[ch] - mean * a
                    y[:, n] = x[:, n] * a + b
        if g("RX_RELU"):
            y = np.maximum(y, 0)
        if g("RX_ADD") >= 0:
            mode = g("RX_ADDMODE")
            arow = s if mode == 0 else (t if mode == 1 else 0)
            add = np.frombuffer(arena, dtype=np.float32, count=C, offset=4 * (g("RX_ADD") + g("RX_ADDLD") * arow))
            y = y + add
        X.set(y.astype(np.float32))

    def _statsx(self, r, arena, sms, s, t):
        i = r["i"]
        src, dst, n = int(i[RV["RT_ST"]]), int(i[RV["RT_DST"]]), int(i[RV["RT_NFLOATS"]])
        tot = np.zeros(n, dtype=np.float32)
        for sm in sms:
            tot = tot + sm[src:src + n]
        for sm in sms:
            sm[dst:dst + n] = tot

    def _ddpm(self, r, arena, sms, s, t, rank, scratch):
        i, f = r["i"], r["f"]
        g = lambda k: int(i[RV[k]])
        sm = sms[rank]
        mode, nc, c0 = g("RD_MODE"), g("RD_NCOLS"), g("RD_COL0")
        p0, npl = rank * self.npl, self.npl
        brows = g("RD_BROWS")
        x = self._sm(sm, g("RD_X"), g("RD_XLD"), npl, nc, p0).get()
        eps = self._sm(sm, g("RD_EPS"), g("RD_ELD"), npl, nc, p0).get()
        assert np.isfinite(x).all() and np.isfinite(eps).all()
        tab = self._w(g("RD_TABLE") + 8 * t, 8)
        noise = np.frombuffer(arena, dtype=np.float32, count=brows * nc, offset=4 * (g("RD_NOISE") + brows * nc * t))
        noise = noise.reshape(brows, nc)[s * self.np + p0:s * self.np + p0 + npl]
        f32 = np.float32
        if mode == 0:
            new = (x - f32(tab[0]) * eps) / f32(tab[1])
            if t > 0:
                new = new + f32(tab[2]) * noise
        elif mode == 2:
            new = x * f32(tab[0]) + (f32(tab[1]) * eps + f32(tab[2]) * noise)
        else:
            c1, c2, pm1, pm2, sig = [f32(v) for v in tab[:5]]
            x0 = c1 * x - c2 * eps
            if f[0] > 0:
                x0 = np.clip(x0, -f[0], f[0])
            if g("RD_X0C") >= 0:
                x0c = self._ar(arena, g("RD_X0C") + s * g("RD_X0CSTRIDE"), g("RD_X0CLD"), npl, nc, p0).get()
                m = self._ar(arena, g("RD_MASK") + s * g("RD_MASKSTRIDE"), 1, npl, 1, p0).get()
                x0 = x0 * m + x0c * (f32(1.0) - m)
            new = pm1 * x0 + pm2 * x
            new = new + (f32(0.0 if t == 0 else 1.0) * sig) * noise
        xg = self._ar(arena, g("RD_XG") + s * g("RD_XGSTRIDE"), g("RD_XGLD"), npl, nc, p0)
        cur = xg.get()
        cur[:, c0:] = new.astype(np.float32)[:, c0:]
        xg.set(cur)

    def _spill(self, r, arena, sms, s, t, rank, scratch):
        i = r["i"]
        o, n, so = int(i[RV["RL_SMEM"]]), int(i[RV["RL_NFLOATS"]]), int(i[RV["RL_SCRATCH"]])
        scratch[so:so + n] = sms[rank][o:o + n]
        sms[rank][o:o + n] = np.nan  # the region is reused by other tensors

    def _fill(self, r, arena, sms, s, t, rank, scratch):
        i = r["i"]
        o, n, so = int(i[RV["RL_SMEM"]]), int(i[RV["RL_NFLOATS"]]), int(i[RV["RL_SCRATCH"]])
        sms[rank][o:o + n] = scratch[so:so + n]


class _View(object):
    """Strided (rows x cols) window into a flat float32 buffer."""

    def __init__(self, buf, idx):
        self.buf, self.idx = buf, idx

    def get(self):
        return np.array(self.buf[self.idx], dtype=np.float32)

    def set(self, v):
        self.buf[self.idx] = v

    def set_rows(self, rows, v):
        self.buf[self.idx[rows]] = v
