"""CPU interpreter of slide programs (the records of include/slide_program.h).

TEST INFRASTRUCTURE ONLY -- never imported by slide_b200/.  It gives every op kind an independent numpy /
torch-CPU meaning so that (a) the lowering in slide_b200/nets.py can be checked against oracle/ref_model.py
without a GPU and (b) the CUDA executor can be checked record by record on the GPU box.
Neighbour search and sampling go through the C oracle (oracle/slide_oracle.c) for exact index semantics.
"""
import numpy as np
import torch

from slide_b200.program import V, KIND, KIND_NAME  # field indices only (parsed from the C header)
from . import ops

EPS = 1e-5
NXF = V["XF_NFIELD"]


class Machine(object):
    def __init__(self, builder):
        self.rec = builder.pack()
        self.arena = np.zeros(builder.arena_bytes, dtype=np.uint8)
        self.weights = builder.weights_blob()
        self.builder = builder

    # ---- raw views ---------------------------------------------------------------------------------
    def _mat(self, buf, off, rows, ld, dtype=np.float32):
        n = rows * ld
        return np.frombuffer(buf, dtype=dtype, count=n, offset=off).reshape(rows, ld)

    def A(self, off, rows, ld, cols=None, dtype=np.float32):
        m = self._mat(self.arena, off, rows, ld, dtype)
        return m if cols is None else m[:, :cols]

    def W(self, off, rows, ld, cols=None):
        m = self._mat(self.weights, off, rows, ld)
        return m if cols is None else m[:, :cols]

    def Wv(self, off, n):
        return np.frombuffer(self.weights, dtype=np.float32, count=n, offset=off)

    def view(self, t):
        dt = {"f32": np.float32, "i32": np.int32, "f64": np.float64}[t.dtype]
        return self.A(t.off, t.rows, t.ld, t.C, dt)

    def upload(self, t, value):
        v = np.asarray(value.detach().cpu().numpy() if hasattr(value, "detach") else value)
        self.view(t)[...] = v.reshape(t.rows, t.C)

    def download(self, t):
        return torch.from_numpy(np.array(self.view(t)))

    def step(self):
        return int(np.frombuffer(self.arena, dtype=np.int32, count=1, offset=self.builder.step.off)[0])

    def set_step(self, v):
        np.frombuffer(self.arena, dtype=np.int32, count=1, offset=self.builder.step.off)[0] = v

    # ---- transform-on-load ---------------------------------------------------------------------------
    def xf(self, x, p, base, step_off):
        """x: (rows, C) float32 array (copy).  Applies the XF block at p[base:base+NXF]."""
        stats, cg, nnorm, choff, gamma, beta, R, count, relu, addvec, addld, addmode = [int(v) for v in p[base:base + NXF]]
        rows, C = x.shape
        y = x.astype(np.float32).copy()
        if stats >= 0:
            G = nnorm // cg
            nB = (rows + R - 1) // R
            st = self.A(stats, nB, 2 * G, None, np.float64).reshape(nB, G, 2)
            mean = st[:, :, 0] / count
            var = np.maximum(st[:, :, 1] / count - mean * mean, 0.0)
            rstd = 1.0 / np.sqrt(var + EPS)
            ch = choff + np.arange(C)
            sel = ch < nnorm
            if sel.any():
                g = ch[sel] // cg
                s = np.arange(rows) // R
                gam = self.Wv(gamma, nnorm)[ch[sel]]
                bet = self.Wv(beta, nnorm)[ch[sel]]
                m = mean[s][:, g].astype(np.float32)
                r = rstd[s][:, g].astype(np.float32)
                y[:, sel] = (y[:, sel] - m) * r * gam + bet
        if relu:
            y = np.maximum(y, 0)
        if addvec >= 0:
            if addmode == 0:
                nB = (rows + R - 1) // R
                vec = self.A(addvec, nB, addld, C)[np.arange(rows) // R]
            elif addmode == 1:
                t = np.frombuffer(self.arena, dtype=np.int32, count=1, offset=step_off)[0]
                vec = self.A(addvec + 4 * addld * int(t), 1, addld, C)
            else:
                vec = self.A(addvec, 1, addld, C)
            y = y + vec
        return y.astype(np.float32)

    # ---- ops -----------------------------------------------------------------------------------------
    def run(self, first=0, count=None):
        count = len(self.rec) - first if count is None else count
        for i in range(first, first + count):
            r = self.rec[i]
            getattr(self, "op_" + KIND_NAME[int(r["kind"])][len("SLIDE_OP_"):].lower())(r["p"], r["f"])

    def run_segment(self, name):
        self.run(*self.builder.segments[name])

    def op_nop(self, p, f):
        pass

    def op_join(self, p, f):
        pass  # records are stored in a valid sequential order

    def op_step_begin(self, p, f):
        off, n = int(p[V["SB_ZERO_OFF"]]), int(p[V["SB_ZERO_BYTES"]])
        if n > 0:
            self.arena[off:off + n] = 0
        np.frombuffer(self.arena, dtype=np.int32, count=1, offset=int(p[V["SB_STEP"]]))[0] -= 1

    def op_knn(self, p, f):
        g = lambda k: int(p[V[k]])
        B, P1, P2, K = g("KNN_B"), g("KNN_P1"), g("KNN_P2"), g("KNN_K")
        q = torch.from_numpy(np.array(self.A(g("KNN_Q"), B * P1, g("KNN_LDQ"), 3))).reshape(B, P1, 3)
        r = torch.from_numpy(np.array(self.A(g("KNN_REF"), B * P2, g("KNN_LDR"), 3))).reshape(B, P2, 3)
        res = ops.knn_points(q, r, K=K)
        self.A(g("KNN_IDX"), B * P1, K, None, np.int32)[...] = res.idx.reshape(B * P1, K).numpy().astype(np.int32)
        if g("KNN_D2") >= 0:
            self.A(g("KNN_D2"), B * P1, K)[...] = res.dists.reshape(B * P1, K).numpy()

    def op_group(self, p, f):
        g = lambda k: int(p[V[k]])
        B, N, npnt, K, C, mode = g("GRP_B"), g("GRP_N"), g("GRP_NP"), g("GRP_K"), g("GRP_C"), g("GRP_MODE")
        idx = self.A(g("GRP_IDX"), B * npnt, K, None, np.int32).reshape(B, npnt, K).astype(np.int64)
        xyz = self.A(g("GRP_XYZ"), B * N, g("GRP_LDX"), 3).reshape(B, N, 3)
        ctr = self.A(g("GRP_CTR"), B * npnt, g("GRP_LDCTR"), 3).reshape(B, npnt, 3)
        bi = np.arange(B)[:, None, None]
        xj = xyz[bi, idx]                                 # (B,np,K,3)
        ci = np.broadcast_to(ctr[:, :, None, :], xj.shape)
        parts = []
        if C > 0:
            F = self.A(g("GRP_F"), B * N, g("GRP_LDF"), C).reshape(B, N, C)
            parts.append(F[bi, idx])
        if mode == 0:
            parts.append(xj - ci)
            if g("GRP_ABS"):
                parts.append(xj)
            if g("GRP_CENTER"):
                parts.append(ci)
        else:
            d2 = self.A(g("GRP_D2"), B * npnt, K).reshape(B, npnt, K, 1)
            inv = (np.float32(1.0) / (d2 + np.float32(1e-8))).astype(np.float32)
            w = inv / inv.sum(axis=2, keepdims=True, dtype=np.float32)
            parts += [d2, w.astype(np.float32), xj, xj - ci, ci]
        out = np.concatenate(parts, axis=3).astype(np.float32)
        Ct = out.shape[3]
        self.A(g("GRP_OUT"), B * npnt * K, g("GRP_LDO"), Ct)[...] = out.reshape(B * npnt * K, Ct)

    def op_gemm(self, p, f):
        g = lambda k: int(p[V[k]])
        M, K, N = g("GEMM_M"), g("GEMM_K"), g("GEMM_N")
        step_off = g("GEMM_STEP")
        a = self.xf(np.array(self.A(g("GEMM_A"), M, g("GEMM_LDA"), K)), p, V["GEMM_XFA"], step_off)
        w = self.W(g("GEMM_W_W"), N, g("GEMM_LDW"), K)
        c = (torch.from_numpy(a) @ torch.from_numpy(np.array(w)).t()).numpy()
        if g("GEMM_BIAS_W") >= 0:
            c = c + self.Wv(g("GEMM_BIAS_W"), N)
        if g("GEMM_EV") >= 0:
            div = g("GEMM_EVDIV")
            ev = self.A(g("GEMM_EV"), M // div, g("GEMM_EVLD"), N)
            c = c + ev[np.arange(M) // div]
        smk = g("GEMM_SMK")
        if smk > 0:
            val = self.xf(np.array(self.A(g("GEMM_RES"), M, g("GEMM_LDR"), N)), p, V["GEMM_XFR"], step_off)
            w = torch.softmax(torch.from_numpy(c.astype(np.float32).reshape(M // smk, smk, N)), dim=1).numpy()
            out = (val.reshape(M // smk, smk, N) * w).sum(axis=1, dtype=np.float32)
            self.A(g("GEMM_C"), M // smk, g("GEMM_LDC"), N)[...] = out
            return
        if g("GEMM_RES") >= 0:
            c = c + self.xf(np.array(self.A(g("GEMM_RES"), M, g("GEMM_LDR"), N)), p, V["GEMM_XFR"], step_off)
        act = g("GEMM_ACT")
        if act == 1:
            c = np.maximum(c, 0)
        elif act == 2:
            c = c * (1.0 / (1.0 + np.exp(-c)))
        c = c.astype(np.float32)
        self.A(g("GEMM_C"), M, g("GEMM_LDC"), N)[...] = c
        if g("GEMM_ST_STATS") >= 0:
            cg, nnorm, choff, R, wgt = g("GEMM_ST_CG"), g("GEMM_ST_NNORM"), g("GEMM_ST_CHOFF"), g("GEMM_ST_R"), g("GEMM_ST_WEIGHT")
            G = nnorm // cg
            nB = M // R
            st = self.A(g("GEMM_ST_STATS"), nB, 2 * G, None, np.float64).reshape(nB, G, 2)
            ch = choff + np.arange(N)
            sel = ch < nnorm
            if sel.any():
                v = c[:, sel].astype(np.float64).reshape(nB, R, -1)
                grp = ch[sel] // cg
                for gi in np.unique(grp):
                    cols = grp == gi
                    st[:, gi, 0] += wgt * v[:, :, cols].sum(axis=(1, 2))
                    st[:, gi, 1] += wgt * (v[:, :, cols] ** 2).sum(axis=(1, 2))

    def op_softmax_wsum(self, p, f):
        g = lambda k: int(p[V[k]])
        rows, K, C = g("SM_ROWS"), g("SM_K"), g("SM_C")
        s = np.array(self.A(g("SM_S"), rows * K, g("SM_LDS"), C)).reshape(rows, K, C)
        v = self.xf(np.array(self.A(g("SM_V"), rows * K, g("SM_LDV"), C)), p, V["SM_XFV"], g("SM_STEP")).reshape(rows, K, C)
        w = torch.softmax(torch.from_numpy(s), dim=1).numpy()
        self.A(g("SM_OUT"), rows, g("SM_LDO"), C)[...] = (v * w).sum(axis=1, dtype=np.float32)

    def op_copy_cols(self, p, f):
        g = lambda k: int(p[V[k]])
        rows, n = g("CP_ROWS"), g("CP_NCOLS")
        self.A(g("CP_DST"), rows, g("CP_LDD"), n)[...] = np.array(self.A(g("CP_SRC"), rows, g("CP_LDS"), n))

    def op_ddpm_update(self, p, f):
        g = lambda k: int(p[V[k]])
        rows, nc, c0 = g("DD_ROWS"), g("DD_NCOLS"), g("DD_COL0")
        t = int(np.frombuffer(self.arena, dtype=np.int32, count=1, offset=g("DD_STEP"))[0])
        tab = self.Wv(g("DD_TABLE_W") + 32 * t, 8)
        x = self.A(g("DD_X"), rows, g("DD_LDX"), nc)
        eps = self.A(g("DD_EPS"), rows, g("DD_LDE"), nc)
        noise = self.A(g("DD_NOISE") + 4 * rows * nc * t, rows, nc)
        f32 = np.float32
        if g("DD_MODE") == 0:
            k1, sa, sig = f32(tab[0]), f32(tab[1]), f32(tab[2])
            new = (x - k1 * eps) / sa
            if t > 0:
                new = new + sig * noise
        elif g("DD_MODE") == 2:
            a_, c_, sig = f32(tab[0]), f32(tab[1]), f32(tab[2])
            new = x * a_ + (c_ * eps + sig * noise)
        else:
            c1, c2, pm1, pm2, sig = [f32(v) for v in tab[:5]]
            x0 = c1 * x - c2 * eps
            if f[0] > 0:
                x0 = np.clip(x0, -f[0], f[0])
            if g("DD_X0C") >= 0:  # local resampling (diffusion.py:76-79)
                x0c = self.A(g("DD_X0C"), rows, g("DD_LDX0C"), nc)
                m = self.A(g("DD_MASK"), rows, 1, 1)
                x0 = x0 * m + x0c * (f32(1.0) - m)
            new = pm1 * x0 + pm2 * x
            m = f32(0.0 if t == 0 else 1.0)
            new = new + (m * sig) * noise
        x[:, c0:] = new.astype(np.float32)[:, c0:]

    def op_fps(self, p, f):
        g = lambda k: int(p[V[k]])
        B, N, m = g("FPS_B"), g("FPS_N"), g("FPS_M")
        xyz = torch.from_numpy(np.array(self.A(g("FPS_XYZ"), B * N, g("FPS_LDX"), 3))).reshape(B, N, 3)
        if g("FPS_MODE") == 0:
            out = ops.furthest_point_sampling(xyz, m).numpy()
        else:
            start = None
            if g("FPS_START") >= 0:
                start = torch.from_numpy(np.array(self.A(g("FPS_START"), 1, B, None, np.int32))[0].astype(np.int64))
            else:
                start = torch.zeros(B, dtype=torch.int64)
            _, idx = ops.sample_farthest_points(xyz, K=m, start_idx=start)
            out = idx.numpy().astype(np.int32)
        self.A(g("FPS_OUT"), B, m, None, np.int32)[...] = out

    def op_gather_rows(self, p, f):
        g = lambda k: int(p[V[k]])
        B, N, m, nc = g("GA_B"), g("GA_N"), g("GA_M"), g("GA_NCOLS")
        src = self.A(g("GA_SRC"), B * N, g("GA_LDS"), nc).reshape(B, N, nc)
        idx = self.A(g("GA_IDX"), B, m, None, np.int32).astype(np.int64)
        self.A(g("GA_DST"), B * m, g("GA_LDD"), nc)[...] = src[np.arange(B)[:, None], idx].reshape(B * m, nc)

    def op_upsample(self, p, f):
        g = lambda k: int(p[V[k]])
        rows, fac, Fd, cc = g("UP_ROWS"), g("UP_FACTOR"), g("UP_F"), g("UP_COARSE_C")
        coarse = np.zeros((rows, Fd), dtype=np.float32)
        coarse[:, :cc] = self.A(g("UP_COARSE"), rows, g("UP_LDC"), cc)
        disp = np.array(self.A(g("UP_DISP"), rows, g("UP_LDD"), fac * Fd)).reshape(rows, fac, Fd)
        out = coarse[:, None, :] + (disp * np.float32(f[0])) * np.float32(f[1])
        self.A(g("UP_OUT"), rows * fac, g("UP_LDO"), Fd)[...] = out.reshape(rows * fac, Fd).astype(np.float32)

    def op_pair(self, p, f):
        g = lambda k: int(p[V[k]])
        B, Ns, npnt, K, N = g("PR_B"), g("PR_NSRC"), g("PR_NP"), g("PR_K"), g("PR_N")
        idx = self.A(g("PR_IDX"), B * npnt, K, None, np.int32).reshape(B, npnt, K).astype(np.int64)
        U = self.A(g("PR_U"), B * Ns, g("PR_LDU"), N).reshape(B, Ns, N)
        xyz = self.A(g("PR_XYZ"), B * Ns, g("PR_LDX"), 3).reshape(B, Ns, 3)
        ctr = self.A(g("PR_CTR"), B * npnt, g("PR_LDCTR"), 3).reshape(B, npnt, 3)
        bi = np.arange(B)[:, None, None]
        wc = self.W(g("PR_WC_W"), N, 3)
        if g("PR_WX_W") >= 0:
            wx = self.W(g("PR_WX_W"), N, 3)
            out = U[bi, idx] + xyz[bi, idx] @ wx.T + (ctr @ wc.T)[:, :, None, :]
        else:  # the neighbour-coordinate term is already part of U
            out = U[bi, idx] + (ctr @ wc.T)[:, :, None, :]
        if g("PR_BIAS_W") >= 0:
            out = out + self.Wv(g("PR_BIAS_W"), N)
        if g("PR_D2") >= 0:
            d2 = self.A(g("PR_D2"), B * npnt, K).reshape(B, npnt, K, 1)
            inv = (np.float32(1.0) / (d2 + np.float32(1e-8))).astype(np.float32)
            w = inv / inv.sum(axis=2, keepdims=True, dtype=np.float32)
            out = out + d2 * self.Wv(g("PR_WD_W"), N) + w * self.Wv(g("PR_WW_W"), N)
        M = B * npnt * K
        c = out.reshape(M, N).astype(np.float32)
        if g("PR_RES") >= 0:
            c = c + self.xf(np.array(self.A(g("PR_RES"), M, g("PR_LDR"), N)), p, V["PR_XFR"], g("PR_STEP"))
        act = g("PR_ACT")
        if act == 1:
            c = np.maximum(c, 0)
        c = c.astype(np.float32)
        self.A(g("PR_OUT"), M, g("PR_LDO"), N)[...] = c
        if g("PR_ST_STATS") >= 0:
            cg, nnorm, choff, wgt = g("PR_ST_CG"), g("PR_ST_NNORM"), g("PR_ST_CHOFF"), g("PR_ST_WEIGHT")
            G = nnorm // cg
            R = npnt * K
            st = self.A(g("PR_ST_STATS"), B, 2 * G, None, np.float64).reshape(B, G, 2)
            ch = choff + np.arange(N)
            sel = ch < nnorm
            if sel.any():
                v = c[:, sel].astype(np.float64).reshape(B, R, -1)
                grp = ch[sel] // cg
                for gi in np.unique(grp):
                    cols = grp == gi
                    st[:, gi, 0] += wgt * v[:, :, cols].sum(axis=(1, 2))
                    st[:, gi, 1] += wgt * (v[:, :, cols] ** 2).sum(axis=(1, 2))

    def op_colmax(self, p, f):
        g = lambda k: int(p[V[k]])
        B, R, C = g("CM_B"), g("CM_R"), g("CM_C")
        x = self.xf(np.array(self.A(g("CM_X"), B * R, g("CM_LDX"), C)), p, V["CM_XF"], g("CM_STEP"))
        self.A(g("CM_OUT"), B, g("CM_LDO"), C)[...] = x.reshape(B, R, C).max(axis=1)

    def op_kl(self, p, f):
        g = lambda k: int(p[V[k]])
        rows, C = g("KL_ROWS"), g("KL_C")
        P = torch.from_numpy(np.array(self.A(g("KL_P"), rows, g("KL_LDP"), 2 * C)))
        mean, logvar = P[:, :C], P[:, C:]
        if g("KL_NOISE") >= 0:
            noise = torch.from_numpy(np.array(self.A(g("KL_NOISE"), rows, g("KL_LDN"), C)))
            mean = mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise
        self.A(g("KL_OUT"), rows, g("KL_LDO"), C)[...] = mean.numpy()

    def op_temb(self, p, f):
        g = lambda k: int(p[V[k]])
        rows, half = g("TE_ROWS"), g("TE_HALF")
        ts = torch.from_numpy(np.array(np.frombuffer(self.arena, dtype=np.float32, count=rows, offset=g("TE_TS"))))
        freq = torch.from_numpy(np.array(self.Wv(g("TE_FREQ_W"), half)))
        arg = ts.unsqueeze(1) * freq
        out = torch.cat((torch.sin(arg), torch.cos(arg)), 1).numpy()
        self.A(g("TE_OUT"), rows, g("TE_LDO"), 2 * half)[...] = out
