"""Record-level parity at the BASELINE.json batch sizes, DEFAULT dispatch (no environment overrides).

The kernel a GEMM record runs is chosen from its shape (tile width from N, persistent warp-specialised kernel from 148
tiles up), so the instantiations that bench.py times at batch 32 / 128 / 256 are not the ones a batch-8 test launches.
These tests run every record of one sampling step of both denoisers at the batch sizes of BASELINE configs 2 (B=32),
3 (B=128), 4 (B=32 of 256) and the headline (B=256) with backend "auto" and compare each record's outputs with the CPU
interpreter (oracle/ir_exec.py) on identical inputs.  After every record the interpreter's outputs are copied to the
GPU, so each comparison isolates ONE kernel launch (no drift across layers) and the tolerance is the kernel's own:
TF32 operands, fp32 accumulate.  Reference behaviour: pointnet2/util.py:225-259, diffusion_utils/diffusion.py:58-95.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ir_exec
from slide_b200 import engine, lib
from slide_b200.program import Program, KIND_NAME, V
from tests import common

pytestmark = pytest.mark.gpu

TOL = {"simt": 2e-4, "auto": 4e-3}

OUT_FIELDS = {
    "SLIDE_OP_KNN": ["KNN_IDX", "KNN_D2"], "SLIDE_OP_GROUP": ["GRP_OUT"], "SLIDE_OP_GEMM": ["GEMM_C", "GEMM_ST_STATS"],
    "SLIDE_OP_SOFTMAX_WSUM": ["SM_OUT"], "SLIDE_OP_COPY_COLS": ["CP_DST"], "SLIDE_OP_DDPM_UPDATE": ["DD_X"],
    "SLIDE_OP_FPS": ["FPS_OUT"], "SLIDE_OP_GATHER_ROWS": ["GA_DST"], "SLIDE_OP_UPSAMPLE": ["UP_OUT"],
    "SLIDE_OP_TEMB": ["TE_OUT"], "SLIDE_OP_COLMAX": ["CM_OUT"], "SLIDE_OP_KL": ["KL_OUT"],
    "SLIDE_OP_PAIR": ["PR_OUT", "PR_ST_STATS"],
}


def outputs_of(b, rec):
    """Arena tensors a record writes (looked up from its output offsets; a column view maps to its base tensor)."""
    kind = KIND_NAME[int(rec["kind"])]
    outs = []
    for f in OUT_FIELDS.get(kind, []):
        off = int(rec["p"][V[f]])
        if off < 0:
            continue
        for t in b.tensors:
            if t.off <= off < t.off + max(t.nbytes, 1):
                outs.append(t)
                break
        else:
            raise AssertionError("no tensor at offset %d (%s)" % (off, f))
    return outs


def isolated_records(b, m, prog, first, count, rtol, log):
    """Run records [first, first+count) on both machines, compare what each wrote, then copy the interpreter's result
    over the GPU's so the next record starts from identical inputs."""
    rec = b.pack()
    raw = prog.raw_arena()
    raw.copy_(torch.from_numpy(m.arena))
    worst = []
    for i in range(first, first + count):
        kind = KIND_NAME[int(rec[i]["kind"])]
        m.run(i, 1)
        prog.run(i, 1)
        torch.cuda.synchronize()
        touched = outputs_of(b, rec[i])
        if kind == "SLIDE_OP_STEP_BEGIN":
            raw.copy_(torch.from_numpy(m.arena))
            continue
        gpu = {}
        bad = []
        for t in touched:
            g = raw[t.off:t.off + t.nbytes].cpu().numpy()
            dt = {"f32": np.float32, "i32": np.int32, "f64": np.float64}[t.dtype]
            a = np.frombuffer(m.arena, dtype=dt, count=t.rows * t.ld, offset=t.off).reshape(t.rows, t.ld)[:, :t.C]
            gg = np.frombuffer(g, dtype=dt).reshape(t.rows, t.ld)[:, :t.C]
            if t.dtype == "i32":
                if not np.array_equal(a, gg):
                    bad.append((t.name, float((a != gg).mean()), 1.0))
            elif not np.isfinite(gg).all():
                bad.append((t.name, float("nan"), 0.0))
            else:
                scale = float(np.abs(a).max()) if a.size else 0.0
                err = float(np.abs(a.astype(np.float64) - gg.astype(np.float64)).max()) if a.size else 0.0
                # statistics (fp64 sums of ~1e5 values): relative to the sum-of-squares scale
                if err > rtol * max(scale, 1e-6) + 1e-7:
                    bad.append((t.name, err, scale))
            raw[t.off:t.off + t.nbytes].copy_(torch.from_numpy(m.arena[t.off:t.off + t.nbytes]))
        log.append("op %3d %-22s %-28s %s" % (i, kind, b.ops[i][3], "ok" if not bad else bad[:4]))
        if bad:
            worst.append((i, kind, b.ops[i][3], bad[:4]))
    return worst


@pytest.mark.parametrize("which,B", [("pos", 32), ("lat", 32), ("lat", 128), ("pos", 256), ("lat", 256)])
def test_step_records_at_baseline_batch(which, B, pipeline_cfg):
    """BASELINE config 2 (position DDPM, B=32), config 3 (feature DDPM, B=128), config 4 (32 per GPU) and the headline
    batch 256: every record of one sampling step under the default dispatch."""
    b, h, pc, sd = common.ddpm_program(pipeline_cfg, which, B, with_noise=True, T=4)
    m = ir_exec.Machine(b)
    common.init_machine(m, h, np.arange(B) % 13)
    g = torch.Generator().manual_seed(1000 + B)
    m.upload(h["x"], torch.randn(B * 16, h["C"], generator=g))
    m.upload(h["noise"], torch.randn(h["noise"].rows, h["C"], generator=g))
    m.set_step(3)
    m.run_segment("setup")  # setup records are batch-size independent GEMMs: checked at B=8 elsewhere
    prog = Program(b)
    prog.set_gemm_backend("auto")
    log = []
    lib.reset_launch_count()
    bad = isolated_records(b, m, prog, *b.segments["step"], rtol=TOL["auto"], log=log)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/records_%s_auto_b%d.log" % (which, B), "w") as f:
        f.write("\n".join(log) + "\n")
    assert lib.load().slide_tc_error() == 0, "tcgen05 pipeline wait timed out"
    assert not bad, bad[:6]


@pytest.mark.parametrize("stage", ["decode", "encode"])
def test_autoencoder_records_default_dispatch(stage, golden, pipeline_cfg):
    """Decode / encode programs under the default dispatch (their pair-level GEMMs have >= 148 row tiles already at
    B=2, so the persistent tcgen05 kernels run), record by record against the interpreter."""
    B = 2
    sd = common.state_dict("ae")
    aec = pipeline_cfg["autoencoder"]
    if stage == "decode":
        b, h = engine.build_decode(aec["decoders"], sd, B)
        m = ir_exec.Machine(b)
        engine.init_constants(m, h)
        m.upload(h["labels"], golden["label"].astype(np.int32))
        m.upload(h["keypoint"], golden["dec_kp"])
        m.upload(h["feature"], golden["dec_feat"])
        for t, s in zip(h["starts"], golden["dec_starts"]):
            m.upload(t, s.astype(np.int32))
        segs = ["setup", "decode"]
    else:
        b, h = engine.build_encode(aec["encoder"], aec["decoders"][0], sd, B, 2048, sample_posterior=True)
        m = ir_exec.Machine(b)
        common.load_encode_inputs(m, h, golden, True)
        segs = ["encode"]
    prog = Program(b)
    prog.set_gemm_backend("auto")
    log, bad = [], []
    for s in segs:
        bad += isolated_records(b, m, prog, *b.segments[s], rtol=TOL["auto"], log=log)
    with open("gpurun_out/records_%s_auto.log" % stage, "w") as f:
        f.write("\n".join(log) + "\n")
    assert lib.load().slide_tc_error() == 0
    assert not bad, bad[:6]
