"""CPU suite: host logic of the block schedule (restormer_b200_arch.run_block / run_stack) with the kernel wrappers
replaced by recorders -- which launches a stack issues, that the LayerNorm that follows a residual row is requested
from the conv that finishes it with the RIGHT norm's parameters (norm2 of the block, norm1 of the next block, also
across the decoder_level1 -> refinement boundary), and that wide / Res-fusion blocks keep the standalone norm."""
import types

import pytest
import torch

from textualdegremoval_b200.archs import restormer_b200_arch as A


class _Recorder:
    def __init__(self, fuse=True, gdfn=False):
        self.calls = []
        self.fuse = fuse
        self.gdfn = gdfn          # fused GDFN tail (tdr_gdfn_tail) available for the shape?

    def gdfn_tail_ok(self, hid, w_out, Cc):
        return self.gdfn

    def gdfn_tail(self, hid, w9, b9, w_out, Cc, **kw):
        self.calls.append(("gdfn", Cc, kw.get("res1") is not None, kw.get("scale_ptr") is not None))
        return kw.get("out")

    def conv_ln_ok(self, Co):
        return self.fuse and Co <= 96 and Co % 8 == 0

    def rows16(self, B, H, W, C, dev, dt=torch.bfloat16):
        return torch.zeros(B, H, W, C, dtype=dt)

    def rownorm(self, x, mode, w=None, b=None, eps=1e-5, out=None, **kw):
        self.calls.append(("rownorm", w))
        return out if out is not None else torch.zeros(x.shape, dtype=torch.bfloat16)

    def conv_gemm(self, x, w, Co, **kw):
        self.calls.append(("conv", Co, kw.get("ln"), kw.get("res1") is not None))
        B, H, W, _ = x.shape
        o32 = kw.get("out_f32")
        if o32 is None and (kw.get("want") == "f32"):
            o32 = torch.zeros(B, H, W, Co)
        o16 = kw.get("out_bf16")
        if o32 is None and o16 is None:
            o16 = torch.zeros(B, H, W, Co, dtype=torch.bfloat16)
        return o32, o16

    def dwconv3x3(self, x, w, b, gate=0, out=None):
        self.calls.append(("dw", gate))
        B, H, W, C = x.shape
        return out if out is not None else torch.zeros(B, H, W, C // 2 if gate else C, dtype=torch.bfloat16)

    def mdta_weff(self, qkv, C, heads, temp, w_po, **kw):
        self.calls.append(("weff",))
        return torch.zeros(qkv.shape[0], C, C, dtype=torch.bfloat16)


def _prep(C, tag, alpha=None):
    t = lambda name: torch.full((1,), 0.0).new_tensor([hash((tag, name)) % 997], dtype=torch.float32)
    return dict(C=C, heads=1, hp=2 * C, alpha=alpha, dt=torch.float16, ln_mode=1, ln1_w=t("ln1_w"), ln1_b=t("ln1_b"), ln2_w=t("ln2_w"),
                ln2_b=t("ln2_b"), w_qkv=None, b_qkv=None, w_qkv_dw=None, b_qkv_dw=None, temp=None, w_po=None, b_po=None,
                w_in=None, b_in=None, w_dw=None, b_dw=None, w_out=None, b_out=None)


@pytest.fixture
def rec(monkeypatch):
    r = _Recorder()
    monkeypatch.setattr(A, "ops", r)
    return r


def _ln_args(calls):
    return [c[2] for c in calls if c[0] == "conv" and c[2] is not None]


def test_stack_chains_norms_through_the_convs(rec):
    preps = [_prep(96, i) for i in range(3)]
    A.run_stack(torch.zeros(1, 8, 8, 96), preps)
    norms = [c for c in rec.calls if c[0] == "rownorm"]
    assert len(norms) == 1 and norms[0][1] is preps[0]["ln1_w"]          # only the first norm1 runs standalone
    ln = _ln_args(rec.calls)
    assert len(ln) == 5                                                   # 3 x norm2 + 2 x next norm1
    assert [a[1] for a in ln] == [preps[0]["ln2_w"], preps[1]["ln1_w"], preps[1]["ln2_w"], preps[2]["ln1_w"],
                                  preps[2]["ln2_w"]]
    assert all(a[0] == 1 and a[3] == 1e-5 for a in ln)
    assert sum(c[0] == "conv" for c in rec.calls) == 12 and sum(c[0] == "dw" for c in rec.calls) == 6   # 9 launches / block


def test_fused_gdfn_tail_replaces_the_gate_and_project_out_launches(monkeypatch):
    """With tdr_gdfn_tail the block is 7 launches: norm1 -> qkv -> dw -> Gram/fold -> attn.v.proj(+norm2) -> project_in ->
    fused tail; the next block's norm1 runs standalone (the tail emits fp32 rows only)."""
    r = _Recorder(gdfn=True)
    monkeypatch.setattr(A, "ops", r)
    preps = [_prep(96, i) for i in range(3)]
    A.run_stack(torch.zeros(1, 8, 8, 96), preps)
    assert sum(c[0] == "rownorm" for c in r.calls) == 3                    # one norm1 per block
    assert [a[1] for a in _ln_args(r.calls)] == [p["ln2_w"] for p in preps]   # norm2 still comes out of the attn conv
    assert sum(c[0] == "conv" for c in r.calls) == 9 and sum(c[0] == "dw" for c in r.calls) == 3
    assert [c for c in r.calls if c[0] == "gdfn"] == [("gdfn", 96, False, False)] * 3
    r.calls.clear()
    alpha = torch.ones(1)
    A.run_stack(torch.zeros(1, 8, 8, 96), [_prep(96, 0, alpha)])
    assert [c for c in r.calls if c[0] == "gdfn"] == [("gdfn", 96, True, True)]   # Res-fusion epilogue: res1 + alpha


def test_stack_boundary_hands_over_the_normalised_rows(rec):
    dec, ref = [_prep(96, "d0"), _prep(96, "d1")], [_prep(96, "r0")]
    tail = []
    x = torch.zeros(1, 8, 8, 96)
    A.run_stack(x, dec, nxt=ref[0], tail=tail)
    assert tail[0] is not None and _ln_args(rec.calls)[-1][1] is ref[0]["ln1_w"]
    n_before = sum(c[0] == "rownorm" for c in rec.calls)
    A.run_stack(x, ref, xn=tail[0])
    assert sum(c[0] == "rownorm" for c in rec.calls) == n_before            # refinement starts without a norm launch


def test_wide_and_fusion_blocks_keep_the_standalone_norm(rec):
    A.run_stack(torch.zeros(1, 4, 4, 192), [_prep(192, 0), _prep(192, 1)])
    assert sum(c[0] == "rownorm" for c in rec.calls) == 4 and not _ln_args(rec.calls)
    rec.calls.clear()
    alpha = torch.ones(1)
    A.run_stack(torch.zeros(1, 4, 4, 96), [_prep(96, 0, alpha), _prep(96, 1, alpha)])
    assert sum(c[0] == "rownorm" for c in rec.calls) == 4 and not _ln_args(rec.calls)
    assert sum(1 for c in rec.calls if c[0] == "conv" and c[3]) == 2        # the two-residual epilogue (res1) per block


def test_fusion_off_is_the_plain_schedule(monkeypatch):
    r = _Recorder(fuse=False)
    monkeypatch.setattr(A, "ops", r)
    A.run_stack(torch.zeros(1, 8, 8, 48), [_prep(48, 0), _prep(48, 1)])
    assert sum(c[0] == "rownorm" for c in r.calls) == 4 and not _ln_args(r.calls)


# ------------------------------------------------------------------------------------------------ training forward
class _TrainRecorder(_Recorder):
    def conv_gemm(self, x, w, Co, **kw):
        ln = kw.get("ln")
        if ln is not None:
            ln[4].fill_(float(len(self.calls)))            # mark the emitted norm tensor with its producer's position
        return super().conv_gemm(x, w, Co, **kw)

    def dwconv3x3_gated_train(self, x, w, b, gate):
        self.calls.append(("dw_train", gate))
        B, H, W, C = x.shape
        return torch.zeros(B, H, W, C // 2, dtype=torch.bfloat16), torch.zeros(B, H, W, C, dtype=torch.bfloat16)

    def mdta_weff(self, qkv, C, heads, temp, w_po, save=None, **kw):
        if save is not None:
            save.update(shat=None, attn=None, weff=None, weff_t=None)
        return super().mdta_weff(qkv, C, heads, temp, w_po)

    def scale_add(self, x, y, scale_ptr=None, **kw):
        self.calls.append(("scale_add",))
        return torch.zeros_like(x)


def test_training_tape_keeps_the_norms_the_convs_emitted(monkeypatch):
    from textualdegremoval_b200.archs import restormer_train as TR
    r = _TrainRecorder()
    monkeypatch.setattr(A, "ops", r)
    monkeypatch.setattr(TR, "ops", r)
    preps = [dict(_prep(96, i), train=True) for i in range(3)]
    tape = []
    out, span = TR.run_stack_train(torch.zeros(1, 8, 8, 96), preps, [None] * 3, tape)
    assert span == (0, 3) and len(tape) == 3 and out.shape == (1, 8, 8, 96)
    assert sum(c[0] == "rownorm" for c in r.calls) == 1                   # only the first norm1 is a launch
    for i, sv in enumerate(tape):
        assert sv["xn1"] is not sv["xn2"] and sv["xn1"].dtype == torch.bfloat16
        if i:                                                              # block i's xn1 is what block i-1's last conv wrote
            assert sv["xn1"] is not tape[i - 1]["xn2"]
            assert float(sv["xn1"].flatten()[0]) > float(tape[i - 1]["xn2"].flatten()[0])
    ln = _ln_args(r.calls)
    assert [a[1] for a in ln] == [preps[0]["ln2_w"], preps[1]["ln1_w"], preps[1]["ln2_w"], preps[2]["ln1_w"],
                                  preps[2]["ln2_w"]]
    # every tape entry keeps its own tensors (the backward reads them after later blocks ran)
    kept = [id(sv[k]) for sv in tape for k in ("xn1", "xn2")]
    assert len(set(kept)) == len(kept)


def test_every_ops_attribute_the_schedules_use_exists():
    """The kernel schedules run on the GPU only; catch renamed / deleted wrappers (``ops.xyz``) and C-ABI entry points
    (``lib.call("tdr_xyz", ...)`` / ``_call("tdr_xyz", ...)``) here, on the CPU."""
    import ast
    import glob
    import os

    from textualdegremoval_b200 import lib, ops
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "textualdegremoval_b200", "**", "*.py"), recursive=True) + \
        [os.path.join(root, "bench.py"), os.path.join(root, "tests", "gpu_checks.py")]
    missing = []
    for f in files:
        tree = ast.parse(open(f).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id == "ops":
                if not hasattr(ops, node.attr):
                    missing.append(f"{os.path.relpath(f, root)}:{node.lineno} ops.{node.attr}")
            if isinstance(node, ast.Call) and node.args and isinstance(node.args[0], ast.Constant) \
                    and isinstance(node.args[0].value, str) and node.args[0].value.startswith("tdr_"):
                if node.args[0].value not in lib.SIGNATURES:
                    missing.append(f"{os.path.relpath(f, root)}:{node.lineno} {node.args[0].value}")
    assert not missing, "\n".join(missing)
