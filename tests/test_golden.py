"""Golden fixtures: outputs of the reference's OWN, unmodified CUDA extensions (oracle/build_ref.py ->
oracle/_ref/*.so, run on a B200 by oracle/make_golden.py) for the seeded cases of tests/golden_cases.py.

  * not-gpu tests pin the CPU oracle (oracle/vfidkr_oracle.c) against them -- this is what makes the oracle
    more than a restatement: it reproduces what the reference kernels actually wrote on the same inputs;
  * gpu tests compare the sm_100a kernels (through the Python front-end and the C ABI) with the same fixtures.

Tolerances: bit-exact for the integer-valued FlowProjection count; 1e-5 for fp32 results; 1e-4 where the
reference accumulates with atomics (order-dependent rounding).  Error measure: tests/util.py:max_err.
"""
from pathlib import Path

import numpy as np
import pytest

import golden_cases as G
import util as U

GOLDEN = Path(__file__).resolve().parent / "golden"
NAMES = sorted(G.CASES)


def load(name):
    p = GOLDEN / f"{name}.npz"
    if not p.exists():
        pytest.fail(f"golden fixture {p} is missing (regenerate with oracle/make_golden.py on a GPU box)")
    z = np.load(p)
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
    if G.CASES[name].get("wide"):
        # inputs are regenerated from the seeded recipe and held to the stored checksum
        ins = G.build_inputs(name)
        if G.CASES[name]["op"] == "corr":
            c = G.CASES[name]
            ins["gradoutput"] = G.corr_gradoutput(name, (c["B"], 81, c["H"], c["W"]))
        assert np.array_equal(G.input_checksum(ins), z["checksum"]), f"{name}: the seeded recipe no longer reproduces the inputs"
        ins["__shapes__"] = {k[6:]: tuple(z[k]) for k in z.files if k.startswith("shape_")}
    else:
        ins = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    return ins, ref


def tol(op, key):
    return U.RTOL_ATOMIC if key in G.ATOMIC_OUTPUTS[op] else U.RTOL_FWD


def compare(name, got: dict, ref: dict, ins: dict):
    op = G.CASES[name]["op"]
    mask = ins.get("contract_mask")
    checked = 0
    for key, r in ref.items():
        if key not in got:
            continue
        g = np.asarray(got[key], dtype=np.float64)
        r = np.asarray(r, dtype=np.float64)
        if key in ins.get("__shapes__", {}):     # a `wide` case's sampled output
            assert g.shape == ins["__shapes__"][key], (name, key, g.shape)
            g = g.ravel()[G.sample_indices(name, key, g.size)]
        assert np.isfinite(g).all(), f"{name}.{key}: non-finite values"
        if key in G.EXACT_OUTPUTS.get(op, ()):
            assert np.array_equal(g, r), f"{name}.{key}: not bit-exact"
        elif mask is not None and key in ("out", "gi2", "gi3", "gi4"):
            # per-pixel outputs of a DKR case with symmetric offsets: compare inside the contract domain only
            m = np.broadcast_to(mask, r.shape)
            U.assert_close(np.where(m, g, 0.0), np.where(m, r, 0.0), tol(op, key), f"{name}.{key} (in-contract pixels)")
        elif mask is not None and key == "gi1":
            continue   # scatter target: out-of-contract pixels contribute reference-undefined values to neighbours
        else:
            U.assert_close(g, r, tol(op, key), f"{name}.{key}")
        checked += 1
    assert checked >= 1, f"{name}: nothing compared"


def test_fixture_inputs_match_their_recipe():
    """The stored inputs are what tests/golden_cases.py generates from the seed (so the recipe is the source)."""
    for name in NAMES:
        if G.CASES[name].get("wide"):
            continue                  # inputs not stored; load() holds the recipe to the stored checksum
        ins, _ = load(name)
        fresh = G.build_inputs(name)
        for k, v in fresh.items():
            assert np.array_equal(ins[k], v), f"{name}: stored input {k} differs from the seeded recipe"


# ----------------------------------------------------------------------------------------- oracle vs golden
def run_oracle(O, name, ins):
    c = G.CASES[name]
    op = c["op"]
    if op.startswith("fi_"):
        variant = {"fi_ori": "ori", "fi_dkr": "dkr", "fi_deforconv": "deforconv", "fi_nofilter": "nofilterwithdeforconv"}[op]
        i4 = ins.get("input4")
        out = O.fi_forward(variant, ins["input1"], ins["input2"], ins["input3"], i4)
        gi1, gi2, gi3, gi4 = O.fi_backward(variant, ins["input1"], ins["input2"], ins["input3"], i4, ins["gradoutput"])
        got = dict(out=out, gi1=gi1, gi2=gi2, gi3=gi3)
        if gi4 is not None:
            got["gi4"] = gi4
        return got
    if op in ("flowproj", "depthflowproj"):
        d = ins.get("input2")
        out, count = O.flowprojection_forward(ins["input1"], d, 0)
        out_fill, count_fill = O.flowprojection_forward(ins["input1"], d, 1)
        gi1, gi2 = O.flowprojection_backward(ins["input1"], d, count, out if d is not None else None, ins["gradoutput"])
        got = dict(out=out, count=count, out_fill=out_fill, count_fill=count_fill, gi1=gi1)
        if gi2 is not None:
            got["gi2"] = gi2
        return got
    if op in ("interp", "interpch"):
        out = O.interpolation_forward(ins["input1"], ins["input2"])
        gi1, gi2 = O.interpolation_backward(ins["input1"], ins["input2"], ins["gradoutput"])
        return dict(out=out, gi1=gi1, gi2=gi2)
    if op == "sepconv":
        out = O.sepconv_forward(ins["input1"], ins["input2"], ins["input3"])
        gi1, gi2, gi3 = O.sepconv_backward(ins["input1"], ins["input2"], ins["input3"], ins["gradoutput"])
        return dict(out=out, gi1=gi1, gi2=gi2, gi3=gi3)
    if op == "sepconvflow":
        out = O.sepconvflow_forward(ins["input2"], ins["input3"])
        gi2, gi3 = O.sepconvflow_backward(ins["input2"], ins["input3"], ins["gradoutput"])
        return dict(out=out, gi1=np.zeros_like(ins["input1"], dtype=np.float64), gi2=gi2, gi3=gi3)
    if op == "corr":
        a = (c["pad"], c["k"], c["md"], c["s1"], c["s2"])
        out = O.correlation_forward(ins["input1"], ins["input2"], *a)
        if not c.get("bwd", True):
            return dict(out=out)
        gi1, gi2 = O.correlation_backward(ins["input1"], ins["input2"], ins["gradoutput"], *a)
        return dict(out=out, gi1=gi1, gi2=gi2)
    raise KeyError(op)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_kernels(oracle, name):
    ins, ref = load(name)
    compare(name, run_oracle(oracle, name, ins), ref, ins)


# ----------------------------------------------------------------------------------------- CUDA path vs golden
def run_cuda(V, name, ins):
    import torch

    def cu(a, grad=False):
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        return t.requires_grad_() if grad else t

    def host(t):
        return t.detach().cpu().numpy()

    c = G.CASES[name]
    op = c["op"]
    g = cu(ins["gradoutput"])
    if op.startswith("fi_"):
        t1, t2, t3 = cu(ins["input1"], True), cu(ins["input2"], True), cu(ins["input3"], True)
        if op == "fi_ori":
            out = V.FilterInterpolationModule()(t1, t2, t3)
        elif op == "fi_nofilter":
            out = V.FilterInterpolationModule("nofilterwithdeforconv")(t1, t2, t3)
        else:
            t4 = cu(ins["input4"], True)
            out = V.FilterInterpolationModule("dkr" if op == "fi_dkr" else "deforconv")(t1, t2, t3, t4)
        out.backward(g)
        got = dict(out=host(out), gi1=host(t1.grad), gi2=host(t2.grad), gi3=host(t3.grad))
        if op in ("fi_dkr", "fi_deforconv"):
            got["gi4"] = host(t4.grad)
        return got
    if op in ("flowproj", "depthflowproj"):
        from vfidkr_b200 import _lib
        from vfidkr_b200._common import ptr, stream_ptr
        t1 = cu(ins["input1"], True)
        B, _, H, W = t1.shape
        sp = stream_ptr(t1.device)
        got = {}
        if op == "flowproj":
            out = V.FlowProjectionModule(True)(t1)
            out.backward(g)
            got.update(out=host(out), gi1=host(t1.grad))
            got["out_fill"] = host(V.FlowProjectionModule(False)(t1.detach()))
            cnt, o = torch.empty(B, 1, H, W, device="cuda"), torch.empty(B, 2, H, W, device="cuda")
            for fill, key in ((0, "count"), (1, "count_fill")):
                _lib.call("vfidkr_flowprojection_forward", ptr(t1), ptr(cnt), ptr(o), B, H, W, fill, sp)
                got[key] = host(cnt)
        else:
            t2 = cu(ins["input2"], True)
            out = V.DepthFlowProjectionModule(True)(t1, t2)
            out.backward(g)
            got.update(out=host(out), gi1=host(t1.grad), gi2=host(t2.grad))
            got["out_fill"] = host(V.DepthFlowProjectionModule(False)(t1.detach(), t2.detach()))
            cnt, o = torch.empty(B, 1, H, W, device="cuda"), torch.empty(B, 2, H, W, device="cuda")
            for fill, key in ((0, "count"), (1, "count_fill")):
                _lib.call("vfidkr_depthflowprojection_forward", ptr(t1), ptr(t2), ptr(cnt), ptr(o), B, H, W, fill, sp)
                got[key] = host(cnt)
        return got
    if op in ("interp", "interpch"):
        t1, t2 = cu(ins["input1"], True), cu(ins["input2"], True)
        out = (V.InterpolationModule() if op == "interp" else V.InterpolationChModule(c["C"]))(t1, t2)
        out.backward(g)
        return dict(out=host(out), gi1=host(t1.grad), gi2=host(t2.grad))
    if op in ("sepconv", "sepconvflow"):
        t1, t2, t3 = cu(ins["input1"], True), cu(ins["input2"], True), cu(ins["input3"], True)
        mod = V.SeparableConvModule(c["F"]) if op == "sepconv" else V.SeparableConvFlowModule(c["F"])
        out = mod(t1, t2, t3)
        out.backward(g)
        gi1 = host(t1.grad) if t1.grad is not None else np.zeros_like(ins["input1"])
        return dict(out=host(out), gi1=gi1, gi2=host(t2.grad), gi3=host(t3.grad))
    if op == "corr":
        t1, t2 = cu(ins["input1"], True), cu(ins["input2"], True)
        out = V.Correlation(c["pad"], c["k"], c["md"], c["s1"], c["s2"], 1)(t1, t2)
        if not c.get("bwd", True):
            return dict(out=host(out))
        out.backward(g)
        return dict(out=host(out), gi1=host(t1.grad), gi2=host(t2.grad))
    raise KeyError(op)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_reproduces_reference_kernels(lib, name):
    ins, ref = load(name)
    before = lib.launch_count()
    got = run_cuda(lib, name, ins)
    assert lib.launch_count() > before
    compare(name, got, ref, ins)
