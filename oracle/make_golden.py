#!/usr/bin/env python
"""Generates the golden fixtures tests/golden/*.npz from the reference's OWN CUDA kernels.

TEST INFRASTRUCTURE ONLY.  Needs a CUDA device and the extension modules built by oracle/build_ref.py
(oracle/_ref/*.so: the unmodified reference sources compiled for sm_100a).  Run on the GPU box:

    gpurun -- 'python oracle/make_golden.py --out gpurun_out/golden'
    cp gpurun_out/golden/*.npz tests/golden/

Each fixture stores the inputs of one case of tests/golden_cases.py and what the reference extension wrote
for them, called exactly as the reference's Python layers call it (zero-filled outputs allocated by the
caller, e.g. FilterInterpolationLayer.py:34,62-64; correlation resizes its own, correlation.py:24-31).
Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "oracle"))

import build_ref  # noqa: E402
import golden_cases as G  # noqa: E402


def main():
    import torch

    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden"))
    ap.add_argument("cases", nargs="*")
    a = ap.parse_args()
    out_dir = Path(a.out)
    out_dir.mkdir(parents=True, exist_ok=True)
    dev = torch.device("cuda", 0)
    mods = {}

    def mod(name):
        if name not in mods:
            mods[name] = build_ref.load(name)
        return mods[name]

    def cu(x):
        return torch.from_numpy(np.ascontiguousarray(x)).to(dev)

    def zeros_like(t):
        return torch.zeros_like(t)

    for name in (a.cases or list(G.CASES)):
        c = G.CASES[name]
        op = c["op"]
        d = G.build_inputs(name)
        t = {k: cu(v) for k, v in d.items()}
        res = {}
        if op.startswith("fi_"):
            m = mod("filterinterpolation_cuda")
            i1, i2, i3, g = t["input1"], t["input2"], t["input3"], t["gradoutput"]
            out, gi1, gi2, gi3 = zeros_like(i1), zeros_like(i1), zeros_like(i2), zeros_like(i3)
            if op == "fi_ori":
                e1 = m.FilterInterpolationLayer_gpu_forward_ori(i1, i2, i3, out)
                e2 = m.FilterInterpolationLayer_gpu_backward_ori(i1, i2, i3, g, gi1, gi2, gi3)
            elif op == "fi_nofilter":
                e1 = m.FilterInterpolationLayer_gpu_forward_nofilterwithdeforconv(i1, i2, i3, out)
                e2 = m.FilterInterpolationLayer_gpu_backward_nofilterwithdeforconv(i1, i2, i3, g, gi1, gi2, gi3)
            else:
                i4 = t["input4"]
                gi4 = zeros_like(i4)
                fwd = m.FilterInterpolationLayer_gpu_forward if op == "fi_dkr" else m.FilterInterpolationLayer_gpu_forward_deforconv
                bwd = m.FilterInterpolationLayer_gpu_backward if op == "fi_dkr" else m.FilterInterpolationLayer_gpu_backward_deforconv
                e1 = fwd(i1, i2, i3, i4, out)
                e2 = bwd(i1, i2, i3, i4, g, gi1, gi2, gi3, gi4)
                res["gi4"] = gi4
            assert e1 == 0 and e2 == 0, (name, e1, e2)
            res.update(out=out, gi1=gi1, gi2=gi2, gi3=gi3)
            if "neg_offsets" in c and not c["neg_offsets"]:
                d["contract_mask"] = G.contract_mask(d["input2"], c["H"], c["W"])
        elif op == "flowproj":
            m = mod("flowprojection_cuda")
            i1, g = t["input1"], t["gradoutput"]
            B, _, H, W = i1.shape
            for fill, key in ((0, "out"), (1, "out_fill")):
                count = torch.zeros(B, 1, H, W, device=dev)
                out = zeros_like(i1)
                assert m.FlowProjectionLayer_gpu_forward(i1, count, out, fill) == 0
                res[key] = out
                res["count" if fill == 0 else "count_fill"] = count
            gi1 = zeros_like(i1)
            assert m.FlowProjectionLayer_gpu_backward(i1, res["count"], g, gi1) == 0
            res["gi1"] = gi1
        elif op == "depthflowproj":
            m = mod("depthflowprojection_cuda")
            i1, i2, g = t["input1"], t["input2"], t["gradoutput"]
            B, _, H, W = i1.shape
            for fill, key in ((0, "out"), (1, "out_fill")):
                count = torch.zeros(B, 1, H, W, device=dev)
                out = zeros_like(i1)
                assert m.DepthFlowProjectionLayer_gpu_forward(i1, i2, count, out, fill) == 0
                res[key] = out
                res["count" if fill == 0 else "count_fill"] = count
            gi1, gi2 = zeros_like(i1), zeros_like(i2)
            assert m.DepthFlowProjectionLayer_gpu_backward(i1, i2, res["count"], res["out"], g, gi1, gi2) == 0
            res.update(gi1=gi1, gi2=gi2)
        elif op in ("interp", "interpch"):
            m = mod("interpolation_cuda" if op == "interp" else "interpolationch_cuda")
            pre = "InterpolationLayer" if op == "interp" else "InterpolationChLayer"
            i1, i2, g = t["input1"], t["input2"], t["gradoutput"]
            out, gi1, gi2 = zeros_like(i1), zeros_like(i1), zeros_like(i2)
            assert getattr(m, pre + "_gpu_forward")(i1, i2, out) == 0
            assert getattr(m, pre + "_gpu_backward")(i1, i2, g, gi1, gi2) == 0
            res.update(out=out, gi1=gi1, gi2=gi2)
        elif op == "sepconv":
            m = mod("separableconv_cuda")
            i1, i2, i3, g = t["input1"], t["input2"], t["input3"], t["gradoutput"]
            out = zeros_like(g)
            gi1, gi2, gi3 = zeros_like(i1), zeros_like(i2), zeros_like(i3)
            assert m.SeparableConvLayer_gpu_forward(i1, i2, i3, out) == 0
            assert m.SeparableConvLayer_gpu_backward(i1, i2, i3, g, gi1, gi2, gi3) == 0
            res.update(out=out, gi1=gi1, gi2=gi2, gi3=gi3)
        elif op == "sepconvflow":
            m = mod("separableconvflow_cuda")
            i1, i2, i3, g = t["input1"], t["input2"], t["input3"], t["gradoutput"]
            out = zeros_like(g)
            gi1, gi2, gi3 = zeros_like(i1), zeros_like(i2), zeros_like(i3)
            assert m.SeparableConvFlowLayer_gpu_forward(i1, i2, i3, out) == 0
            assert m.SeparableConvFlowLayer_gpu_backward(i1, i2, i3, g, gi1, gi2, gi3) == 0
            res.update(out=out, gi1=gi1, gi2=gi2, gi3=gi3)
        elif op == "corr":
            m = mod("correlation_cuda")
            i1, i2 = t["input1"], t["input2"]
            args = (c["pad"], c["k"], c["md"], c["s1"], c["s2"], 1)
            rb1, rb2, out = i1.new_empty(0), i2.new_empty(0), i1.new_empty(0)
            m.forward(i1, i2, rb1, rb2, out, *args)
            gnp = G.corr_gradoutput(name, tuple(out.shape))
            d["gradoutput"] = gnp
            res["out"] = out
            if c.get("bwd", True):   # the reference backward is only in bounds for stride1 == 1
                rb1, rb2, gi1, gi2 = i1.new_empty(0), i2.new_empty(0), i1.new_empty(0), i2.new_empty(0)
                m.backward(i1, i2, rb1, rb2, cu(gnp), gi1, gi2, *args)
                res.update(gi1=gi1, gi2=gi2)
        else:
            raise KeyError(op)
        torch.cuda.synchronize()
        if c.get("wide"):    # inputs come back from the seeded recipe; large outputs are stored as a seeded sample
            arrays = {"checksum": G.input_checksum(d)}
            for k, v in res.items():
                arr = v.detach().cpu().numpy()
                idx = G.sample_indices(name, k, arr.size)
                arrays[f"ref_{k}"] = arr if idx is None else arr.ravel()[idx]
                if idx is not None:
                    arrays[f"shape_{k}"] = np.asarray(arr.shape, np.int64)
        else:
            arrays = {f"in_{k}": np.asarray(v) for k, v in d.items()}
            arrays.update({f"ref_{k}": v.detach().cpu().numpy() for k, v in res.items()})
        np.savez_compressed(out_dir / f"{name}.npz", **arrays)
        print(f"{name}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in arrays.items() if k.startswith('ref_')))
    print(f"wrote {len(a.cases or G.CASES)} fixtures to {out_dir} "
          f"(device {torch.cuda.get_device_name(0)}, torch {torch.__version__})")


if __name__ == "__main__":
    main()
