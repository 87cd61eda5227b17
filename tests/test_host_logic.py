"""Host-side behaviour that needs no GPU: argument checks, error behaviour, drop-in aliases, sharding."""
import os
import sys

import numpy as np
import pytest
import torch


def test_cpu_tensors_are_rejected_no_fallback(lib):
    x = torch.rand(1, 3, 8, 8)
    f = torch.zeros(1, 2, 8, 8)
    w = torch.rand(1, 16, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lib.FilterInterpolationModule()(x, f, w)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lib.FlowProjectionModule(False)(f)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lib.Correlation(4, 1, 4, 1, 1, 1)(x, x)


def test_missing_library_fails_loudly(lib, monkeypatch, tmp_path):
    from vfidkr_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_c_abi_rejects_bad_arguments_without_touching_the_gpu(lib):
    from ctypes import c_void_p
    from vfidkr_b200 import _lib
    null = c_void_p(0)
    # null pointers / non-positive sizes are rejected before any CUDA call is made
    with pytest.raises(_lib.VfidkrError, match="status 1"):
        _lib.call("vfidkr_filterinterpolation_forward_ori", null, null, null, null, 1, 3, 8, 8, 4, null)
    with pytest.raises(_lib.VfidkrError, match="status 1"):
        _lib.call("vfidkr_flowprojection_forward", null, null, null, 0, 8, 8, 0, null)
    with pytest.raises(_lib.VfidkrError, match="status 1"):
        _lib.call("vfidkr_interpolation_forward", c_void_p(16), c_void_p(16), c_void_p(16), 1, 4, 8, 8, 1, null)  # C != 3


def test_reference_import_lines_resolve_to_this_package(lib):
    """The four operator import lines of the reference's networks (more in tests/test_dropin_network.py, which
    imports the reference's real networks through the aliases)."""
    names = lib.install_reference_aliases()
    try:
        assert "my_package.FilterInterpolation" in names
        from my_package.FilterInterpolation import FilterInterpolationModule          # networks/DAIN.py:11
        from my_package.FlowProjection import FlowProjectionModule                    # networks/DAIN.py:12
        from my_package.DepthFlowProjection import DepthFlowProjectionModule          # networks/DAIN.py:13
        from PWCNet.correlation_package_pytorch1_0.correlation import Correlation     # PWCNet/PWCNet.py:15
        assert FilterInterpolationModule is lib.FilterInterpolationModule
        assert FlowProjectionModule is lib.FlowProjectionModule
        assert DepthFlowProjectionModule is lib.DepthFlowProjectionModule
        assert Correlation is lib.Correlation
    finally:
        lib.compat.remove_reference_aliases()
        for k in [k for k in sys.modules if k.startswith("PWCNet")]:
            del sys.modules[k]


def test_module_signatures_match_the_reference(lib):
    import inspect
    assert list(inspect.signature(lib.FilterInterpolationModule().forward).parameters) == ["input1", "input2", "input3", "input4"]
    assert list(inspect.signature(lib.FlowProjectionModule.__init__).parameters) == ["self", "requires_grad"]
    assert list(inspect.signature(lib.DepthFlowProjectionModule(True).forward).parameters) == ["input1", "input2"]
    assert list(inspect.signature(lib.InterpolationChModule.__init__).parameters) == ["self", "ch"]
    assert list(inspect.signature(lib.SeparableConvModule.__init__).parameters) == ["self", "filtersize"]
    assert list(inspect.signature(lib.Correlation.__init__).parameters) == [
        "self", "pad_size", "kernel_size", "max_displacement", "stride1", "stride2", "corr_multiply"]


def test_pair_sharding_is_a_partition():
    import bench
    for n_pairs in (1, 7, 8, 13):
        for world in (1, 2, 4, 8):
            parts = [bench.shard_pairs(n_pairs, r, world) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n_pairs))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    # every rank "times" its own shard; the reported time is the max over ranks, units are summed
    t = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
    units = torch.tensor([float(len(bench.shard_pairs(9, rank, world)))], dtype=torch.float64)
    tmax, usum = bench.reduce_timing(t, units)
    q.put((rank, float(tmax), float(usum)))
    dist.destroy_process_group()


def test_world_size_2_timing_reduction_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, 15.0, 9.0), (1, 15.0, 9.0)]


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    # the 4k_stream bookkeeping: P pairs dealt round-robin, every round each rank contributes the frame of its pair (a
    # blank when it has none left), rank 0 receives them at slot round * world + rank = the pair's own index
    P, shape = 5, (1, 3, 4, 6)
    mine = bench.shard_pairs(P, rank, world)
    rounds = (P + world - 1) // world
    out = torch.full((rounds * world,) + shape[1:], -1.0) if rank == 0 else None
    for r in range(rounds):
        frame = torch.full(shape, float(mine[r])) if r < len(mine) else torch.zeros(shape)
        dst = [out[r * world + k].unsqueeze(0) for k in range(world)] if rank == 0 else [None]
        w = bench.gather_frames(dist, frame, dst, rank, world, async_op=True)
        if w is not None:
            w.wait()
    if rank == 0:
        q.put([float(out[i].mean()) for i in range(rounds * world)])
    dist.destroy_process_group()


def test_world_size_2_frame_gather_over_gloo():
    """bench.py --workload 4k_stream: pair i ends up at slot i of rank 0's output (blank slots past the stream)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + ((os.getpid() + 977) % 2000)
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [0.0, 1.0, 2.0, 3.0, 4.0, 0.0]


def test_both_bench_arms_print_the_same_config():
    """`same_config` was false because the two arms spelt the workload differently."""
    import bench
    assert bench.workload_config(1) == bench.workload_config(1)
    src = open(bench.__file__).read()
    assert src.count('"config": workload_config(') == 3           # our arm, reference (CUDA), reference (CPU port)


def test_pair_stream_refuses_cpu():
    """No CPU path: the host-side streamer insists on a CUDA device."""
    import pytest
    import torch
    import vfidkr_b200 as V
    with pytest.raises(ValueError):
        V.PairStream(torch.device("cpu"), lambda d: ())


def test_frame_padding_rule():
    """demo_MiddleBury.py:286-301: next multiple of 128, split floor/ceil; + 32 + 32 when already a multiple."""
    import vfidkr_b200 as V
    assert V.frame_padding(1080) == (36, 1152)      # (1152 - 1080) / 2
    assert V.frame_padding(1920) == (32, 1984)      # already a multiple of 128
    assert V.frame_padding(2160) == (8, 2176)
    assert V.frame_padding(3840) == (32, 3904)
    assert V.frame_padding(480) == (16, 512) and V.frame_padding(640) == (32, 704)   # the demo's own comment: (512, 704)
    assert V.frame_padding(1) == (63, 128) and V.frame_padding(129) == (63, 256)
