"""Host-side streaming of independent frame pairs through the hot path.

Frame pairs are independent, so a batch that starts in HOST memory need not be copied as a whole
before the first kernel runs.  `PairStream` cuts the batch into chunks of pairs and runs three CUDA streams:
host -> device copies of chunk i + 1, the operators on chunk i and device -> host copies of the results of chunk
i - 1 overlap, ordered by events.  With the 1080p workload (495 MB of inputs per pair, 0.3 ms of kernels) the
end-to-end time becomes the PCIe host -> device time alone.  The reference has nothing of the kind: it runs one
stream with a device-wide synchronise in the middle of the network (networks/DAIN.py:119-120, :213).

PyTorch supplies streams, events and device memory; the operators are called exactly as a user calls them.
"""
from __future__ import annotations

from typing import Callable, Mapping, Sequence

import torch


def bind_to_gpu_numa_node(device) -> list[int] | None:
    """Pin the calling process to the CPUs NVML reports as local to `device` (its NUMA node), so that pinned host buffers
    allocated AFTERWARDS are first-touched there and host -> device copies do not cross the socket interconnect.
    Returns the CPU list, or None when NVML / the affinity call is unavailable (nothing is changed then).  Host-side
    plumbing for the end-to-end path of several ranks on one box; no reference counterpart."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device)
        bus = f"{props.pci_domain_id:08X}:{props.pci_bus_id:02X}:{props.pci_device_id:02X}.0"
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (int(words[i // 64]) >> (i % 64)) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


class PairStream:
    """step_fn(chunk_inputs: dict[str, Tensor on device]) -> sequence of device tensors with the pair dimension first.

    run(host_inputs, host_outputs): every tensor of `host_inputs` is pinned host memory with the pair dimension first;
    result j of step_fn for pairs [i, i + n) is copied into host_outputs[j][i:i + n] (pinned host memory).
    The call returns when everything is enqueued; `synchronize()` waits for the results."""

    def __init__(self, device: torch.device, step_fn: Callable[[Mapping[str, torch.Tensor]], Sequence[torch.Tensor]],
                 pairs_per_chunk: int = 1):
        if device.type != "cuda":
            raise ValueError("PairStream needs a CUDA device (there is no CPU path)")
        self.device, self.step_fn, self.chunk = device, step_fn, max(1, int(pairs_per_chunk))
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(device) for _ in range(3))

    def run(self, host_inputs: Mapping[str, torch.Tensor], host_outputs: Sequence[torch.Tensor]) -> None:
        n = next(iter(host_inputs.values())).shape[0]
        for k, t in host_inputs.items():
            if t.shape[0] != n or not t.is_pinned():
                raise ValueError(f"input {k!r}: pinned host tensor with {n} pairs expected")
        caller = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event()
        start.record(caller)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_event(start)           # nothing of this batch overtakes work the caller has already enqueued
        for i in range(0, n, self.chunk):
            j = min(i + self.chunk, n)
            with torch.cuda.stream(self.s_in):
                dev = {k: t[i:j].to(self.device, non_blocking=True) for k, t in host_inputs.items()}
                copied = torch.cuda.Event()
                copied.record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(copied)
                for t in dev.values():
                    t.record_stream(self.s_run)      # allocated on s_in, consumed here
                results = self.step_fn(dev)
                done = torch.cuda.Event()
                done.record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                for dst, r in zip(host_outputs, results):
                    r.record_stream(self.s_out)
                    dst[i:j].copy_(r, non_blocking=True)
        finished = torch.cuda.Event()
        finished.record(self.s_out)
        caller.wait_event(finished)       # work the caller enqueues next sees the results

    def synchronize(self) -> None:
        self.s_out.synchronize()
