"""Multi-GPU host logic: streams shard across ranks, only per-stream result records are exchanged.

Every dongle stream (gsm_sync_demod.m:112) / scanned channel (multi_rtl_sdr_gsm_FCCH_scanner.m:163) is
independent, so there is no data-path collective; one all_gather of fixed-size records (88 B per stream)
over NCCL/NVLink (gloo in the CPU tests) assembles the whole-box result on every rank.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ._lib import StreamResult

REC_BYTES = C.sizeof(StreamResult)


def shard_range(n_streams: int, rank: int, world: int):
    """Contiguous block of streams owned by `rank` (first n_streams % world ranks get one extra)."""
    base, extra = divmod(n_streams, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def records_to_tensor(records, device="cpu") -> torch.Tensor:
    buf = np.frombuffer(bytes(records), dtype=np.uint8).copy()
    return torch.from_numpy(buf).to(device)


def tensor_to_records(t: torch.Tensor, n: int):
    raw = t.cpu().numpy().tobytes()[: n * REC_BYTES]
    return (StreamResult * n).from_buffer_copy(raw)


def gather_records(local_records, n_streams: int, device="cpu", group=None):
    """all_gather of the per-stream result records; returns a (StreamResult * n_streams) array in stream order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_records
    rank = dist.get_rank(group)
    sizes = [shard_range(n_streams, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    mine = torch.zeros(cap * REC_BYTES, dtype=torch.uint8, device=device)
    t = records_to_tensor(local_records, device)
    mine[: t.numel()] = t
    out = torch.empty(world * cap * REC_BYTES, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    out = out.cpu().numpy()
    merged = bytearray()
    for r, (lo, hi) in enumerate(sizes):
        merged += out[r * cap * REC_BYTES: r * cap * REC_BYTES + (hi - lo) * REC_BYTES].tobytes()
    assert sizes[rank][1] - sizes[rank][0] == len(local_records)
    return (StreamResult * n_streams).from_buffer_copy(bytes(merged))


def calibrate_sharded(raw_all, carrier_freq, template, coef, compute, device="cpu", group=None):
    """Runs `compute(raw_shard) -> (StreamResult * n)` on this rank's block of streams and gathers the records.

    `compute` is gsmcal.calibrate_batch(..., details=False) on a GPU box."""
    n = raw_all.shape[0]
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_range(n, rank, world)
    local = compute(raw_all[lo:hi], carrier_freq, template, coef)
    return gather_records(local, n, device, group)
