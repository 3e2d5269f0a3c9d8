"""Multi-GPU plumbing of the update_packets() path: one process per GPU, packets sharded, estimators summed.

The reference gives every MPI rank its own MPKTS packets and, after update_packets(), sums the estimator arrays
over ranks with MPI_Allreduce (sn3d.cc:565-631, radfield.cc:988-1030). Here the only exchange is ONE all-reduce
(sum, f64) of the library's packed estimator buffer [J|nuJ|ffheating|colheating|gamma|bfheating|dep_*|ts.scalars]
plus the integer counters; packets never move between ranks and all tables are replicated.
torch.distributed is plumbing only (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
import numpy as np

ESTIMATOR_ORDER = ["est.J", "est.nuJ", "est.ffheating", "est.colheating", "est.gamma", "est.bfheating", "est.dep_gamma",
                   "est.dep_positron", "est.dep_electron", "est.dep_alpha", "ts.scalars", "est.bins_J_raw", "est.bins_nuJ_raw", "est.bfrate_raw"]


def rank_seed(base_seed, rank):
    """Philox key word of a rank: distinct streams per rank (the reference offsets pre_zseed by rank, input.cc:1911-1916)"""
    return int(base_seed) + (int(rank) << 32)


def shard_bounds(npackets_total, rank, world):
    """contiguous shard [begin, end) of a global packet array; sizes differ by at most one packet"""
    base, extra = divmod(int(npackets_total), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def device_estimator_tensor(eng, device_index):
    """the library's packed estimator buffer as a torch tensor (no copy)"""
    import torch

    ptr, count = eng.estimator_device_buffer()

    class _Holder:
        def __init__(self, p, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (p, False), "version": 3}

    return torch.as_tensor(_Holder(ptr, count), device=torch.device("cuda", device_index))


def allreduce_estimators_device(eng, device_index, group=None):
    """in-place NCCL all-reduce of the packed device buffer (on the caller's current stream)"""
    import torch.distributed as dist

    t = device_estimator_tensor(eng, device_index)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allreduce_estimators_host(est, group=None):
    """Sum a dict of host estimator arrays (ArtisB200.estimators()) over ranks with one packed all-reduce for the f64
    arrays and one for the integer counters. Works with any torch.distributed backend that reduces CPU tensors."""
    import torch
    import torch.distributed as dist

    names = [n for n in ESTIMATOR_ORDER if n in est]
    packed = torch.from_numpy(np.concatenate([np.asarray(est[n], dtype=np.float64).ravel() for n in names]))
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    out = dict(est)
    off = 0
    flat = packed.numpy()
    for n in names:
        size = np.asarray(est[n]).size
        out[n] = flat[off:off + size].reshape(np.asarray(est[n]).shape).copy()
        off += size
    ints = [n for n in ("counters", "ts.pellet_decays", "diag") if n in est]
    if ints:
        packed_i = torch.from_numpy(np.concatenate([np.asarray(est[n], dtype=np.int64).ravel() for n in ints]))
        dist.all_reduce(packed_i, op=dist.ReduceOp.SUM, group=group)
        off = 0
        flat_i = packed_i.numpy()
        for n in ints:
            size = np.asarray(est[n]).size
            out[n] = flat_i[off:off + size].copy()
            off += size
    return out


SPECTRA_SUMMED = ["flux", "emission", "trueemission", "absorption", "lc_lum", "lc_lumcmf", "gamma_lc_lum", "gamma_lc_lumcmf"]


def allreduce_binned_host(binned, group=None):
    """Sum the arrays of artis_b200.spectra.binned() over ranks with ONE packed f64 all-reduce: every rank bins the escaped
    packets it owns (artisb200_bin_escaped_packets with nprocs_exspec = number of ranks), the sums are the spectra and light
    curves of the run - the reference's MPI_Allreduce calls at spectrum_lightcurve.cc:293-310. The frequency grid and the
    per-packet direction bins are rank-local and stay as they are."""
    import torch
    import torch.distributed as dist

    names = [n for n in SPECTRA_SUMMED if n in binned]
    packed = torch.from_numpy(np.concatenate([np.asarray(binned[n], dtype=np.float64).ravel() for n in names]))
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    out = dict(binned)
    flat = packed.numpy()
    off = 0
    for n in names:
        shape = np.asarray(binned[n]).shape
        size = int(np.prod(shape))
        out[n] = flat[off:off + size].reshape(shape).copy()
        off += size
    return out
