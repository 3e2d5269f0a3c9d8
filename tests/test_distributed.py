"""Multi-rank path on CPU: world_size-2 gloo processes, each propagating its shard of the packets through the
host-simulation build (TEST-ONLY), estimators summed with artis_b200.distributed -- must reproduce the single-process
run: the same packets (a packet's history depends on nothing but the packet), the same event counters, and estimator
sums equal up to the order of the floating-point additions. On GPUs the same helpers run over NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from artis_b200 import distributed as abdist
from tests import fixtures

CONFIG, NTS = "kilonova_toy", 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_shard(lib, fx, begin, end, seed):
    """propagate packets [begin, end) of the fixture; returns (packets as structured array, estimators)"""
    before = dict(fx["before"])
    stride = int(before["packets.stride"][0])
    before["packets.aos"] = before["packets.aos"][begin * stride:end * stride].copy()
    before["packets.count"] = np.array([end - begin], dtype=before["packets.count"].dtype)
    pk, est, _, _ = fixtures.run_fixture(lib, dict(fx, before=before), rng="philox", seed=seed, options={"schedule": 1, "wf_tail": 0})
    return pk, est


def _worker(rank, world, port, lib, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ARTISB200_ALLOW_HOSTSIM="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fx = fixtures.load_golden(CONFIG, NTS)
        n = int(fx["before"]["packets.count"][0])
        begin, end = abdist.shard_bounds(n, rank, world)
        # same Philox seed on every rank here: the key also holds the packet number, and the check below compares with
        # a single-process run of all packets (production uses abdist.rank_seed, every rank owning its own packets)
        pk, est = _run_shard(lib, fx, begin, end, seed=77)
        total = abdist.allreduce_estimators_host(est)
        np.save(os.path.join(outdir, f"pk{rank}.npy"), pk)
        if rank == 0:
            np.savez(os.path.join(outdir, "est.npz"), **total)
    finally:
        dist.destroy_process_group()


def test_two_ranks_reproduce_one(tmp_path):
    lib = fixtures.hostsim_library(fixtures.PRESET_OF[CONFIG])
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), lib, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    fx = fixtures.load_golden(CONFIG, NTS)
    n = int(fx["before"]["packets.count"][0])
    pk_ref, est_ref = _run_shard(lib, fx, 0, n, seed=77)
    pk = np.concatenate([np.load(tmp_path / f"pk{r}.npy") for r in range(world)])
    for name in pk_ref.dtype.names:  # field by field: np.save does not keep the bytes of NaN payloads / struct padding
        x, y = pk[name], pk_ref[name]
        assert np.array_equal(x, y, equal_nan=(x.dtype.kind == "f")), f"sharded packets differ from the single-process run in {name}"
    est = np.load(tmp_path / "est.npz")
    assert np.array_equal(est["counters"][:32], est_ref["counters"][:32])  # UPDATECELL (32) counts per-rank table builds
    assert int(est["counters"][33]) == int(est_ref["counters"][33])
    assert int(est["ts.pellet_decays"][0]) == int(est_ref["ts.pellet_decays"][0])
    for name in abdist.ESTIMATOR_ORDER:
        if name not in est_ref:  # optional estimators of other presets
            continue
        a, b = est[name], est_ref[name]
        scale = max(np.abs(b).max(), 1e-300)
        assert np.abs(a - b).max() / scale < 1e-12, name


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 1000, 10_000_001):
        for world in (1, 2, 3, 8):
            bounds = [abdist.shard_bounds(n, r, world) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in bounds]
            assert max(sizes) - min(sizes) <= 1


def test_rank_seeds_are_distinct():
    seeds = {abdist.rank_seed(20260101, r) for r in range(8)}
    assert len(seeds) == 8


def _spectra_worker(rank, world, port, lib, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ARTISB200_ALLOW_HOSTSIM="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from artis_b200 import spectra as spectra_mod
        fx = fixtures.load_golden(CONFIG, NTS)
        after = fx["after"]
        n, stride = int(after["packets.count"][0]), int(after["packets.stride"][0])
        begin, end = abdist.shard_bounds(n, rank, world)
        eng = fixtures.ablib.ArtisB200(libpath=lib)
        eng.set_arrays(fx["static"])
        eng.commit_static()
        eng.upload_packets(after["packets.aos"][begin * stride:end * stride].copy(), end - begin, stride)
        eng.bin_escaped_packets(direction_bins=True, emission_absorption=1, nprocs_exspec=1)
        total = abdist.allreduce_binned_host(spectra_mod.binned(eng))
        eng.close()
        if rank == 0:
            np.savez(os.path.join(outdir, "binned.npz"), **total)
    finally:
        dist.destroy_process_group()


def test_two_ranks_bin_the_spectra_of_one(tmp_path):
    # every rank bins the escaped packets it owns, one packed all-reduce sums the sets (spectrum_lightcurve.cc:293-310): the
    # result is the reference's own binning of all packets (tests/golden/kilonova_toy_spectra_ts4.npz)
    lib = fixtures.hostsim_library(fixtures.PRESET_OF[CONFIG])
    world = 2
    mp.start_processes(_spectra_worker, args=(world, _free_port(), lib, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    got = np.load(tmp_path / "binned.npz")
    ref = np.load(os.path.join(fixtures.GOLDEN_DIR, f"{CONFIG}_spectra_ts{NTS}.npz"))
    for key, sel, refkey in (("flux", 0, "ref.spec.flux"), ("emission", 0, "ref.spec.emission"), ("absorption", 0, "ref.spec.absorption"),
                             ("lc_lum", 0, "ref.lc.lum"), ("lc_lumcmf", 0, "ref.lc.lumcmf"), ("gamma_lc_lum", None, "ref.lc.gamma_lum"),
                             ("flux", slice(1, None), "ref.spec.flux_res"), ("lc_lum", slice(1, None), "ref.lc.lum_res")):
        a = (got[key] if sel is None else got[key][sel]).ravel()
        b = ref[refkey].ravel()
        assert np.array_equal(a != 0., b != 0.), key
        assert (np.abs(a - b) / np.maximum(np.abs(b), 1e-300)).max() <= 1e-12, key
