#!/usr/bin/env python3
"""Benchmark of the update_packets() hot path (BASELINE.json metric: packet-interactions/sec per timestep).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm: the CUDA library through its C ABI
  python bench.py --impl reference [--gpus N] --steps K --warmup W   the reference's own CPU update_packets

Workload (config.workload): BASELINE.json configs[1], the kilonova LTE 2D cylindrical model with 1e7 packets, one
timestep of it. The state at the start of that timestep (static tables, cell state, 1e7 packets) is produced on the
box by the drop-in host binary (the reference's own driver and grid update with update_packets() bound to this
library, integration/), saved as a snapshot, and every "step" replays update_packets() for that timestep:
begin_timestep (zero estimators + per-cell table build, which the reference's GPU_ON update_packets also does up
front, update_packets.cc:551-563) followed by the propagation of all packets to the end of the timestep.

  value      interactions / device time with the packets already resident in HBM (CUDA events on the library stream)
  e2e        the same through artisb200_update_packets_host with pinned HOST packet buffers: H2D of the AoS packets,
             propagation, D2H of the packets and of the estimators inside the timed region
  roofline   algorithmic bytes (SURVEY.md 8d formula, from the device work counters) / propagation-kernel time
  cpu_baseline  the compiled reference (oracle/_ref, production flags) timed on this box on a bounded packet sample
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

# --workload: BASELINE.json configs. The bench line the driver reads is the default one (configs[1], the configuration the
# metric is quoted on); the others are the same measurement on the other configurations (lines kept under profiles/).
WORKLOADS = {
    "kilonova_2d": dict(
        preset="kilonova_lte", ts=2, cpu_config="kilonova_2d_cpu",
        workload="kilonova LTE 2D cylindrical r-process ejecta (BASELINE configs[1])", model_grid="2D cylindrical 50 x 100",
        atomic_data="synthetic, 5 elements x 4 ions x 120 levels (54 892 lines, 1 475 bound-free continua)"),
    "classic_1d3d": dict(
        preset="classic", ts=4, cpu_config="classic_1d3d_cpu",
        # few active packets (~1e4) with ~3e5 interactions each in this optically thick phase: the whole-history kernel with
        # the packets spread over the warps takes over early
        options={"wf_tail": 16384},
        workload="classic LTE W7-like 1D model on a 3D Cartesian 100^3 grid, 1e5 packets (BASELINE configs[0])",
        model_grid="1D model, 100 shells, on a 3D Cartesian 100^3 propagation grid",
        atomic_data="synthetic, 7 elements x 4 ions x <= 40 levels (2 002 lines)"),
    "asym3d": dict(
        # timestep 2: the last one in which cells with grey optical depth > 8 are treated grey (input.txt line 19). From
        # timestep 3 on the optically thick interior gets the detailed treatment and an active packet takes ~2e5 interactions
        # per timestep one after the other (measured with the reference: 71 s per 1e5 packets and timestep on one core);
        # a step of that phase does not fit a bench run on any hardware
        preset="classic", ts=2, cpu_config="asym3d_cpu",
        # (the step ends with ~2e4 packets random-walking through very thick grey cells, 1e4-1e5 scatterings each in sequence;
        # keeping them in the wavefront down to 2048 packets was measured slower: 9.8 s against 4.5 s)
        workload="3D Cartesian asymmetric SN Ia model, classic macro-atom mode (BASELINE configs[2]; 1e6 packets per run)",
        model_grid="3D Cartesian 100^3 (ellipsoidal density with an off-centre Ni blob)",
        atomic_data="synthetic, 7 elements x 4 ions x <= 40 levels (2 002 lines)"),
    "gamma_3d50": dict(
        preset="classic", ts=1, cpu_config="gamma_3d50_cpu",
        workload="gamma-packet-only Ni56/Co56 deposition run (Compton/photoelectric/pair) on a 3D 50^3 grid from 20 d, every cell "
                 "grey for r-packets (BASELINE configs[3]; 1e7 packets per run)",
        model_grid="3D Cartesian 50^3", atomic_data="synthetic, 3 elements x 3 ions x 6 levels (93 lines)"),
}
WORKLOAD = os.environ.get("ARTISB200_BENCH_CONFIG", "kilonova_2d")
PRESET = "kilonova_lte"
BENCH_TS = int(os.environ.get("ARTISB200_BENCH_TS", "2"))
CPU_SAMPLE_CONFIG = os.environ.get("ARTISB200_BENCH_CPU_CONFIG", "kilonova_2d_cpu")


def select_workload(name):
    global WORKLOAD, PRESET, BENCH_TS, CPU_SAMPLE_CONFIG
    if name not in WORKLOADS:
        raise SystemExit(f"unknown workload {name!r}; known: {sorted(WORKLOADS)}")
    w = WORKLOADS[name]
    WORKLOAD = name
    PRESET = w["preset"]
    BENCH_TS = int(os.environ.get("ARTISB200_BENCH_TS", str(w["ts"])))
    CPU_SAMPLE_CONFIG = os.environ.get("ARTISB200_BENCH_CPU_CONFIG", w["cpu_config"])


def make_rundir(config, builddir, rundir):
    """a fresh run folder of `config`: the input files written at build time (builddir/inputs) or, for the models too large
    to keep in the tree, generated here once; large read-only files are linked, input.txt is copied (it gets edited)"""
    import gen_inputs
    inputs = os.path.join(builddir, "inputs")
    if not os.path.isdir(inputs):
        inputs = os.path.join(CACHE, f"inputs_{config}")
        if not os.path.isdir(inputs):
            log(f"writing the input files of {config} ...")
            gen_inputs.generate(config, inputs + ".tmp")
            os.replace(inputs + ".tmp", inputs)
    if os.path.isdir(rundir):
        shutil.rmtree(rundir)
    os.makedirs(rundir)
    for f in os.listdir(inputs):
        src = os.path.join(inputs, f)
        if f == "data":
            continue
        if f == "input.txt" or os.path.getsize(src) < (1 << 20):
            shutil.copy(src, os.path.join(rundir, f))
        else:
            os.symlink(os.path.abspath(src), os.path.join(rundir, f))
    os.symlink(os.path.join(ROOT, "oracle", "_ref", "data"), os.path.join(rundir, "data"))
FLAVOR = os.environ.get("ARTISB200_BENCH_FLAVOR", "fast")
CPU_FLAVOR = os.environ.get("ARTISB200_BENCH_CPU_FLAVOR", "fast")
# several GB of snapshots and run folders: outside the repository tree
CACHE = os.environ.get("ARTISB200_BENCH_CACHE", os.path.join(tempfile.gettempdir(), "artis_b200_bench_cache"))
INTERACTIONS = 26


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------------------
# workload preparation (host driver with the GPU update_packets; not timed)
# ------------------------------------------------------------------------------------------------------------

def prepare_workload(device=0):
    """run the drop-in host binary up to BENCH_TS and snapshot the inputs of update_packets(BENCH_TS)"""
    dump = os.path.join(CACHE, f"{WORKLOAD}_ts{BENCH_TS}")
    static_path = os.path.join(dump, "static.abt")
    before_path = os.path.join(dump, f"ts{BENCH_TS}_before.abt")
    if os.path.exists(static_path) and os.path.exists(before_path):
        return static_path, before_path
    bdir = os.path.join(ROOT, "integration", "_build", WORKLOAD, FLAVOR)
    binary = os.path.join(bdir, "sn3d_b200")
    if not os.path.exists(binary):
        raise RuntimeError(f"{binary} missing (python __graft_entry__.py build in the development container)")
    rundir = os.path.join(CACHE, f"{WORKLOAD}_run")
    make_rundir(WORKLOAD, bdir, rundir)
    # run only timesteps 0..BENCH_TS
    inp = os.path.join(rundir, "input.txt")
    lines = open(inp).read().split("\n")
    lines[2] = f"000 {BENCH_TS + 1:03d}".ljust(24) + " #  2"
    open(inp, "w").write("\n".join(lines))
    os.makedirs(dump, exist_ok=True)
    from artis_b200 import lib as ablib
    env = dict(os.environ, ARTISB200_LIB=ablib.library_path(PRESET), ARTISB200_DUMP_DIR=dump, ARTISB200_DUMP_TS=str(BENCH_TS),
               ARTISB200_DEVICE=str(device), ARTISB200_MODE="gpu")
    t0 = time.time()
    log(f"preparing workload: {binary} (timesteps 0..{BENCH_TS}) ...")
    with open(os.path.join(rundir, "stdout.txt"), "w") as out:
        subprocess.run([binary], cwd=rundir, env=env, stdout=out, stderr=subprocess.STDOUT, check=True)
    log(f"workload prepared in {time.time() - t0:.1f} s")
    for f in os.listdir(dump):  # the *_after snapshot is not needed
        if f.endswith("_after.abt"):
            os.remove(os.path.join(dump, f))
    return static_path, before_path


# ------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------------------

class ClockSampler:
    def __init__(self, device):
        self.device = device
        self.samples = []
        self.stop = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=5)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------
# roofline
# ------------------------------------------------------------------------------------------------------------

def algorithmic_bytes(diag, counters, log2_lines, nions, mean_ncoolingterms_log2, mean_ntrans_log2):
    """minimum compulsory traffic of one update_packets call, from the device work counters"""
    n_step_r, n_line, n_conteval, n_contterm, n_binsearch, n_est, n_ma, n_k, n_step_g = (int(diag[i]) for i in range(9))
    b = 0
    b += n_step_r * (2 * 144 + 24)      # packet hot state read+write, cell scalars
    b += n_line * 40                    # line data (24 B) + two level populations (16 B)
    b += n_contterm * 40                # kept continuum terms
    b += n_binsearch * 8                # binary-search probes (linelist, continuum list, cumulative tables)
    b += n_est * 16                     # f64 estimator read-modify-writes
    b += n_ma * 72                      # 9 macro-atom process rates per transition (search probes counted above)
    b += n_k * 0                        # k-packet selection = two searches, counted in n_binsearch
    b += n_step_g * (2 * 112 + 16)      # gamma step: packet hot state + rho, nnetot, ffegrp
    return b


def bench_config(npackets=None):
    """the workload both arms are quoted on; identical in the two arms' JSON lines"""
    import configs as _configs
    w = WORKLOADS[WORKLOAD]
    if npackets is None:
        npackets = int(_configs.get(WORKLOAD)["opts"]["constexpr int MPKTS"].split("=")[1].strip(" ;"))
    return {"workload": f"{WORKLOAD}: {w['workload']}, timestep {BENCH_TS}",
            "packets_per_gpu": int(npackets), "timestep": BENCH_TS, "model_grid": w["model_grid"],
            "atomic_data": w["atomic_data"],
            "l2_policy": "inputs larger than L2 (packet records + per-cell tables > 126 MB)"}


STAGE_NAMES = ["other", "rpkt_thin", "rpkt_thick", "macroatom", "history_tail"]
STAGE_KERNELS = {"other": "k_wf_stage<other>", "rpkt_thin": "k_wf_stage<rpkt_thin>", "rpkt_thick": "k_wf_stage<rpkt_thick>",
                 "macroatom": "k_wf_stage<macroatom>", "history_tail": "k_propagate"}


def dram_traffic_profile():
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per stage kernel and step, from the committed ncu pass over
    every launch of one full-size step (profiles/r2_dram_traffic.json, written by tools/ncu_dram_traffic.py)"""
    path = os.path.join(ROOT, "profiles", "r2_dram_traffic.json")
    if WORKLOAD != "kilonova_2d":
        return None  # the committed ncu pass is of the default workload
    if os.path.exists(path):
        return json.load(open(path))
    return None


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------

def spectra_binning_leg(eng, n, peak):
    """SURVEY 8f row 2, reported beside the headline (never part of the timed step): the spectra / light-curve binning of the
    packets as the last step left them on the device, all 1 + 100 direction sets in one pass. Algorithmic bytes: the 32-byte
    type sector of every packet + kinematics, energies, escape type and time (136 B) of every escaped one."""
    try:
        eng.set_option("spec_record_dirbin", 1)
        eng.bin_escaped_packets(direction_bins=True, emission_absorption=0, nprocs_exspec=1)
        escaped = int((eng.get_array("spec.dirbin") >= 0).sum())
        eng.set_option("spec_record_dirbin", 0)
        times = []
        for _ in range(5):
            eng.bin_escaped_packets(direction_bins=True, emission_absorption=0, nprocs_exspec=1)
            times.append(eng.last_binning_ms())
        ms = sorted(times)[len(times) // 2]
        b_alg = 32 * n + 136 * escaped
        filled = int((eng.get_array("spec.flux") != 0.).sum())
        return {"kernel": "k_bin_escaped", "packets": int(n), "escaped": escaped, "direction_sets": 101, "ms": ms,
                "algorithmic_bytes": b_alg, "bound": "hbm", "achieved": b_alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": b_alg / (ms * 1e-3) / 1e9 / peak, "flux_bins_filled": filled,
                "reference_does": "1 + 100 passes over the host packets (spectrum_lightcurve.cc:316-337)"}
    except Exception as e:  # an extra: it must never take the bench line down
        return {"error": str(e)}


def grid_update_leg(eng, before):
    """SURVEY 8f row 1, reported beside the headline (never part of the timed step): the LTE grid update of every cell on the
    device copies of the cell state (partition functions, Saha ion balance, electron density). The workload's state was
    balanced by the reference's own update_grid, so the electron densities must come back unchanged: `nne_max_rel_change` is a
    parity property at bench scale (float32 steps are 6e-8)."""
    try:
        import numpy as np
        if "cell.elem_numberdens" not in before:
            return {"skipped": "the workload snapshot has no cell.elem_numberdens (older drop-in binary)"}
        nne_before = eng.get_array("cell.nne", dtype=np.float32)
        times = []
        for _ in range(3):
            eng.update_grid_lte()
            times.append(eng.last_gridupdate_ms())
        nne_after = eng.get_array("cell.nne", dtype=np.float32)
        change = np.abs(nne_after.astype(np.float64) - nne_before) / np.maximum(np.abs(nne_before.astype(np.float64)), 1e-300)
        return {"kernels": "k_lte_perion<partition functions>, k_lte_perion<Saha factors>, k_lte_ion_balance", "cells": int(nne_before.size),
                "ms": sorted(times)[1], "nne_max_rel_change": float(change.max()), "cells_changed_more_than_1e-6": int((change > 1e-6).sum())}
    except Exception as e:  # an extra: it must never take the bench line down
        return {"error": str(e)}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from artis_b200 import distributed as abdist
    from artis_b200 import lib as ablib
    from artis_b200 import snapshot as snap

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the library has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # workload: rank 0 prepares the snapshot, everybody loads it (weak scaling: every rank propagates its own
    # 1e7 packets with rank-distinct Philox keys, as every MPI rank of the reference owns its own MPKTS packets)
    if rank == 0:
        prepare_workload(local_rank)
    if world > 1:
        dist.barrier()
    static_path, before_path = prepare_workload(local_rank)
    static = snap.read_snapshot(static_path)
    before = snap.read_snapshot(before_path)
    n = int(before["packets.count"][0])
    stride = int(before["packets.stride"][0])
    n_sub = int(os.environ.get("ARTISB200_BENCH_NPACKETS", "0"))  # profiling aid only: propagate a packet subset
    if 0 < n_sub < n:
        # evenly spaced over the array: the workload run returns the packets in completion order, so its head holds the
        # packets that had nothing to do in the previous timestep
        pick = (np.arange(n_sub, dtype=np.int64) * n) // n_sub
        before["packets.aos"] = np.ascontiguousarray(before["packets.aos"].reshape(n, stride)[pick]).reshape(-1)
        n = n_sub
    if args.scaling == "strong" and world > 1:
        # strong scaling: the workload's packets are divided among the ranks (the reference divides MPKTS among its MPI ranks
        # the same way when the total is fixed), tables replicated
        per = n // world
        # interleaved, not contiguous: the snapshot holds the packets in completion order of the previous timestep
        before["packets.aos"] = np.ascontiguousarray(before["packets.aos"].reshape(n, stride)[rank::world][:per]).reshape(-1)
        n = per

    eng = ablib.ArtisB200(preset=PRESET, device=local_rank)
    eng.set_option("rng_mode", 0)
    eng.set_option("seed", abdist.rank_seed(20260101, rank))
    eng.set_option("rank", rank)
    eng.set_option("nranks", world)
    eng.set_option("max_steps_per_launch", int(os.environ.get("ARTISB200_MAXSTEPS", "0")))
    # tuning aid: ARTISB200_OPTS="schedule=0,wf_tail=4096,..." (artisb200_set_option names)
    user_opts = {k: str(v) for k, v in WORKLOADS[WORKLOAD].get("options", {}).items()}
    user_opts.update(dict(kv.split("=") for kv in os.environ.get("ARTISB200_OPTS", "").split(",") if "=" in kv))
    for name, value in user_opts.items():
        eng.set_option(name, int(value))
    eng.set_arrays(static)
    eng.commit_static()
    eng.set_arrays(before)
    eng.begin_timestep(BENCH_TS)

    host_packets = torch.from_numpy(before["packets.aos"]).pin_memory()
    work = host_packets.clone().pin_memory()
    eng.upload_packets(host_packets.numpy(), n, stride)
    eng.save_packets_device()

    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local_rank))
    est_ptr, est_count = eng.estimator_device_buffer()

    def reduce_estimators():
        if world > 1:
            # one packed all-reduce replaces the per-array MPI_Allreduce calls of sn3d.cc:565-625 / radfield.cc:988-1030
            abdist.allreduce_estimators_device(eng, local_rank)

    def step_device():
        eng.restore_packets_device()
        eng.begin_timestep(BENCH_TS)
        eng.update_packets(BENCH_TS)
        reduce_estimators()

    def timed(fn, restore=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            ev0.record(stream)
            fn()
            ev1.record(stream)
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1)

    def device_step_timed():
        eng.restore_packets_device()   # not part of update_packets: resets the replayed timestep's input state
        torch.cuda.synchronize()

        def body():
            eng.begin_timestep(BENCH_TS)
            eng.update_packets(BENCH_TS)
            reduce_estimators()
        return timed(body)

    if args.one_step:
        ms = device_step_timed()
        log(f"one step: {ms:.1f} ms (under a profiler this is not a bench value)")
        eng.close()
        return
    log(f"rank {rank}: {n} packets, stride {stride}; warm-up {args.warmup} steps")
    for _ in range(args.warmup):
        device_step_timed()

    step_ms, prop_ms, interactions = [], [], []
    diag_sum = None
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.steps):
            step_ms.append(device_step_timed())
            prop_ms.append(eng.last_timing_ms()[1])
            est = eng.estimators()
            interactions.append(int(est["counters"][INTERACTIONS]))
            diag_sum = est["diag"] if diag_sum is None else diag_sum + est["diag"]
            counters = est["counters"]
    clock_summary = clocks.summary()
    sched_stats = eng.last_schedule_stats()

    # one extra, untimed pass with every stage kernel bracketed by CUDA events: how the step splits over the kernels
    eng.set_option("wf_stage_timing", 1)
    device_step_timed()
    stage_stats = eng.last_schedule_stats()
    stage_total_ms = eng.last_timing_ms()[0]
    stage_diag = eng.get_array("diag_stage").reshape(len(STAGE_NAMES), -1)
    iters = max(1, int(stage_stats["iterations"]))
    ma_per_iter = 1 if int(user_opts.get("wf_refill_masteps", 0)) > 0 else int(user_opts.get("wf_ma_rounds", 3)) | 1
    stage_launches = {"other": iters, "rpkt_thin": iters, "rpkt_thick": iters, "macroatom": iters * ma_per_iter,
                      "history_tail": 1 if stage_stats["tail_ms"] > 0 else 0}
    eng.set_option("wf_stage_timing", 0)

    # end-to-end through the host-buffer call (the drop-in signature): H2D packets, propagate, D2H packets + estimators
    # (as the drop-in binding calls it: finished packets are copied back while the wavefront drains, the array comes
    # back in completion order - the reference's own update_packets permutes it too)
    eng.set_option("stream_download", int(os.environ.get("ARTISB200_BENCH_STREAM_DOWNLOAD", "1")))
    e2e_ms = []
    for it in range(max(1, min(args.steps, 3))):
        work.copy_(host_packets)
        torch.cuda.synchronize()

        def body():
            eng.begin_timestep(BENCH_TS)
            eng.update_packets_host(BENCH_TS, work.numpy(), n, stride)
            reduce_estimators()
            eng.estimators()
        e2e_ms.append(timed(body))

    t_step = sum(step_ms) / len(step_ms)
    t_prop = sum(prop_ms) / len(prop_ms)
    t_e2e = sum(e2e_ms) / len(e2e_ms)
    n_int = sum(interactions) / len(interactions)
    n_gamma = float(diag_sum[9]) / len(step_ms)
    vals = torch.tensor([t_step, t_prop, t_e2e, n_int, n_gamma], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = vals.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = vals.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t_step, t_prop, t_e2e = (float(x) for x in tmax[:3])
        n_int_total = float(tsum[3])
        n_gamma_total = float(tsum[4])
    else:
        n_int_total = n_int
        n_gamma_total = n_gamma

    if rank == 0:
        nlines = static["line.nu"].size
        diag_mean = diag_sum / len(step_ms)
        b_alg = algorithmic_bytes(diag_mean, counters, np.log2(max(nlines, 2)), static["ion.nlevels"].size, 0, 0)
        peak, peak_kind = measured_peak_gbs()
        achieved = b_alg / (t_prop * 1e-3) / 1e9
        launches = int(diag_mean[10])
        ncells = static["cell.ffegrp"].size
        # per kernel family: algorithmic bytes of the family's own work counters / the family's CUDA-event time of the
        # stage-timing pass (kernels run one after the other there), per launch; DRAM traffic from the committed ncu pass
        traffic = dram_traffic_profile()
        stage_rooflines = []
        for si, sname in enumerate(STAGE_NAMES):
            ms = stage_stats["stage_ms"].get(sname, 0.0) if sname != "history_tail" else stage_stats["tail_ms"]
            b_stage = algorithmic_bytes(stage_diag[si], counters, 0, 0, 0, 0)
            nlaunch = stage_launches.get(sname, 0)
            if ms <= 0 or nlaunch <= 0:
                continue
            tr = traffic["kernels"].get(sname) if traffic else None
            stage_rooflines.append({
                "kernel": STAGE_KERNELS[sname], "bound": "hbm", "launches_per_step": nlaunch, "ms_per_step": ms,
                "algorithmic_bytes_per_launch": b_stage / nlaunch, "avg_launch_ms": ms / nlaunch,
                "achieved": b_stage / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": b_stage / (ms * 1e-3) / 1e9 / peak,
                "traffic": (tr["dram_bytes_per_step"] / tr["launches_per_step"]) if tr else None,
                "traffic_over_algorithmic": (tr["dram_bytes_per_step"] / b_stage) if (tr and b_stage > 0) else None})
        dominant = max(stage_rooflines, key=lambda r: r["ms_per_step"]) if stage_rooflines else None
        roofline = dict(dominant) if dominant else {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                                                      "frac": achieved / peak, "traffic": None}
        roofline.update({"peak_source": peak_kind, "algorithmic_bytes": roofline.get("algorithmic_bytes_per_launch"),
                         "traffic_source": (traffic.get("source") if traffic else None),
                         "note": "dominant kernel of the step (largest share of the stage-timing pass); per launch; the whole step is in roofline_step"})
        out = {
            "metric": "packet-interactions/sec per timestep", "value": n_int_total / (t_step * 1e-3), "unit": "interactions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(n),
            "workload_details": {"model_cells": int(ncells), "lines": int(nlines), "levels": int(static["level.epsilon"].size),
                                 "bf_continua": int(static["cont.nu_edge"].size), "rng": "philox4x32-10",
                                 "interactions_per_step_per_gpu": n_int},
            "e2e": {"value": n_int_total / (t_e2e * 1e-3), "unit": "interactions/s", "h2d_bytes_per_step": int(n * stride),
                    "d2h_bytes_per_step": int(n * stride + est_count * 8), "ms_per_step": t_e2e},
            # per step: the propagation launches (stage kernels, sort, resets) + the 5 per-cell table-build kernels
            "gpu_launches": int(launches + 5),
            "roofline": roofline,
            "roofline_step": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                              "kernel": "all propagation kernels of the step", "kernel_ms": t_prop, "algorithmic_bytes": int(b_alg),
                              "traffic": (sum(v["dram_bytes_per_step"] for v in traffic["kernels"].values()) if traffic else None)},
            "roofline_stages": stage_rooflines,
            "schedule": {"options": user_opts, "iterations": sched_stats["iterations"], "launches": sched_stats["launches"],
                         "tail_packets": sched_stats["tail_packets"], "tail_ms": sched_stats["tail_ms"],
                         "stage_timing_pass": {"total_ms": stage_total_ms, "stage_ms": stage_stats["stage_ms"],
                                               "tail_ms": stage_stats["tail_ms"]}},
            "event_counters": {k: int(counters[i]) for k, i in (("interactions", 26), ("electron_scatterings", 27), ("ma_activation_bb", 4),
                                                                 ("ma_activation_bf", 5), ("ma_deactivation_bb", 9), ("k_from_ff", 19),
                                                                 ("k_from_bf", 20), ("cellcrossings", 29), ("escapes", 33),
                                                                 ("resonance_scatterings", 28))},
            "clocks": clock_summary,
            "work_counters": {k: int(diag_mean[i]) for i, k in enumerate(
                ["rpkt_steps", "lines_visited", "cont_evals", "cont_terms", "binsearch_probes", "estimator_adds", "ma_steps",
                 "k_steps", "gamma_steps", "gamma_events", "launches", "packet_segments", "table_passes"])},
        }
        # gamma transport does not touch the reference's INTERACTIONS counter, so the gamma-only configuration
        # is quoted in physical gamma events (Compton / photoelectric / pair, gammapkt.cc:720-747) per second as well
        out["gamma_events_per_s"] = n_gamma_total / (t_step * 1e-3)
        out["spectra_binning"] = spectra_binning_leg(eng, n, peak)
        out["grid_update_lte"] = grid_update_leg(eng, before)
        out["table_windows"] = {"passes_per_step": int(diag_mean[12]), "cells": int(ncells),
                                "note": "1 = the per-cell tables of every cell are resident; > 1 = cell-batched tables"}
        if not args.no_cpu_baseline and world == 1:
            try:
                # same model, same timestep, independent random numbers: the two codes must agree within Monte Carlo noise
                out["cpu_baseline"], out["crosscheck"] = cpu_baseline(n_int / n)
            except Exception as e:  # the baseline must never take the bench line down
                out["cpu_baseline"] = {"value": None, "unit": "interactions/s", "cores": 1, "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(out), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the compiled reference's own update_packets on this box
# ------------------------------------------------------------------------------------------------------------

def _set_input_line(path, index, text):
    lines = open(path).read().split("\n")
    lines[index] = text.ljust(24) + f" # {index:2d}"
    open(path, "w").write("\n".join(lines))


def _run_reference_processes(rundirs, binary):
    """one single-rank reference process per run folder, all at once (the reference's production parallelism is one MPI
    rank per core with no communication inside update_packets, update_packets.cc:561/631); returns the ARTISB200_TIMING
    record of timestep BENCH_TS of every process"""
    import run_oracle
    procs = []
    for rundir in rundirs:
        env = dict(os.environ, ARTISB200_MODE="ref")
        env.pop("ARTISB200_DUMP_DIR", None)
        out = open(os.path.join(rundir, "stdout.txt"), "w")
        procs.append((rundir, subprocess.Popen([binary], cwd=rundir, env=env, stdout=out, stderr=subprocess.STDOUT)))
    res = []
    for rundir, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"reference process failed in {rundir}")
        res.append([t for t in run_oracle.timing_lines(rundir) if t["nts"] == BENCH_TS][-1])
    return res


def cpu_reference_setup(nproc, rundir_root):
    """fresh run folders of the bounded CPU sample (same model and atomic data, MPKTS = 1e5 packets per process, distinct
    seeds)"""
    import run_oracle
    odir = run_oracle.oracle_dir(CPU_SAMPLE_CONFIG, CPU_FLAVOR)
    binary = os.path.join(odir, "sn3d_ref")
    if not os.path.exists(binary):
        raise RuntimeError(f"{binary} missing")
    rundirs = []
    for r in range(nproc):
        rundir = os.path.join(rundir_root, f"cpu_rank{r}")
        make_rundir(CPU_SAMPLE_CONFIG, odir, rundir)
        _set_input_line(os.path.join(rundir, "input.txt"), 0, f"{20260101 + 1000 * r}")
        rundirs.append(rundir)
    return rundirs, binary


def cpu_reference_evolve(rundirs, binary):
    """timesteps 0..BENCH_TS from scratch. Leaves gridsave_ts<BENCH_TS>.tmp / packets_0000_ts<BENCH_TS>.tmp and an
    input.txt that resumes there (grid.cc:974, packet.cc:253-311, sn3d.cc:663-716), so that every further call of
    cpu_reference_restart() times update_packets(BENCH_TS) alone"""
    for rundir in rundirs:
        _set_input_line(os.path.join(rundir, "input.txt"), 2, f"000 {BENCH_TS + 1:03d}")
    res = _run_reference_processes(rundirs, binary)
    for rundir in rundirs:
        for f in (f"gridsave_ts{BENCH_TS}.tmp", f"packets_0000_ts{BENCH_TS}.tmp"):
            if not os.path.exists(os.path.join(rundir, f)):
                raise RuntimeError(f"{f} was not written in {rundir}")
        _set_input_line(os.path.join(rundir, "input.txt"), 2, f"{BENCH_TS:03d} {BENCH_TS + 1:03d}")
        _set_input_line(os.path.join(rundir, "input.txt"), 16, "1")
    return res


def cpu_reference_restart(rundirs, binary):
    """resume at BENCH_TS from the saved grid and packets: one update_packets(BENCH_TS) per process"""
    return _run_reference_processes(rundirs, binary)


def cpu_restart_from_gpu_state(nproc, rundir_root, npackets_each):
    """run folders that resume at BENCH_TS from the state the GPU workload run saved: its cell state
    (gridsave_ts<BENCH_TS>.tmp, written by the reference's own driver in the drop-in binary) and disjoint slices of its
    packets. The CPU reference then propagates a sample of exactly the packets, in exactly the cell state, of the GPU step."""
    gpu_run = os.path.join(CACHE, f"{WORKLOAD}_run")
    gridsave = os.path.join(gpu_run, f"gridsave_ts{BENCH_TS}.tmp")
    pktfile = os.path.join(gpu_run, f"packets_0000_ts{BENCH_TS}.tmp")
    if not (os.path.exists(gridsave) and os.path.exists(pktfile)):
        return None
    import struct
    import numpy as np
    rundirs, binary = cpu_reference_setup(nproc, rundir_root)
    with open(pktfile, "rb") as f:
        total = struct.unpack("<q", f.read(8))[0]
    stride = (os.path.getsize(pktfile) - 8) // total
    raw = np.memmap(pktfile, dtype=np.uint8, mode="r", offset=8, shape=(total, stride))
    # an UNBIASED sample: the drop-in run returns the packets in completion order (streamed download), so the head of the
    # file holds the packets that finished the previous timestep first (e.g. pellets that did not decay); every process
    # takes every step-th packet, interleaved with the other processes
    each = max(1, min(npackets_each, total // nproc))
    for r, rundir in enumerate(rundirs):
        # evenly spaced over the WHOLE file (its tail holds the packets that took longest in the previous timestep)
        idx = (((np.arange(each, dtype=np.int64) * nproc) + r) * total) // (each * nproc)
        chunk = np.ascontiguousarray(raw[idx])
        with open(os.path.join(rundir, f"packets_0000_ts{BENCH_TS}.tmp"), "wb") as g:
            g.write(struct.pack("<q", each))
            g.write(chunk.tobytes())
        shutil.copy(gridsave, os.path.join(rundir, f"gridsave_ts{BENCH_TS}.tmp"))
        _set_input_line(os.path.join(rundir, "input.txt"), 2, f"{BENCH_TS:03d} {BENCH_TS + 1:03d}")
        _set_input_line(os.path.join(rundir, "input.txt"), 16, "1")
    return rundirs, binary, each


def cpu_baseline(gpu_interactions_per_packet=None):
    """the compiled reference on the host cores of this box, K = min(8, cores) single-rank processes side by side: a bounded
    sample (1e5 packets each) of the GPU step's own packets in the GPU step's own cell state when the workload run left
    its restart files, else of an independent CPU evolution of the same model. Also the cross-check of the two codes:
    interactions per packet of the GPU step against the K CPU samples (mean, standard error, z)."""
    root = os.path.join(CACHE, "cpu_baseline")
    os.makedirs(root, exist_ok=True)
    import configs as _configs
    k = max(1, min(8, os.cpu_count() or 1))
    npk = int(_configs.get(CPU_SAMPLE_CONFIG)["opts"]["constexpr int MPKTS"].split("=")[1].strip(" ;"))
    prepared = cpu_restart_from_gpu_state(k, root, npk)
    if prepared is not None:
        rundirs, binary, npk = prepared
        res = cpu_reference_restart(rundirs, binary)
        how = (f"{k} processes x {npk} packets: disjoint interleaved samples of the GPU step's own packets, resumed by the reference "
               f"from the cell state of the GPU workload run (gridsave_ts{BENCH_TS}.tmp)")
    else:
        rundirs, binary = cpu_reference_setup(k, root)
        res = cpu_reference_evolve(rundirs, binary)
        how = f"{k} processes x {npk} packets of an independent CPU evolution of timesteps 0..{BENCH_TS} of the same model"
    total_int = sum(r["interactions"] for r in res)
    wall = max(r["wall_s"] for r in res)
    per_packet = [r["interactions"] / r["npackets"] for r in res]
    mean = sum(per_packet) / len(per_packet)
    sd = (sum((x - mean) ** 2 for x in per_packet) / max(1, len(per_packet) - 1)) ** 0.5
    out = {"value": total_int / wall, "unit": "interactions/s", "cores": k, "kind": "reference",
           "sample": f"{CPU_SAMPLE_CONFIG}: {how}; update_packets({BENCH_TS}) wall {wall:.2f} s (slowest process), {total_int} interactions",
           "wall_s": wall}
    cross = {"interactions_per_packet_cpu_reference": mean, "cpu_seeds": len(per_packet), "cpu_std_between_seeds": sd,
             "cpu_standard_error": sd / len(per_packet) ** 0.5 if len(per_packet) > 1 else None,
             "same_cell_state_and_packets": prepared is not None}
    if gpu_interactions_per_packet is not None:
        cross["interactions_per_packet_gpu"] = gpu_interactions_per_packet
        if len(per_packet) > 1 and sd > 0:
            z = (gpu_interactions_per_packet - mean) / (sd / len(per_packet) ** 0.5)
            cross.update(z=z, criterion="|z| < 4 (GPU mean over all its packets against the mean of the CPU seeds, standard error from "
                                        "the scatter between the seeds)", passed=bool(abs(z) < 4.0))
    return out, cross


def run_reference(args):
    """the reference's own CPU update_packets on every host core: evolve timesteps 0..BENCH_TS-1 once (not timed), then
    `warmup` + `steps` passes, each resuming from the saved grid and packets and running update_packets(BENCH_TS) alone"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_begin = time.time()
    budget_s = float(os.environ.get("ARTISB200_REFERENCE_BUDGET_S", "270"))
    cores = os.cpu_count() or 1
    root = os.path.join(CACHE, "reference_arm")
    os.makedirs(root, exist_ok=True)
    rundirs, binary = cpu_reference_setup(cores, root)
    log(f"reference arm: evolving timesteps 0..{BENCH_TS} on {cores} cores (not timed) ...")
    first = cpu_reference_evolve(rundirs, binary)
    log(f"evolved in {time.time() - t_begin:.1f} s; update_packets({BENCH_TS}) took {max(r['wall_s'] for r in first):.2f} s")
    values, walls = [], []
    for it in range(args.warmup + args.steps):
        t0 = time.time()
        res = cpu_reference_restart(rundirs, binary)
        values.append(sum(r["interactions"] for r in res) / max(r["wall_s"] for r in res))
        walls.append(max(r["wall_s"] for r in res))
        remaining_needed = (time.time() - t0) * (args.warmup + args.steps - it - 1)
        if it + 1 >= args.steps and (time.time() - t_begin) + remaining_needed > budget_s:
            log(f"reference arm: stopping after {it + 1} passes to stay within {budget_s:.0f} s")
            break
    n_timed = min(args.steps, len(values))
    use, use_walls = values[-n_timed:], walls[-n_timed:]
    v = sum(use) / len(use)
    npk = res[0]["npackets"]
    sample = (f"{CPU_SAMPLE_CONFIG}: {cores} concurrent single-rank reference processes x {npk} packets of the same model and "
              f"atomic data (a bounded sample of the workload: {cores * npk} of the {bench_config()['packets_per_gpu']} packets), "
              f"update_packets({BENCH_TS}) resumed from saved grid and packets each pass")
    out = {"impl": "reference", "metric": "packet-interactions/sec per timestep", "value": v, "unit": "interactions/s",
           "n_gpus": args.gpus, "steps": n_timed, "warmup": len(values) - n_timed, "ms_per_step": 1e3 * sum(use_walls) / len(use_walls),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": bench_config(),
           "cpu_baseline": {"value": v, "unit": "interactions/s", "cores": cores, "kind": "reference", "sample": sample},
           "e2e": {"value": v, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "elapsed_s": time.time() - t_begin}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--one-step", action="store_true", help="profiling aid: run exactly one device step and exit (no JSON line)")
    ap.add_argument("--workload", default=os.environ.get("ARTISB200_BENCH_CONFIG", "kilonova_2d"), choices=sorted(WORKLOADS),
                    help="BASELINE.json configuration (default: configs[1], the one the metric is quoted on)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU propagates the workload's packet count; strong: the packets are divided among the GPUs")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
