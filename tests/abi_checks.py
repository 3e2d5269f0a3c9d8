"""Error behaviour of the C ABI (include/artis_b200.h): every misuse returns nonzero with a message naming the problem,
nothing is silently accepted. The reference's convention is assert_always -> log -> abort (mpi_logging.h:123-130); the
binding aborts on any nonzero return, so what matters here is that misuse IS reported. Shared by the host-simulation
suite (CPU) and the GPU suite."""
import numpy as np
import pytest

from artis_b200 import lib as ablib
from tests import fixtures


def check_abi_errors(libpath):
    fx = fixtures.load_golden("classic_toy_1d", 3)
    eng = ablib.ArtisB200(libpath=libpath)
    err = ablib.ArtisB200Error
    with pytest.raises(err, match="required table"):
        eng.commit_static()
    with pytest.raises(err, match="unknown array"):
        eng.set_array("no.such.array", np.zeros(3))
    with pytest.raises(err, match="dtype"):
        eng.set_array("line.nu", np.zeros(3, dtype=np.float32))  # line.nu is f64
    with pytest.raises(err, match="cannot be set"):
        eng.set_array("est.J", np.zeros(3))  # outputs belong to the library
    with pytest.raises(err, match="unknown option"):
        eng.set_option("no_such_option", 1)
    with pytest.raises(err, match="schedule"):
        eng.set_option("schedule", 7)
    with pytest.raises(err, match="commit_static"):
        eng.begin_timestep(0)

    eng.set_arrays(fx["static"])
    eng.commit_static()
    with pytest.raises(err, match="per-timestep array"):
        eng.begin_timestep(fx["nts"])  # cell state missing
    eng.set_arrays(fx["before"])
    with pytest.raises(err, match="out of range"):
        eng.begin_timestep(10_000)
    with pytest.raises(err, match="begin_timestep"):
        eng.update_packets(fx["nts"])  # before begin_timestep
    eng.begin_timestep(fx["nts"])
    with pytest.raises(err, match="no packets"):
        eng.update_packets(fx["nts"])
    n = int(fx["before"]["packets.count"][0])
    stride = int(fx["before"]["packets.stride"][0])
    aos = fx["before"]["packets.aos"].copy()
    with pytest.raises(err, match="stride"):
        eng.upload_packets(aos, n, 200)
    eng.set_option("rng_mode", 1)
    with pytest.raises(err, match="xoshiro"):
        eng.upload_packets(aos[: n * 240].copy(), n, 240)  # the reference RNG state needs the 256-byte GPU_ON Packet
    eng.upload_packets(aos, n, stride)
    with pytest.raises(err, match="differ"):
        eng.download_packets(aos, n - 1, stride)
    with pytest.raises(err, match="begin_timestep"):
        eng.update_packets(fx["nts"] + 1)  # a different timestep than the one begun
    # wrong cell-state length is caught when the timestep begins
    eng.set_array("cell.rho", np.zeros(3, dtype=np.float32))
    with pytest.raises(err, match="lengths"):
        eng.begin_timestep(fx["nts"])
    eng.set_array("cell.rho", fx["before"]["cell.rho"])
    eng.begin_timestep(fx["nts"])
    eng.update_packets(fx["nts"])  # and the context still works after all of that
    assert int(eng.get_array("counters")[fixtures.INTERACTIONS]) == int(fx["after"]["counters"][fixtures.INTERACTIONS])
    with pytest.raises(err, match="mismatch"):
        eng.lib.artisb200_get_array  # noqa: B018  (attribute exists)
        out = np.zeros(5)
        eng._check(eng.lib.artisb200_get_array(eng.ctx, b"est.J", b"d", out.ctypes.data, 5), "get_array")  # wrong count
    eng.close()


def check_device_error_record(libpath):
    """the device-side stand-in for the reference's assert_always: a packet of a type no code path knows
    (update_packets.cc:312) is reported by update_packets with the failed assertion and the packet index, instead of
    being coerced silently"""
    fx = fixtures.load_golden("kilonova_toy", 4)
    before = fx["before"]
    n = int(before["packets.count"][0])
    stride = int(before["packets.stride"][0])
    for options in ({"schedule": 0}, {"schedule": 1, "wf_tail": 0}):
        eng = fixtures.make_engine(libpath, fx, rng="philox", options=options)
        aos = before["packets.aos"].copy()
        pk = aos.view(fixtures.snap.packet_dtype(stride))
        victim = int(np.nonzero(pk["type"] == 11)[0][3])
        pk["type"][victim] = 77
        with pytest.raises(ablib.ArtisB200Error, match="unknown packet type") as info:
            eng.update_packets_host(fx["nts"], aos, n, stride)
        assert f"packet index {victim}" in str(info.value)
        record = eng.get_array("dev_error")
        assert record[0] == 6 and record[1] == victim and record[2] == 77 and record[3] == 1
        # the record is cleared when the next timestep begins, and a clean run passes again
        eng.set_arrays(before)
        eng.begin_timestep(fx["nts"])
        good = before["packets.aos"].copy()
        eng.update_packets_host(fx["nts"], good, n, stride)
        assert eng.get_array("dev_error")[0] == 0
        eng.close()


def check_options_summary(libpath, preset):
    eng = ablib.ArtisB200(libpath=libpath)
    summary = eng.options_summary()
    assert f"preset={preset}" in summary
    assert eng.lib.artisb200_options_hash() != 0
    eng.close()
