"""Packet files of the reference written by the library (the I/O part; csrc/packetio.h).

The text file is compared BYTE FOR BYTE with the packets00_0000.out that the compiled reference itself wrote at the end of a
complete run (development container: needs oracle/_ref), from the same final packets; the binary restart file with the
reference's packets_0000_ts<N>.tmp. Without the oracle build the format is checked on a committed fixture's packets."""
import os

import numpy as np
import pytest

from artis_b200 import snapshot as snap
from tests import fixtures


def _engine(preset):
    return fixtures.ablib.ArtisB200(libpath=fixtures.hostsim_library(preset))


@pytest.mark.parametrize("config", ["classic3d_toy", "kilonova_toy", "classic_toy_1d"])
def test_text_and_restart_files_equal_the_reference_s(config, tmp_path):
    import run_oracle
    binary = os.path.join(run_oracle.oracle_dir(config, "parity"), "sn3d_ref")
    if not os.path.exists(binary):
        pytest.skip("oracle/_ref not built (development container only)")
    rundir = run_oracle.run(config, "parity", "ref_perpacket", "all", rundir=str(tmp_path / "run"))
    last = int(open(os.path.join(rundir, "input.txt")).read().split("\n")[2].split()[1]) - 1
    after = snap.read_snapshot(os.path.join(rundir, "dump", f"ts{last}_after.abt"))
    n, stride = int(after["packets.count"][0]), int(after["packets.stride"][0])
    eng = _engine(fixtures.PRESET_OF[config])
    mine = tmp_path / "packets00_0000.out"
    eng.write_text_packets(mine, after["packets.aos"], n, stride, keep_escaped_gammas=True)
    theirs = open(os.path.join(rundir, "packets00_0000.out"), "rb").read()
    assert theirs.count(b"\n") == n + 1
    assert open(mine, "rb").read() == theirs, "packets00_0000.out differs from the reference's"
    # binary restart file of the last completed timestep
    restart = [f for f in os.listdir(rundir) if f.startswith("packets_0000_ts") and f.endswith(".tmp")]
    assert restart, "the reference wrote no restart file"
    their_bytes = open(os.path.join(rundir, sorted(restart)[-1]), "rb").read()
    raw, count = eng.read_temp_packetsfile(os.path.join(rundir, sorted(restart)[-1]), stride)
    assert count == n and raw.tobytes() == their_bytes[8:]
    eng.write_temp_packetsfile(tmp_path / "packets_0000_ts0.tmp", raw, count, stride)
    assert open(tmp_path / "packets_0000_ts0.tmp", "rb").read() == their_bytes
    eng.close()


def test_text_file_format_on_a_committed_fixture(tmp_path):
    fx = fixtures.load_golden("classic3d_toy", 2)
    after = fx["after"]
    n, stride = int(after["packets.count"][0]), int(after["packets.stride"][0])
    pk = snap.packets_view(after)
    eng = _engine("classic")  # POL_ON: two more columns
    path = tmp_path / "packets.out"
    eng.write_text_packets(path, after["packets.aos"], n, stride, keep_escaped_gammas=False)
    lines = open(path).read().split("\n")
    assert lines[0].startswith("#number where type_id posx") and lines[0].endswith("pellet_nucindex pellet_decaytype")
    ncols = len(lines[0].split())
    assert ("stokes_q" in lines[0]) and ncols == 34
    escaped_gammas = int(((pk["type"] == 32) & (pk["escape_type"] == 10)).sum())
    body = [ln for ln in lines[1:] if ln]
    assert escaped_gammas > 0 and len(body) == n - escaped_gammas
    kept = pk[~((pk["type"] == 32) & (pk["escape_type"] == 10))]
    for row, p in list(zip(body, kept))[::37]:
        tok = row.split()
        assert len(tok) == ncols and int(tok[0]) == p["number"] and int(tok[2]) == p["type"]
        assert tok[3] == "%g" % p["pos"][0] and tok[13] == "%g" % p["nu_rf"] and tok[15] == "%g" % p["escape_time"]
        assert int(tok[16]) == p["emissiontype"] and tok[25] == "%g" % p["stokes_q"]
    with pytest.raises(fixtures.ablib.ArtisB200Error, match="cannot open"):
        eng.write_text_packets(tmp_path / "no_such_dir" / "p.out", after["packets.aos"], n, stride)
    eng.close()
