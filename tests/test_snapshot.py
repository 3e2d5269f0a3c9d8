import os

import numpy as np

from artis_b200 import snapshot as snap


def test_snapshot_roundtrip(tmp_path):
    arrays = {"scalar.tmin": np.array([1.5]), "line.nu": np.linspace(5e15, 1e14, 17), "ion.nlevels": np.arange(7, dtype=np.int32),
              "trans.forbidden": np.array([0, 1, 1], dtype=np.uint8), "built.cont_keepbits": np.array([2**63 + 5], dtype=np.uint64),
              "cell.rho": np.arange(5, dtype=np.float32), "counters": np.arange(34, dtype=np.int64), "empty": np.zeros(0)}
    path = os.path.join(tmp_path, "x.abt")
    snap.write_snapshot(path, arrays)
    back = snap.read_snapshot(path)
    assert list(back) == list(arrays)
    for k, v in arrays.items():
        assert back[k].dtype == v.dtype and np.array_equal(back[k], v)


def test_packet_dtype_matches_reference_layout():
    # sizeof(Packet) == 240 (CPU) / 256 (GPU_ON); offsets measured with the reference headers
    d240, d256 = snap.packet_dtype(240), snap.packet_dtype(256)
    assert d240.itemsize == 240 and d256.itemsize == 256
    assert d240.fields["type"][1] == 196 and d240.fields["cellindex"][1] == 200 and d240.fields["pellet_nucindex"][1] == 236
    assert d256.fields["prop_time"][1] == 16 and d256.fields["rngstate"][1] == 0
