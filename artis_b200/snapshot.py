"""Reader/writer of the ABT1 snapshot container (integration/b200_snapshot.h): a flat sequence of named,
typed arrays that carries the update_packets() boundary data between the host driver, tests and bench.py."""
import struct

import numpy as np

MAGIC = b"ARTISB2\n"
_DTYPES = {"d": np.float64, "f": np.float32, "i": np.int32, "q": np.int64, "B": np.uint8, "Q": np.uint64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


def read_snapshot(path):
    """-> dict name -> numpy array (scalars are arrays of length 1)"""
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != MAGIC:
        raise ValueError(f"{path}: not an ABT1 snapshot")
    pos = 8
    n = len(data)
    while pos < n:
        start = pos
        (name_len,) = struct.unpack_from("<I", data, pos)
        pos += 4
        name = data[pos:pos + name_len].decode()
        pos += name_len
        code = chr(data[pos])
        pos += 1
        (count,) = struct.unpack_from("<Q", data, pos)
        pos += 8
        dt = np.dtype(_DTYPES[code])
        nbytes = count * dt.itemsize
        out[name] = np.frombuffer(data, dtype=dt, count=count, offset=pos).copy()
        pos += nbytes
        pos += (8 - ((pos - start) % 8)) % 8
    return out


def write_snapshot(path, arrays):
    with open(path, "wb") as f:
        f.write(MAGIC)
        for name, arr in arrays.items():
            arr = np.ascontiguousarray(arr)
            code = _CODES[arr.dtype]
            nb = name.encode()
            rec = struct.pack("<I", len(nb)) + nb + code.encode() + struct.pack("<Q", arr.size) + arr.tobytes()
            rec += b"\0" * ((8 - (len(rec) % 8)) % 8)
            f.write(rec)


def dtype_code(arr):
    return _CODES[np.dtype(arr.dtype)]


# field layout of the reference's AoS Packet (packet.h:109-156), 240-byte CPU layout
PACKET_FIELDS = [
    ("prop_time", "<f8", 0), ("pos", "<f8", 8, 3), ("dir", "<f8", 32, 3), ("nu_cmf", "<f8", 56), ("e_cmf", "<f8", 64),
    ("nu_rf", "<f8", 72), ("e_rf", "<f8", 80), ("next_trans", "<i4", 88), ("nscatterings", "<i4", 92),
    ("emissiontype", "<i4", 96), ("em_pos", "<f8", 104, 3), ("em_time", "<f4", 128), ("absorptiontype", "<i4", 132),
    ("absorptionfreq", "<f8", 136), ("stokes_q", "<f8", 144), ("stokes_u", "<f8", 152), ("trueemissiontype", "<i4", 160),
    ("trueem_pos", "<f8", 168, 3), ("trueem_time", "<f4", 192), ("type", "<i4", 196), ("cellindex", "<i4", 200),
    ("escape_type", "<i4", 204), ("escape_time", "<f4", 208), ("tdecay", "<f8", 216), ("number", "<i4", 224),
    ("originated_from_particlenotgamma", "u1", 228), ("pellet_decaytype", "<i4", 232), ("pellet_nucindex", "<i4", 236),
]


def packet_dtype(stride):
    """numpy structured dtype over the raw AoS bytes (stride 240, or 256 with the 16-byte rngstate prefix)"""
    base = stride - 240
    names, formats, offsets = [], [], []
    if base == 16:
        names.append("rngstate")
        formats.append(("<u4", (4,)))
        offsets.append(0)
    for fld in PACKET_FIELDS:
        names.append(fld[0])
        formats.append((fld[1], (fld[3],)) if len(fld) == 4 else fld[1])
        offsets.append(base + fld[2])
    return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": stride})


def packets_view(snap):
    stride = int(snap["packets.stride"][0])
    return snap["packets.aos"].view(packet_dtype(stride))
