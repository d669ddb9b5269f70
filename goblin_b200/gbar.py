"""GBAR: the flat "named arrays" container used by the tools in this repo.

Each record is::

    "GBAR" | u32 name_len | name | u32 dtype | u32 ndim | u64 dims[ndim] | data

dtype: 0 = float32, 1 = uint32, 2 = int32, 3 = uint8.  Little endian.
"""
import struct

import numpy as np

_DTYPES = {0: np.float32, 1: np.uint32, 2: np.int32, 3: np.uint8}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


def load(path):
    """Read a GBAR file into an ordered dict name -> numpy array."""
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    off = 0
    while off < len(buf):
        if buf[off:off + 4] != b"GBAR":
            raise ValueError(f"{path}: bad record magic at byte {off}")
        (nl,) = struct.unpack_from("<I", buf, off + 4)
        off += 8
        name = buf[off:off + nl].decode()
        off += nl
        dt, nd = struct.unpack_from("<II", buf, off)
        off += 8
        dims = struct.unpack_from("<%dQ" % nd, buf, off)
        off += 8 * nd
        dtype = np.dtype(_DTYPES[dt])
        n = int(np.prod(dims)) if nd else 1
        arr = np.frombuffer(buf, dtype=dtype, count=n, offset=off).reshape(dims)
        off += n * dtype.itemsize
        out[name] = arr
    return out


def save(path, arrays):
    """Write a dict name -> numpy array as a GBAR file."""
    with open(path, "wb") as f:
        for name, arr in arrays.items():
            arr = np.ascontiguousarray(arr)
            code = _CODES[arr.dtype]
            nb = name.encode()
            f.write(b"GBAR" + struct.pack("<I", len(nb)) + nb)
            f.write(struct.pack("<II", code, arr.ndim))
            f.write(struct.pack("<%dQ" % arr.ndim, *arr.shape))
            f.write(arr.tobytes())
