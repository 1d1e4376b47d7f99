"""Flat little-endian fixture files shared by MATLAB, Python and C.

Layout (all little endian):
    bytes 0-3   magic "CFSB"          bytes 4-7   uint32 version (1)         bytes 8-11  uint32 number of arrays
    per array:  32 bytes name (NUL padded ASCII), uint32 ndim, ndim x uint64 dims, then prod(dims) float64 values in
                COLUMN-MAJOR order (as MATLAB stores them: fread(fid, prod(dims), 'double') + reshape gives the array)
matlab/read_cfs_fixture.m reads the same file into a struct.  Python arrays are written from / returned in their numpy
shape; a (B, n) numpy array of per-problem rows therefore appears in MATLAB as B x n (transpose for the n x B column
layout the C ABI takes)."""
import struct

import numpy as np

MAGIC = b"CFSB"


def write_fixture(path, arrays):
    """arrays: dict name -> array-like (converted to float64)"""
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<II", 1, len(arrays)))
        for name, a in arrays.items():
            a = np.asarray(a, dtype=np.float64)
            nm = name.encode("ascii")
            if len(nm) > 31:
                raise ValueError("fixture array name too long: %s" % name)
            f.write(nm.ljust(32, b"\0") + struct.pack("<I", a.ndim) + struct.pack("<%dQ" % a.ndim, *a.shape))
            f.write(np.asfortranarray(a).tobytes(order="F"))


def read_fixture(path):
    out = {}
    with open(path, "rb") as f:
        if f.read(4) != MAGIC:
            raise ValueError("not a CFSB fixture: %s" % path)
        version, count = struct.unpack("<II", f.read(8))
        if version != 1:
            raise ValueError("unsupported fixture version %d" % version)
        for _ in range(count):
            name = f.read(32).rstrip(b"\0").decode("ascii")
            (ndim,) = struct.unpack("<I", f.read(4))
            dims = struct.unpack("<%dQ" % ndim, f.read(8 * ndim)) if ndim else ()
            n = int(np.prod(dims)) if ndim else 1
            data = np.frombuffer(f.read(8 * n), dtype="<f8")
            out[name] = data.reshape(dims, order="F").copy() if ndim else data.reshape(())
    return out


def write_batch_fixture(path, cfg):
    """the seeded headline batch (synthetic.batch_config_m16ib) as one fixture: everything cfs_solve_batch needs"""
    s = cfg["sys_info"]
    write_fixture(path, dict(H=s["H"], njoint=s["njoint"], QQ=s["QQ"], lim=s["lim"], MAX_input=s["MAX_input"],
                             epsilon_O=s["epsilon_O"], MAX_O_ITER=s["MAX_O_ITER"], obs_l=cfg["obs"][0]["l"], obs_D=cfg["obs"][0]["D"],
                             obs_epsilon=cfg["obs"][0]["epsilon"], x0=cfg["x0"], ff=cfg["ff"], caug=cfg["caug"], xref=cfg["xref"],
                             theta0=cfg["theta0"], thetag=cfg["thetag"]))
