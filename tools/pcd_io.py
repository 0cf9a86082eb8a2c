"""PCD v0.7 reader (ascii / binary / binary_compressed) — TEST INFRASTRUCTURE.

The reference ships its real data as PCL `.pcd` files written with `DATA binary_compressed`
(`Calibration_Tookit/Multi_LiCa/data/demo/lidar_{1,2,3}.pcd`, read by `Lidar.read_pcd`, `Lidar.py:73-95`, through Open3D;
`SensorsCalibration/lidar2lidar/auto_calib/data/000{1,2,3}/*.pcd`, read by `pcl::io::loadPCDFile`, `run_lidar2lidar.cpp`).
Neither PCL nor Open3D is installed here, so this restates the published file format:

* header lines `FIELDS / SIZE / TYPE / COUNT / WIDTH / HEIGHT / POINTS / DATA`;
* `binary`: `POINTS` records of the packed fields (array of structures);
* `binary_compressed`: `uint32 compressed_size, uint32 uncompressed_size`, then an LZF stream whose decompressed bytes are
  the fields one after another (structure of arrays: all x, then all y, ...);
* LZF (liblzf `lzf_d.c`): a control byte c; c < 32 → copy c+1 literal bytes; otherwise a back reference of length
  (c >> 5) + 2 (a length field of 7 is extended by the next byte) at distance ((c & 31) << 8 | next byte) + 1,
  copied byte by byte (the ranges may overlap).

Only tests, tools/make_real_fixtures.py and bench tooling import this module; the product never does.
"""
import numpy as np

_NP = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 1): "<u1", ("U", 2): "<u2", ("U", 4): "<u4", ("U", 8): "<u8",
       ("I", 1): "<i1", ("I", 2): "<i2", ("I", 4): "<i4", ("I", 8): "<i8"}


def lzf_decompress(src: bytes, out_len: int) -> bytes:
    out = bytearray(out_len)
    ip, op, n = 0, 0, len(src)
    while ip < n:
        c = src[ip]; ip += 1
        if c < 32:
            c += 1
            if op + c > out_len or ip + c > n:
                raise ValueError("LZF: literal run overflows")
            out[op:op + c] = src[ip:ip + c]
            ip += c; op += c
        else:
            ln = c >> 5
            if ln == 7:
                ln += src[ip]; ip += 1
            ref = op - (((c & 31) << 8) | src[ip]) - 1; ip += 1
            ln += 2
            if ref < 0 or op + ln > out_len:
                raise ValueError("LZF: bad back reference")
            if ref + ln <= op:
                out[op:op + ln] = out[ref:ref + ln]
            else:                       # overlapping copy: the pattern of length (op - ref) repeats
                pat = bytes(out[ref:op])
                reps = -(-ln // len(pat))
                out[op:op + ln] = (pat * reps)[:ln]
            op += ln
    if op != out_len:
        raise ValueError(f"LZF: produced {op} bytes, header says {out_len}")
    return bytes(out)


def read_pcd(path):
    """Returns a dict field name -> numpy array of length POINTS (COUNT > 1 fields come back as (POINTS, COUNT))."""
    with open(path, "rb") as f:
        raw = f.read()
    hdr = {}
    pos = 0
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, _, val = line.partition(" ")
        hdr[key] = val.split()
        if key == "DATA":
            break
    fields = hdr["FIELDS"]
    sizes = [int(s) for s in hdr["SIZE"]]
    types = hdr["TYPE"]
    counts = [int(c) for c in hdr.get("COUNT", ["1"] * len(fields))]
    npts = int(hdr["POINTS"][0]) if "POINTS" in hdr else int(hdr["WIDTH"][0]) * int(hdr["HEIGHT"][0])
    dts = [np.dtype(_NP[(t, s)]) for t, s in zip(types, sizes)]
    kind = hdr["DATA"][0]
    body = raw[pos:]
    out = {}
    if kind == "ascii":
        rows = np.array([ln.split() for ln in body.decode("ascii").splitlines() if ln.strip()])
        col = 0
        for name, dt, c in zip(fields, dts, counts):
            out[name] = rows[:npts, col:col + c].astype(np.float64).astype(dt).reshape(npts, c).squeeze(axis=1) if c == 1 \
                else rows[:npts, col:col + c].astype(np.float64).astype(dt)
            col += c
    elif kind == "binary":
        rec = np.dtype({"names": [f"f{i}" for i in range(len(fields))], "formats": [(dt, (c,)) if c > 1 else dt for dt, c in zip(dts, counts)]})
        arr = np.frombuffer(body, dtype=rec, count=npts)
        for i, name in enumerate(fields):
            out[name] = np.array(arr[f"f{i}"])
    elif kind == "binary_compressed":
        csz, usz = np.frombuffer(body[:8], dtype="<u4")
        data = lzf_decompress(body[8:8 + int(csz)], int(usz))
        off = 0
        for name, dt, c in zip(fields, dts, counts):
            nb = dt.itemsize * c * npts
            a = np.frombuffer(data, dtype=dt, count=npts * c, offset=off)
            # PCL writes each field contiguously; COUNT > 1 fields are stored point-major inside their block
            out[name] = np.array(a if c == 1 else a.reshape(npts, c))
            off += nb
        if off != usz:
            raise ValueError(f"PCD: fields cover {off} bytes, stream holds {usz}")
    else:
        raise ValueError(f"PCD: unknown DATA {kind}")
    out["_header"] = {k: v for k, v in hdr.items()}
    return out


def xyz(path, finite_only=True):
    d = read_pcd(path)
    p = np.stack([d["x"], d["y"], d["z"]], axis=1).astype(np.float32)
    if finite_only:
        p = p[np.isfinite(p).all(axis=1)]
    return p
