"""Minimal PCD v0.7 reader/writer (ASCII and `DATA binary`), enough for the bundled Velodyne fixtures
(thirdparty/fast_gicp/data/*.pcd: FIELDS x y z intensity, float32) and for save_map-style dumps."""
import numpy as np


def read_pcd(path):
    """Returns an (N, 4) float32 array: x, y, z, intensity (0 when the file has no intensity field)."""
    with open(path, "rb") as f:
        fields, sizes, types, counts, npts, data = [], [], [], [], None, None
        while True:
            line = f.readline()
            if not line:
                raise ValueError("PCD header ended unexpectedly")
            tok = line.decode("ascii", "replace").strip().split()
            if not tok or tok[0].startswith("#"):
                continue
            key = tok[0].upper()
            if key == "FIELDS":
                fields = tok[1:]
            elif key == "SIZE":
                sizes = [int(t) for t in tok[1:]]
            elif key == "TYPE":
                types = tok[1:]
            elif key == "COUNT":
                counts = [int(t) for t in tok[1:]]
            elif key == "POINTS":
                npts = int(tok[1])
            elif key == "DATA":
                data = tok[1].lower()
                break
        if not counts:
            counts = [1] * len(fields)
        if data == "binary":
            dt = []
            for name, s, t, c in zip(fields, sizes, types, counts):
                code = {"F": "f", "I": "i", "U": "u"}[t.upper()] + str(s)
                dt.append((name, "<" + code, (c,)) if c > 1 else (name, "<" + code))
            rec = np.frombuffer(f.read(npts * np.dtype(dt).itemsize), dtype=np.dtype(dt), count=npts)
            cols = {n: rec[n].astype(np.float32) for n in fields}
        elif data == "ascii":
            arr = np.loadtxt(f, dtype=np.float64, ndmin=2)
            cols = {n: arr[:, i].astype(np.float32) for i, n in enumerate(fields)}
        else:
            raise ValueError("unsupported PCD DATA mode: %s" % data)
    out = np.zeros((npts, 4), np.float32)
    for i, n in enumerate(("x", "y", "z")):
        out[:, i] = cols[n]
    if "intensity" in cols:
        out[:, 3] = cols["intensity"]
    return out


def write_pcd(path, pts):
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    n = pts.shape[0]
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\n"
           "COUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (n, n))
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(pts.tobytes())
