"""Writers of synthetic .ply files for the scene-ingest tests (standard 3DGS layout with arbitrary property names / types,
and the PlayCanvas / splat-transform compressed layout). Test infrastructure only."""
import numpy as np

_NP = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1", "char": "i1",
       "short": "<i2", "ushort": "<u2", "int": "<i4", "uint": "<u4"}


def write_ply(props, columns, count=None, eol="\n", extra_header=(), fmt="binary_little_endian", element="vertex"):
    """props: [(type, name)], columns: {name: array}. Returns the file as bytes."""
    n = len(next(iter(columns.values()))) if count is None else count
    head = ["ply", f"format {fmt} 1.0"] + list(extra_header) + [f"element {element} {n}"]
    head += [f"property {t} {name}" for t, name in props] + ["end_header"]
    dt = np.dtype([(name, _NP[t]) for t, name in props])
    body = np.zeros(n, dt)
    for _, name in props:
        body[name] = columns[name][:n]
    return (eol.join(head) + eol).encode() + body.tobytes()


def standard_scene(n=1000, sh_degree=3, seed=0, log_scale=True, logit_opacity=True, placeholders=0, dtype="float"):
    rng = np.random.default_rng(seed)
    k = (sh_degree + 1) ** 2
    cols = {"x": rng.normal(0, 2, n) + 3.0, "y": rng.normal(0, 1, n) - 1.5, "z": rng.normal(0, 3, n) + 7.0,
            "nx": np.zeros(n), "ny": np.zeros(n), "nz": np.zeros(n)}
    props = [(dtype, "x"), (dtype, "y"), (dtype, "z"), ("float", "nx"), ("float", "ny"), ("float", "nz")]
    for c in range(3):
        cols[f"f_dc_{c}"] = rng.normal(0, 1, n)
        props.append(("float", f"f_dc_{c}"))
    for c in range(3 * (k - 1)):
        cols[f"f_rest_{c}"] = rng.normal(0, 0.2, n)
        props.append(("float", f"f_rest_{c}"))
    cols["opacity"] = rng.normal(0, 2, n) if logit_opacity else rng.uniform(0.05, 1.0, n)
    props.append(("float", "opacity"))
    for c in range(3):
        cols[f"scale_{c}"] = rng.normal(-4, 1, n) if log_scale else rng.uniform(0.001, 0.2, n)
        props.append(("float", f"scale_{c}"))
    q = rng.normal(0, 1, (n, 4))
    for c in range(4):
        cols[f"rot_{c}"] = q[:, c]
        props.append(("float", f"rot_{c}"))
    if placeholders:
        idx = rng.choice(n, placeholders, replace=False)
        for c in range(3):
            cols[f"scale_{c}"][idx] = 2.0
        cols["opacity"][idx] = 4.8402
    return props, cols


def compressed_scene(n=1000, seed=0, with_sh_element=False):
    """chunk element (18 floats per 256 vertices) + 4 packed uint32 per vertex (+ an optional trailing sh element)."""
    rng = np.random.default_rng(seed)
    nch = (n + 255) // 256
    cnames = ["min_x", "min_y", "min_z", "max_x", "max_y", "max_z", "min_scale_x", "min_scale_y", "min_scale_z",
              "max_scale_x", "max_scale_y", "max_scale_z", "min_r", "min_g", "min_b", "max_r", "max_g", "max_b"]
    ch = np.zeros(nch, np.dtype([(c, "<f4") for c in cnames]))
    lo = rng.normal(0, 3, (nch, 3)); hi = lo + rng.uniform(0.5, 2.0, (nch, 3))
    slo = rng.normal(-5, 0.5, (nch, 3)); shi = slo + rng.uniform(0.5, 2.0, (nch, 3))
    clo = rng.uniform(0.0, 0.4, (nch, 3)); chi = clo + rng.uniform(0.1, 0.6, (nch, 3))
    for k, ax in enumerate("xyz"):
        ch[f"min_{ax}"] = lo[:, k]; ch[f"max_{ax}"] = hi[:, k]
        ch[f"min_scale_{ax}"] = slo[:, k]; ch[f"max_scale_{ax}"] = shi[:, k]
    for k, cc in enumerate("rgb"):
        ch[f"min_{cc}"] = clo[:, k]; ch[f"max_{cc}"] = chi[:, k]
    vx = np.zeros(n, np.dtype([(c, "<u4") for c in ("packed_position", "packed_rotation", "packed_scale", "packed_color")]))
    for c in vx.dtype.names:
        vx[c] = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    head = ["ply", "format binary_little_endian 1.0", "comment splat-transform style", f"element chunk {nch}"]
    head += [f"property float {c}" for c in cnames] + [f"element vertex {n}"]
    head += [f"property uint {c}" for c in vx.dtype.names]
    tail = b""
    if with_sh_element:
        head += [f"element sh {n}"] + [f"property uchar f_rest_{i}" for i in range(9)]
        tail = rng.integers(0, 256, (n, 9), dtype=np.uint8).tobytes()
    head += ["end_header"]
    return ("\n".join(head) + "\n").encode() + ch.tobytes() + vx.tobytes() + tail
