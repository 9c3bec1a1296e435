"""The reference's on-disk point-cloud format, without the `plyfile` dependency.

`GaussianModel.save_ply` / `load_ply` (scene/gaussian_model.py:890-916, 965-1005) write/read ONE binary little-endian
PLY element `vertex` whose properties are all `float` (f4), in this order:
    x y z  nx ny nz  f_dc_0..2  f_rest_0..(3*((D+1)^2-1)-1)  opacity  scale_0..2  rot_0..3
with the SH features stored channel-major (the reference flattens `features.transpose(1, 2)`), the normals zero,
and every tensor PRE-activation (log-scales, opacity logits, un-normalised quaternions).  `plyfile` emits the header
below byte for byte (format line, element line, one `property float <name>` per attribute, `end_header`).  This module
covers the PLY element only: the reference's `load_ply` also torch.load()s five network state-dicts from beside the
file (scene/gaussian_model.py:1014+) - `checkpoint_io.save_point_cloud` writes the pair the reference (and the SIBR
viewer, which reads the PLY alone) expects.  `double` properties, which other writers emit, are accepted on load.

Host-side IO only (numpy); nothing here is on the hot path.
"""
import os

import numpy as np
import torch


def attribute_names(num_dc=3, num_rest=45, num_scale=3, num_rot=4):
    """Property order of gaussian_model.construct_list_of_attributes (scene/gaussian_model.py:890-903)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += ["f_dc_%d" % i for i in range(num_dc)]
    names += ["f_rest_%d" % i for i in range(num_rest)]
    names.append("opacity")
    names += ["scale_%d" % i for i in range(num_scale)]
    names += ["rot_%d" % i for i in range(num_rot)]
    return names


def _header(n, names):
    lines = ["ply", "format binary_little_endian 1.0", "element vertex %d" % n]
    lines += ["property float %s" % a for a in names]
    lines.append("end_header")
    return ("\n".join(lines) + "\n").encode("ascii")


def save_ply(path, xyz, features_dc, features_rest, opacity, scaling, rotation):
    """xyz [P,3], features_dc [P,1,3], features_rest [P,K,3] (K = (D+1)^2 - 1), opacity [P,1], scaling [P,3],
    rotation [P,4] - the reference's parameter tensors as they are (pre-activation)."""
    def host(t):
        return t.detach().to("cpu", torch.float32).contiguous()
    xyz, opacity, scaling, rotation = host(xyz), host(opacity), host(scaling), host(rotation)
    P = xyz.shape[0]
    f_dc = host(features_dc).transpose(1, 2).reshape(P, -1)          # channel-major, gaussian_model.py:910
    f_rest = host(features_rest).transpose(1, 2).reshape(P, -1)
    table = torch.cat([xyz, torch.zeros_like(xyz), f_dc, f_rest, opacity.reshape(P, -1), scaling, rotation], dim=1)
    names = attribute_names(f_dc.shape[1], f_rest.shape[1], scaling.shape[1], rotation.shape[1])
    if table.shape[1] != len(names):
        raise ValueError("save_ply: %d columns for %d attributes" % (table.shape[1], len(names)))
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)                                   # utils/system_utils.mkdir_p
    with open(path, "wb") as f:
        f.write(_header(P, names))
        f.write(table.numpy().astype("<f4", copy=False).tobytes())


def _read_header(f):
    if f.readline().strip() != b"ply":
        raise ValueError("not a PLY file")
    fmt, count, props, types, in_vertex = None, None, [], [], False
    while True:
        line = f.readline()
        if not line:
            raise ValueError("PLY header not terminated")
        tok = line.decode("ascii").split()
        if not tok or tok[0] in ("comment", "obj_info"):
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            in_vertex = tok[1] == "vertex"
            if in_vertex:
                count = int(tok[2])
            elif count is not None:
                raise ValueError("only a single vertex element is supported")
        elif tok[0] == "property" and in_vertex:
            if tok[1] not in _PLY_TYPES:
                raise ValueError("vertex property %s has unsupported type %s" % (tok[-1], tok[1]))
            props.append(tok[2])
            types.append(_PLY_TYPES[tok[1]])
        elif tok[0] == "end_header":
            break
    if fmt not in ("binary_little_endian", "binary_big_endian", "ascii") or count is None:
        raise ValueError("unsupported PLY header")
    return fmt, count, props, types


# scalar property types plyfile accepts for these attributes (the reference writes 'f4'; other writers use double)
_PLY_TYPES = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def load_ply(path, max_sh_degree=3, device="cpu"):
    """Returns the reference's parameter tensors (gaussian_model.load_ply, :965-1005): xyz [P,3], features_dc [P,1,3],
    features_rest [P,K,3], opacity [P,1], scaling [P,3], rotation [P,4], float32 on `device`."""
    with open(path, "rb") as f:
        fmt, n, props, types = _read_header(f)
        if fmt == "ascii":
            data = np.loadtxt(f, dtype=np.float64, ndmin=2)[:n].astype(np.float32)
            if data.shape != (n, len(props)):
                raise ValueError("load_ply: ASCII body has shape %s, header promises %d x %d" % (data.shape, n, len(props)))
        else:
            order = "<" if fmt == "binary_little_endian" else ">"
            rec = np.dtype([(p, order + t) for p, t in zip(props, types)])
            buf = f.read(rec.itemsize * n)
            if len(buf) != rec.itemsize * n:
                raise ValueError("load_ply: file is truncated (%d of %d body bytes for %d vertices x %d properties)"
                                 % (len(buf), rec.itemsize * n, n, len(props)))
            table = np.frombuffer(buf, dtype=rec)
            data = np.stack([table[p].astype(np.float32) for p in props], axis=1) if props else np.zeros((n, 0), np.float32)
    col = {name: i for i, name in enumerate(props)}

    def cols(prefix):
        names = sorted((p for p in props if p.startswith(prefix)), key=lambda s: int(s.split("_")[-1]))
        return data[:, [col[p] for p in names]]
    xyz = data[:, [col["x"], col["y"], col["z"]]]
    f_dc = cols("f_dc_")
    f_rest = cols("f_rest_")
    K = (max_sh_degree + 1) ** 2 - 1
    if f_dc.shape[1] != 3 or f_rest.shape[1] != 3 * K:
        raise ValueError("load_ply: expected 3 f_dc and %d f_rest properties for SH degree %d" % (3 * K, max_sh_degree))
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=device)
    return {
        "xyz": t(xyz),
        "features_dc": t(f_dc.reshape(n, 3, 1)).transpose(1, 2).contiguous(),
        "features_rest": t(f_rest.reshape(n, 3, K)).transpose(1, 2).contiguous(),
        "opacity": t(data[:, [col["opacity"]]]),
        "scaling": t(cols("scale_")),
        "rotation": t(cols("rot")),
    }
