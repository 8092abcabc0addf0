"""Template meshes of the reference's meshing module, held as arrays.

The reference re-parses modules/meshing/objects/{sphere,cuboid}.obj from disk once per (primitive,
sample) every iteration (meshing/sphere.py:14,30-36; cuboid.py:14,29-33).  Here a template is parsed
once and cached per device.  assets/templates.npz holds the three OBJ assets of the reference as
arrays (written by oracle/make_golden.py); `parse_obj` reads any other OBJ a caller supplies.
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import numpy as np
import torch

_ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets", "templates.npz")
_cache: Dict[Tuple[str, str], Tuple[torch.Tensor, torch.Tensor]] = {}
_raw = None


def parse_obj(path: str) -> Tuple[np.ndarray, np.ndarray]:
    """'v x y z' and 'f a b c' / 'f a//n b//n c//n' rows -> float32 (V,3), int32 (F,3) zero-based."""
    vs, fs = [], []
    with open(path) as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                vs.append([float(x) for x in tok[1:4]])
            elif tok[0] == "f":
                fs.append([int(x.split("/")[0]) - 1 for x in tok[1:4]])
    return np.asarray(vs, dtype=np.float32), np.asarray(fs, dtype=np.int32)


def raw_template(name: str) -> Tuple[np.ndarray, np.ndarray]:
    """name in {'sphere', 'cuboid', 'sphere386'}: vertices as stored in the OBJ, faces int32."""
    global _raw
    if _raw is None:
        _raw = dict(np.load(_ASSETS))
    return _raw[name + "_vertices"], _raw[name + "_faces"]


def template(name: str, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Template as the meshing code uses it: 'sphere' is zero-centred and divided by its mean vertex
    radius (meshing/sphere.py:33-34); 'cuboid' and 'sphere386' are used as stored."""
    key = (name, str(device))
    if key not in _cache:
        v, f = raw_template(name)
        vt = torch.from_numpy(v.copy())
        if name == "sphere":
            vt = vt - torch.mean(vt, 0)
            vt = vt / torch.mean(torch.norm(vt, dim=1))
        _cache[key] = (vt.to(device).contiguous(), torch.from_numpy(f.copy()).to(device).contiguous())
    return _cache[key]
