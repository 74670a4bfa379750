"""``export_ply`` <-> src/model/ply_export.py:31-90 (the on-disk format downstream of
the hot path): the scene a viewer loads -- Gaussians pruned at opacity < 0.005, sorted by descending
opacity, positions, zero normals, SH DC (+ rest) coefficients, logit opacity, log scales and wxyz rotations
as little-endian float32 vertex records of a binary PLY file (the layout 3DGS viewers expect).

Pruning, the sort and the gathers run on the device the tensors live on (torch); only the surviving
records travel to the host, where the file is written in one piece.  No plyfile dependency: the header
and the record layout are written directly (plyfile's binary_little_endian output for an all-'f4'
structured array is exactly header + raw records)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch


def construct_list_of_attributes(num_rest: int) -> list:
    attributes = ["x", "y", "z", "nx", "ny", "nz"]
    attributes += [f"f_dc_{i}" for i in range(3)]
    attributes += [f"f_rest_{i}" for i in range(num_rest)]
    attributes.append("opacity")
    attributes += [f"scale_{i}" for i in range(3)]
    attributes += [f"rot_{i}" for i in range(4)]
    return attributes


def _canonical_wxyz(q_xyzw: torch.Tensor) -> np.ndarray:
    """The reference sends the quaternions through scipy's Rotation (from_quat -> as_matrix -> from_matrix ->
    as_quat, ply_export.py:51-55): unit length, the sign scipy's matrix -> quaternion conversion picks."""
    from scipy.spatial.transform import Rotation as R
    q = R.from_matrix(R.from_quat(q_xyzw.detach().cpu().numpy()).as_matrix()).as_quat()
    return np.stack((q[:, 3], q[:, 0], q[:, 1], q[:, 2]), axis=-1)


@torch.no_grad()
def export_records(means, scales, rotations, harmonics, opacities, save_sh_dc_only: bool = False):
    """-> (attribute names, float32 array (n_kept, n_attributes)) in file order."""
    mask = opacities >= 0.005
    op = opacities[mask]
    op, idx = torch.sort(op, descending=True)
    take = lambda t: t[mask][idx]
    means, scales, rotations, harmonics = take(means), take(scales), take(rotations), take(harmonics)
    f_dc = harmonics[..., 0]
    f_rest = harmonics[..., 1:].flatten(start_dim=1)
    cols = [means, torch.zeros_like(means), f_dc]
    if not save_sh_dc_only:
        cols.append(f_rest)
    cols += [torch.log(op / (1 - op))[:, None], scales.log()]
    left = torch.cat([c.to(torch.float32) for c in cols], dim=1).cpu().numpy()
    rec = np.concatenate([left, _canonical_wxyz(rotations).astype(np.float32)], axis=1)
    return construct_list_of_attributes(0 if save_sh_dc_only else f_rest.shape[1]), np.ascontiguousarray(rec, dtype="<f4")


def export_ply(extrinsics, means, scales, rotations, harmonics, opacities, path, save_sh_dc_only: bool = False) -> None:
    """Same signature as the reference (``extrinsics`` is accepted and unused there too)."""
    names, rec = export_records(means, scales, rotations, harmonics, opacities, save_sh_dc_only)
    path = Path(path)
    path.parent.mkdir(exist_ok=True, parents=True)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {rec.shape[0]}"]
    header += [f"property float {n}" for n in names] + ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(rec.tobytes())


def read_ply(path):
    """Minimal reader of the files export_ply writes: -> (names, float32 array (n, len(names)))."""
    data = Path(path).read_bytes()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").splitlines()
    assert lines[0] == "ply" and lines[1] == "format binary_little_endian 1.0"
    n = int(lines[2].split()[-1])
    names = [ln.split()[-1] for ln in lines[3:-1]]
    return names, np.frombuffer(data[end:], dtype="<f4").reshape(n, len(names))
