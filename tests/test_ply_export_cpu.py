"""export_ply (vicasplat_b200.ply_export <-> src/model/ply_export.py:31-90) against the records the UNMODIFIED
reference produces for the same seeded Gaussians (tests/golden/ply_export.npz), through the written file."""
from pathlib import Path

import numpy as np
import torch

GOLD = Path(__file__).parent / "golden" / "ply_export.npz"


def test_written_file_equals_the_reference_records(tmp_path):
    from oracle.make_ply_golden import seeded_gaussians
    from vicasplat_b200.ply_export import export_ply, read_ply
    gold = np.load(GOLD)
    means, scales, rot, sh, opac = seeded_gaussians()
    for tag, dc_only in (("full", False), ("dc", True)):
        path = tmp_path / "sub" / f"{tag}.ply"
        export_ply(torch.eye(4), means, scales, rot, sh, opac, path, save_sh_dc_only=dc_only)
        names, rec = read_ply(path)
        assert names == list(gold[f"{tag}/names"])
        want = gold[f"{tag}/records"]
        assert rec.shape == want.shape
        assert np.allclose(rec, want, rtol=1e-6, atol=1e-6)
        assert (np.diff(rec[:, names.index("opacity")]) <= 0).all()          # sorted by descending opacity
    # pruning: nothing below the threshold survives (logit(0.005) = -5.29)
    assert rec[:, names.index("opacity")].min() >= np.log(0.005 / 0.995) - 1e-5
    raw = path.read_bytes()
    assert raw.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex ")
