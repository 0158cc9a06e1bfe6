"""Generates the golden vectors under tests/golden/ from the CPU oracle (fp64).

    python tests/golden/make_golden.py

PARITY UNPINNED: the reference records no expected outputs for this path and its arithmetic
(ncut-pytorch==1.7.9, cuml~=24.10) is absent from this image, so these vectors pin the ORACLE
(oracle/ncut_oracle.py), not the reference.  Shapes follow SURVEY.md section 8c:
(8, 2, 2) is the reference's own smoke shape and seed (sandbox/ncut_euclidean.py:13-21); the
others are the BASELINE.json configs (one image each, seeds 1212 + b).
Inputs are regenerated from the seed by the tests; only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-state-vit_b200"))

from oracle import ncut_oracle as O  # noqa: E402
from msvit.synthetic import default_scale, planted_image  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = [  # name, N, D, K planted, k eig, image index
    ("c2_196x768", 196, 768, 8, 8, 0),
    ("c2_196x768_b5", 196, 768, 8, 8, 5),
    ("c3_576x1024", 576, 1024, 16, 16, 0),
    ("c4_1024x768", 1024, 768, 4, 8, 0),
]


def smoke_8x2():
    # sandbox/ncut_euclidean.py:13-16: torch.manual_seed(1212); M = torch.randn((8, 2))
    torch.manual_seed(1212)
    M = torch.randn((8, 2))
    A = O.affinity(M.double(), "rbf", 3.0, scale=1.0)
    V, lam, deg = O.ncut_eig(A, 2)
    labels, centres, C = O.kmeans(V[:, :2], 2, weight=deg)
    # rbf on raw M == cosine on normalised M (ncut_euclidean.py:23-29)
    Mn = torch.nn.functional.normalize(M.double(), dim=-1)
    A_rbf_n = O.affinity(Mn, "rbf", 3.0, scale=1.0)
    A_cos_n = O.affinity(Mn, "cosine", 3.0)
    assert torch.allclose(A_rbf_n, A_cos_n, atol=1e-12)
    np.savez(os.path.join(OUT, "smoke_8x2.npz"), M=M.numpy(), A=A.numpy(), V=V.numpy(), lam=lam.numpy(),
             deg=deg.numpy(), labels=labels.numpy())


def case(name, N, D, K, k, b):
    x, planted = planted_image(b, N, D, K)
    s = default_scale(D)
    xd = x.double()
    A = O.affinity(xd, "rbf", 3.0, s)
    V, lam, deg = O.ncut_eig(A, k + 4)
    labels, centres, C = O.kmeans(V[:, :K], K, weight=deg)
    pooled, counts = O.pool(xd[None], labels[None], K)
    np.savez(os.path.join(OUT, f"{name}.npz"),
             N=N, D=D, K=K, k=k, b=b, scale=s, gamma=3.0,
             lam=lam.numpy(),                     # k + 4 leading eigenvalues (fp64)
             V=V[:, :k].numpy().astype(np.float32),
             deg=deg.numpy().astype(np.float32),
             A_sum=float(A.sum()), A_diag_mean=float(A.diagonal().mean()),
             A_sample=A[:4, :8].numpy(),
             labels=labels.numpy().astype(np.int16), n_child=C,
             planted=planted.numpy().astype(np.int16),
             pooled_sample=pooled[0, :, :16].numpy().astype(np.float32), counts=counts[0].numpy())


if __name__ == "__main__":
    smoke_8x2()
    for c in CASES:
        case(*c)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
