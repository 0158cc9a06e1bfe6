"""Fixtures produced by EXECUTING the reference's own code (run in the build container, where /root/reference exists).

    python tests/golden/make_reference_fixtures.py [--check]

The reference cannot be imported (its packages need matplotlib, ncut_pytorch, cuml, tensordict and an older
transformers), but four pieces of its hot-path arithmetic are plain torch and can be lifted out of the source files
and run on the CPU unchanged.  This script extracts them by AST / line range, executes them on seeded inputs and
writes tests/golden/ref_*.npz.  Nothing of the reference is copied into the repository: only inputs and outputs.

  ref_attention_mask.npz   MultiStateViTEncoderBackbone._construct_attention_mask(_indices)
                           model/multistate_encoder/modeling_msvitencoder.py:426-467
  ref_attention_stats.npz  transmitter / receiver attention sums of compress_tokens_with_cluster_indices
                           model/multistate_encoder/modeling_msvitencoder.py:169,182-190
  ref_closed_form_ncut.npz the in-repo closed-form NCut (normprod distance, exp, degree, I - D^-1/2 A D^-1/2, eigh)
                           sandbox/test.py:100,106-118
  ref_cluster_means.npz    per-label mean centres and nearest-centre assignment
                           model/clustering/modeling_spectral.py:125-127,129

`--check` regenerates in memory and compares with the committed files (used by tests/test_reference_fixtures.py when
the reference checkout is present).
"""
from __future__ import annotations

import ast
import io
import os
import sys
import textwrap
from typing import Dict, Tuple

import numpy as np
import torch

REF = os.environ.get("MSVIT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
MSVIT_ENC = "model/multistate_encoder/modeling_msvitencoder.py"
SPECTRAL = "model/clustering/modeling_spectral.py"
SANDBOX_TEST = "sandbox/test.py"


def _read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def _lines(rel, first, last, must_contain):
    """Source lines first..last (1-based, inclusive), dedented; `must_contain` guards against a shifted checkout."""
    src = _read(rel).splitlines()[first - 1:last]
    text = textwrap.dedent("\n".join(src))
    for needle in must_contain:
        if needle not in text:
            raise RuntimeError(f"{rel}:{first}-{last} no longer holds {needle!r}; the reference checkout changed")
    return text


def _method_sources(rel, class_name, names):
    """Source of the named methods of a class, re-assembled into a stand-alone class of the same name."""
    src = _read(rel)
    tree = ast.parse(src)
    body = []
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in names:
                    seg = ast.get_source_segment(src, item)
                    deco = "".join(f"@{ast.get_source_segment(src, d)}\n" for d in item.decorator_list)
                    body.append(textwrap.indent(deco + textwrap.dedent(seg), "    "))
    if len(body) != len(names):
        raise RuntimeError(f"{rel}: could not find {names} in class {class_name}")
    return f"class {class_name}:\n" + "\n\n".join(body) + "\n"


# ----------------------------------------------------------------------------------------------------------------
def attention_mask_cases():
    g = torch.Generator().manual_seed(1212)
    cases = {
        "a": torch.tensor([[0, 0, 1, 1, 2], [0, 1, 1, 0, 0]]),                 # unequal cluster counts per image
        "b": torch.randint(0, 4, (3, 17), generator=g),
        "c": torch.zeros(2, 9, dtype=torch.long),                              # the caller's initial state
        "d": torch.stack([torch.randperm(12, generator=g) % 5 for _ in range(2)]),
        "e": torch.arange(6)[None, :].repeat(1, 1),                            # every token its own cluster
    }
    return cases


def gen_attention_mask():
    code = _method_sources(MSVIT_ENC, "MultiStateViTEncoderBackbone",
                           ["_construct_attention_mask_indices", "_construct_attention_mask"])
    ns = {"torch": torch, "Dict": Dict, "Tuple": Tuple}
    exec(compile(code, MSVIT_ENC, "exec"), ns)
    fn = ns["MultiStateViTEncoderBackbone"]._construct_attention_mask
    out = {}
    for name, ci in attention_mask_cases().items():
        out[f"{name}_cluster_indices"] = ci.numpy()
        out[f"{name}_mask"] = fn(ci).numpy()
    return out


def gen_attention_stats():
    head = _lines(MSVIT_ENC, 169, 169, ["n_clusters", "torch.max(cluster_indices)"])
    body = _lines(MSVIT_ENC, 182, 190, ["transmitter_attention_probs", "receiver_attention_probs", "masks"])
    g = torch.Generator().manual_seed(2003)
    out = {}
    for name, (B, H, N, C) in {"a": (2, 3, 12, 4), "b": (1, 2, 33, 5)}.items():
        attn = torch.softmax(torch.randn(B, H, N, N, generator=g), dim=-1)
        ci = torch.randint(0, C, (B, N), generator=g)
        ci[:, :C] = torch.arange(C)                    # every cluster occurs in every image (no 0/0 receiver rows)
        ns = {"torch": torch, "attention_probs": attn, "cluster_indices": ci}
        exec(compile(head + "\n" + body, MSVIT_ENC, "exec"), ns)
        out[f"{name}_attention_probs"] = attn.numpy()
        out[f"{name}_cluster_indices"] = ci.numpy()
        out[f"{name}_transmitter"] = ns["transmitter_attention_probs"].numpy()
        out[f"{name}_receiver"] = ns["receiver_attention_probs"].numpy()
    return out


def gen_closed_form_ncut():
    gamma_line = _lines(SANDBOX_TEST, 100, 100, ["affinity_focal_gamma"])
    body = _lines(SANDBOX_TEST, 106, 118, ["normalized_X", "torch.exp(-A / affinity_focal_gamma)", "torch.linalg.eigh(L)"])
    g = torch.Generator().manual_seed(1212)
    out = {}
    for name, (n, D, K) in {"a": (24, 16, 3), "b": (60, 32, 4)}.items():
        centres = torch.randn(K, D, generator=g, dtype=torch.float64)
        lab = torch.randint(0, K, (n,), generator=g)
        x = (centres[lab] + 0.3 * torch.randn(n, D, generator=g, dtype=torch.float64)) * 0.5
        # tokens exactly representable in TF32 (10 mantissa bits): the CUDA path's tensor cores then read the very
        # numbers the reference code sees, and the comparison is about the arithmetic, not the operand rounding
        xi = x.float().view(torch.int32)
        x = ((xi + 0x1000) & ~0x1FFF).view(torch.float32).double()
        states = torch.cat([torch.zeros(1, 1, D, dtype=torch.float64), x[None]], dim=1)   # [1, 1 + n, D]: CLS slot first
        ns = {"torch": torch, "states": states}
        exec(compile(gamma_line, SANDBOX_TEST, "exec"), ns)
        gamma = float(ns["affinity_focal_gamma"])
        # the lines need the affinity before it is overwritten by eigenvectors: run them one statement at a time
        stmts = [s for s in body.split("\n") if s.strip()]
        A_exp = None
        for st in stmts:
            exec(compile(st, SANDBOX_TEST, "exec"), ns)
            if st.strip().startswith("A = torch.exp"):
                A_exp = ns["A"].clone()
        out[f"{name}_x"] = x.numpy()
        out[f"{name}_gamma"] = np.float64(gamma)
        out[f"{name}_affinity"] = A_exp.numpy()
        out[f"{name}_degree"] = ns["D"].numpy()
        out[f"{name}_laplacian_eigvals"] = ns["E"].numpy()            # ascending eigenvalues of I - D^-1/2 A D^-1/2
        out[f"{name}_eigvecs10"] = ns["X"].numpy()                    # V[:, :10], the reference's embedding
    return out


def gen_cluster_means():
    centres_code = _lines(SPECTRAL, 125, 127, ["cluster_centers", "torch.mean(spectral_x[labels == cluster_idx], dim=0)"])
    assign_code = _lines(SPECTRAL, 129, 129, ["torch.argmin(torch.cdist(spectral_x, cluster_centers), dim=1)"])
    g = torch.Generator().manual_seed(7)
    out = {}
    for name, (n, K) in {"a": (40, 4), "b": (196, 8)}.items():
        proto = torch.randn(K, K, generator=g)
        labels = torch.randint(0, K, (n,), generator=g)
        labels[:K] = torch.arange(K)
        spectral_x = proto[labels] + 0.4 * torch.randn(n, K, generator=g)
        ns = {"torch": torch, "spectral_x": spectral_x, "labels": labels, "n_child_clusters": K, "all_labels": {}}
        exec(compile(centres_code + "\n" + assign_code, SPECTRAL, "exec"), ns)
        out[f"{name}_spectral_x"] = spectral_x.numpy()
        out[f"{name}_labels"] = labels.numpy()
        out[f"{name}_cluster_centers"] = ns["cluster_centers"].numpy()
        out[f"{name}_nearest_centre"] = ns["all_labels"]["km_boosted_spectral"].numpy()
    return out


GENERATORS = {
    "ref_attention_mask.npz": gen_attention_mask,
    "ref_attention_stats.npz": gen_attention_stats,
    "ref_closed_form_ncut.npz": gen_closed_form_ncut,
    "ref_cluster_means.npz": gen_cluster_means,
}


def main():
    check = "--check" in sys.argv
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} does not exist: the fixtures can only be generated next to the reference checkout")
    bad = 0
    for fname, gen in GENERATORS.items():
        data = gen()
        path = os.path.join(HERE, fname)
        if check:
            old = np.load(path)
            same = sorted(old.files) == sorted(data) and all(
                old[k].shape == np.asarray(data[k]).shape and np.allclose(old[k], data[k], rtol=1e-12, atol=0) for k in data)
            print(("ok       " if same else "MISMATCH ") + fname)
            bad += 0 if same else 1
        else:
            np.savez_compressed(path, **data)
            print("wrote", path, {k: tuple(np.asarray(v).shape) for k, v in data.items()})
    raise SystemExit(1 if bad else 0)


if __name__ == "__main__":
    main()
