"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/msvit.h
declares, argument checks fire without a GPU, and the Python mirror keeps the reference's interface."""
import ctypes
import re

import pytest
import torch

import msvit
from msvit import _lib


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def header_symbols():
    text = open(_lib.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msvit_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    declared = header_symbols()
    assert declared, "no declarations parsed from include/msvit.h"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in msvit.h but not exported by libmsvit.so"
    assert sorted(_lib.exported_symbols()) == declared, "ctypes signature table out of sync with msvit.h"
    assert lib.msvit_version() >= 100


def test_error_strings(lib):
    assert lib.msvit_error_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5, -6):
        assert len(lib.msvit_error_string(code)) > 3


def test_argument_checks_fire_before_any_cuda_call(lib):
    # NULL pointers / bad shapes are rejected on the host (no GPU needed)
    assert lib.msvit_affinity_degree(None, 0, None, None, 0, 1, 8, 8, 0, 3.0, 1.0, None, None, None) == -1
    buf = ctypes.create_string_buffer(256)
    p = ctypes.addressof(buf)
    p16 = (p + 15) & ~15
    assert lib.msvit_affinity_degree(p16, 7, None, p16, 8, 1, 8, 8, 0, 3.0, 1.0, None, None, None) == -4  # dtype
    assert lib.msvit_affinity_degree(p16, 0, None, p16, 8, 1, 8, 8, 9, 3.0, 1.0, None, None, None) == -4  # mode
    assert lib.msvit_affinity_degree(p16, 0, None, p16, 9, 1, 8, 8, 0, 3.0, 1.0, None, None, None) == -2  # rows != S*N
    assert lib.msvit_affinity_degree(p16, 1, None, p16, 8, 1, 8, 3, 0, 3.0, 1.0, None, None, None) == -3  # row stride
    assert lib.msvit_affinity_degree(p16 + 4, 0, None, p16, 8, 1, 8, 8, 0, 3.0, 1.0, None, None, None) == -3
    assert lib.msvit_ncut_eig(p16, p16, p16, p16, None, 8, 1, 8, 4, 6, 10, 1e-5, 0.0, 0, None, None, None) == -2  # block % 4
    assert lib.msvit_ncut_eig(p16, p16, p16, p16, None, 8, 1, 8, 40, 40, 10, 1e-5, 0.0, 0, None, None, None) == -2  # block > max
    assert lib.msvit_ncut_eig(p16, p16, p16, p16, None, 8, 1, 8, 4, 8, 10, 1e-5, 0.0, 5, None, None, None) == -2  # n_converge > k
    assert lib.msvit_kmeans(p16, None, None, None, p16, p16, None, 8, 1, 8, 4, 0, 0.1, 10, None, None) == -1  # lam needed
    assert lib.msvit_pool(p16, 0, None, p16, p16, 1, 8, 8, 2, None) == -1
    assert lib.msvit_pool(p16, 0, p16, p16, p16, 1, 8, 8, 0, None) == -2
    assert lib.msvit_compose_labels(p16, p16, None, None, p16, 1, 8, 2, None) == -2  # P > 1 needs seg_off
    assert lib.msvit_gkm_assign(p16, 0, None, p16, None, 8, 2, 8, None) == -1
    assert lib.msvit_gkm_assign(p16, 0, p16, p16, None, 8, 2, 3, None) == -3           # row stride % 16
    assert lib.msvit_gkm_assign(p16, 5, p16, p16, None, 8, 2, 8, None) == -4
    assert lib.msvit_gkm_sort(p16, 8, 2, p16, p16, p16, 0, None) == -5                 # workspace too small
    assert lib.msvit_gkm_sort(p16, 8, 2, p16, p16, None, 64, None) == -1
    assert lib.msvit_gkm_accumulate(p16, 1, p16, p16, p16, 8, 2, 12, None, 0, None) == -2       # D % 8 for bf16
    assert lib.msvit_gkm_finalize(p16, None, None, 0, 2, 8, None) == -1
    assert lib.msvit_gkm_workspace_bytes(4096, 10) == (2 + 1) * 10 * 4   # two block histograms + the label totals
    assert lib.msvit_ncut_fused(p16, 0, p16, p16, p16, p16, p16, 8 * 300, 8, 300, 64, 0, 3.0, 1.0, 16, 10, 1e-5, 0.0, 0, None) == -2  # N > 208
    assert lib.msvit_ncut_fused(p16, 0, p16, p16, p16, p16, p16, 8 * 64, 8, 64, 64, 0, 3.0, 1.0, 24, 10, 1e-5, 0.0, 0, None) == -2   # block != 16
    assert lib.msvit_discretise(p16, p16, None, None, p16, p16, None, 8, 1, 8, 4, 2, 0.1, 10, 7, None, None) == -4                  # method
    assert lib.msvit_gkm_assign(p16, 0, p16, p16, None, 0, 2, 8, None) == 0
    assert lib.msvit_attention_mask(None, p16, 1, 8, 2, None) == -1
    assert lib.msvit_attention_mask(p16, p16, 1, 0, 2, None) == -2
    assert lib.msvit_attention_mask(p16, p16 + 1, 1, 8, 2, None) == -3
    assert lib.msvit_attention_mask(p16, p16, 0, 8, 2, None) == 0
    assert lib.msvit_cluster_key_sums(p16, p16, None, 1, 1, 8, 2, None) == -1
    assert lib.msvit_cluster_key_sums(p16, p16, p16, 1, 1, 8, 70, None) == -2         # C > 64
    assert lib.msvit_cluster_key_sums(p16 + 4, p16, p16, 1, 1, 8, 2, None) == -3
    assert lib.msvit_cluster_key_sums(p16, p16, p16, 0, 1, 8, 2, None) == 0
    assert lib.msvit_cluster_attention_stats(p16, p16, p16, None, 1, 1, 8, 2, None) == -1
    assert lib.msvit_cluster_attention_stats(p16, p16, p16, p16, 1, 1, 2048, 2, None) == -2     # N > 256
    assert lib.msvit_cluster_attention_stats(p16, p16, p16, p16, 1, 1, 301, 2, None) == -2      # N > 256
    assert lib.msvit_cluster_attention_stats(p16 + 4, p16, p16, p16, 1, 1, 8, 2, None) == -3
    assert lib.msvit_cluster_attention_stats(p16, p16, p16, p16, 1, 1, 8, 17, None) == -2       # C > 16
    assert lib.msvit_cluster_attention_stats(p16, p16, p16, p16, 0, 1, 8, 2, None) == 0
    # empty batches are a no-op
    assert lib.msvit_affinity_degree(p16, 0, None, p16, 0, 0, 8, 8, 0, 3.0, 1.0, None, None, None) == 0
    assert lib.msvit_pool(p16, 0, p16, p16, p16, 0, 8, 8, 2, None) == 0


def test_plugin_interface_mirrors_reference():
    # model/clustering/__init__.py:7-10, modeling.py:12-36, modeling_spectral.py:42-47
    cfg = msvit.SpectralClusteringConfig(ncut_dim=8, ncut_dist="rbf", eigenvalue_threshold=0.1,
                                         cluster_size_threshold=0.07)
    assert cfg.model_type == "spectral" and cfg.ncut_dim == 8 and cfg.affinity_focal_gamma == 3.0
    mod = msvit.CLUSTERING_CLASSES[cfg.model_type](cfg)
    assert isinstance(mod, msvit.ClusteringModule) and isinstance(mod, torch.nn.Module)
    assert len(list(mod.parameters())) == 0  # stateless: nothing to checkpoint
    with pytest.raises(NotImplementedError):
        msvit.ClusteringModule()(torch.zeros(1, 4, dtype=torch.long), torch.zeros(1, 4, 8))


def test_no_cpu_fallback():
    cfg = msvit.SpectralClusteringConfig(ncut_dim=4, ncut_dist="rbf", eigenvalue_threshold=0.1)
    mod = msvit.SpectralClustering(cfg)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mod(torch.zeros(2, 16, dtype=torch.long), torch.randn(2, 16, 32))
    with pytest.raises(RuntimeError):
        msvit.pool(torch.randn(1, 4, 8), torch.zeros(1, 4, dtype=torch.long), 2)


def test_block_width_rule():
    from msvit.functional import default_block
    assert default_block(8) == 16 and default_block(16) == 24 and default_block(2) == 12
    assert default_block(30) == 32 and default_block(32) == 32
    assert default_block(33) == 0 and default_block(100) == 0     # dense solver (ncut_dim up to 128)
    with pytest.raises(ValueError):
        default_block(129)
