import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import conos_oracle
    conos_oracle.build()
    return conos_oracle


@pytest.fixture(scope="session")
def ctx():
    import conos_b200
    return conos_b200.default_context(0)


@pytest.fixture(scope="session")
def small_panel():
    """Synthetic 3-sample panel small enough for the oracle (seconds)."""
    from conos_b200.panel import make_panel
    return make_panel(7, 3, [420, 380, 300], n_genes=300, n_types=6, n_pcs=20, use_torch=False)


def to_oracle_samples(oracle, panel):
    return {n: oracle.Sample(counts=s.counts, genes=s.genes, cells=s.cells, varinfo_v=s.varinfo_v,
                             varinfo_qv=s.varinfo_qv, pca=s.pca, od_genes=s.od_genes) for n, s in panel.items()}


def _golden_loader():
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "make_golden_vectors.py")
    spec = importlib.util.spec_from_file_location("make_golden_vectors", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def golden():
    """(module with load_reference_panel / knn_inputs, the committed oracle vectors)"""
    import os
    import numpy as np
    mod = _golden_loader()
    vec = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.npz"))
    return mod, vec
