import importlib
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) where there is no CUDA device or the library has not been built."""
    try:
        have_gpu = importlib.import_module("spada-sim_b200").device_count() > 0
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and the built libspada_b200.so")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def spada():
    """The product package (directory name has a hyphen)."""
    return importlib.import_module("spada-sim_b200")


@pytest.fixture(scope="session")
def oracle():
    """CPU restatement (test infrastructure only)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as o  # noqa
    o.build()
    return o


@pytest.fixture(scope="session")
def cari():
    z = np.load(os.path.join(GOLDEN, "cari_csr.npz"))
    a = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
    return a


@pytest.fixture(scope="session", params=["fused", "two_phase"])
def engine(spada, request):
    """Both engine modes: single pass (fused light rows + look-back) and the scratch pass (rows sorted once, then placed)."""
    if spada.device_count() == 0:
        pytest.skip("no CUDA device")
    e = spada.Engine(two_phase=(request.param == "two_phase"), single_pass=(request.param == "fused"))
    yield e
    e.close()


def random_csr(m, n, density=None, row_nnz=None, seed=0, values="uniform"):
    """Canonical CSR with seeded structure.  row_nnz: int or array of per-row counts."""
    rng = np.random.default_rng(seed)
    if row_nnz is None:
        a = sp.random(m, n, density=density, format="csr", random_state=rng, dtype=np.float64)
    else:
        counts = np.broadcast_to(np.asarray(row_nnz, dtype=np.int64), (m,)).copy()
        counts = np.minimum(counts, n)
        rows = np.repeat(np.arange(m), counts)
        cols = np.concatenate([rng.choice(n, size=c, replace=False) if c else np.empty(0, dtype=np.int64)
                               for c in counts]) if m else np.empty(0, dtype=np.int64)
        vals = np.ones(len(rows))
        a = sp.coo_matrix((vals, (rows, cols)), shape=(m, n)).tocsr()
    a.sum_duplicates()
    a.sort_indices()
    if values == "uniform":
        a.data = rng.uniform(0.001, 1.0, size=a.nnz)
    elif values == "signed":
        a.data = rng.uniform(-1.0, 1.0, size=a.nnz)
    elif values == "int":
        a.data = rng.integers(-3, 4, size=a.nnz).astype(np.float64)
    return a
