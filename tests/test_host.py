"""Host-side mirror of the reference interface: loaders, GEMM, storage types, CLI grammar."""
import io
import json
import os
import pickle
from contextlib import redirect_stdout

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

from conftest import GOLDEN, random_csr


def test_rust_debug_float_format(spada):
    import importlib
    r = importlib.import_module("spada-sim_b200.rustfmt")
    cases = {2.14086603168: "2.14086603168", 1.0: "1.0", 0.1: "0.1", 1e-5: "1e-5", 1.5e16: "1.5e16", 1e16: "1e16",
             123456789012345.0: "123456789012345.0", 1e-4: "0.0001", -0.5: "-0.5", 0.0: "0.0", 1e300: "1e300",
             0.12834653759999964: "0.12834653759999964", 4.0: "4.0", -1.0: "-1.0"}
    for v, s in cases.items():
        assert r.debug_f64(v) == s
    assert r.debug_list(np.array([0, 1, 2], dtype=np.uint64)) == "[0, 1, 2]"
    assert r.debug_list(np.array([1.0, 0.5])) == "[1.0, 0.5]"


def test_gemm_from_mat_square_and_rect(spada):
    sq = random_csr(30, 30, density=0.2, seed=1)
    g = spada.GEMM.from_mat("sq", sq)
    assert g.b is g.a                                    # square => A x A (gemm.rs:42-43)
    rc = random_csr(20, 50, density=0.2, seed=2)
    g = spada.GEMM.from_mat("rc", rc)
    assert g.b.shape == (50, 20)                         # otherwise A x A^T as CSR (gemm.rs:44-46)
    assert (g.b != rc.T.tocsr()).nnz == 0 and g.b.has_sorted_indices


def test_gemm_display_matches_reference_quirk(spada, cari):
    g = spada.GEMM.from_mat("cari", cari)
    known = json.load(open(os.path.join(GOLDEN, "cari_known_answers.json")))
    lines = str(g).split("\n")
    assert lines[0] == "---- cari ----" and lines[1] == "--A: (400, 1200)" and lines[5] == "--B: (1200, 400)"
    assert lines[2] == "data: [0.052016, 0.038102, 0.024245, 0.01071, 0.003553] .. "
    assert lines[3] == "indices: [0, 1, 2, 3, 4] ..." and lines[4] == "indptr: [0, 382, 764, 1146, 1528] ..."
    assert lines[6] == "data: [0.052016, 0.038102, 0.024245, 0.01071, 0.003553] ..."   # A's data under B (gemm.rs:79)
    assert lines[7] == "indices: [0, 1, 2, 3, 4] ..."
    assert lines[8].startswith("indptr: [0, ") and known["a_head"]["indptr"] == [0, 382, 764, 1146, 1528]


@pytest.mark.parametrize("field,symmetry", [("real", "general"), ("real", "symmetric"), ("integer", "general"),
                                            ("pattern", "general")])
def test_load_mm_mat(spada, tmp_path, field, symmetry):
    m = random_csr(12, 12, density=0.3, seed=3, values="int" if field == "integer" else "uniform")
    if symmetry == "symmetric":
        m = (m + m.T).tocsr()
    if field == "pattern":
        m.data[:] = 1.0
    scipy.io.mmwrite(str(tmp_path / "w.mtx"), m.astype(np.int64) if field == "integer" else m, field=field,
                     symmetry=symmetry)
    buf = io.StringIO()
    with redirect_stdout(buf):
        got = spada.load_mm_mat(str(tmp_path), "w")
    assert buf.getvalue().splitlines() == ["---- Python Interface ----", f"% Load w from {tmp_path}"]
    assert got.dtype == np.float64 and got.has_canonical_format
    assert (got != m).nnz == 0


def test_load_mm_mat_array_format_errors_like_reference(spada, tmp_path):
    scipy.io.mmwrite(str(tmp_path / "d.mtx"), np.eye(3))
    with pytest.raises(AttributeError), redirect_stdout(io.StringIO()):
        spada.load_mm_mat(str(tmp_path), "d")


def test_load_pickled_gemms(spada, tmp_path):
    a = random_csr(8, 6, density=0.4, seed=4)
    b = random_csr(6, 9, density=0.4, seed=5)
    unsorted = sp.csr_matrix((np.array([1.0, 2.0, 3.0]), np.array([2, 0, 0]), np.array([0, 3])), shape=(1, 6))
    d = {"csr": (a, b), "csc": (a.tocsc(), b.tocsc()), "coo": (a.tocoo(), b.tocoo()), "nd": (a.toarray(), b.toarray()),
         "bad": (a.tolil(), b), "raw": (unsorted, b)}
    fp = str(tmp_path / "nn_gemm.pkl")
    pickle.dump(d, open(fp, "wb"))
    for key in ("csr", "csc", "coo", "nd"):
        buf = io.StringIO()
        with redirect_stdout(buf):
            g = spada.load_pickled_gemms(fp, key)
        out = buf.getvalue().splitlines()
        assert out[0] == "---- Python Interface ----" and out[1] == f"% Load {key} from {fp}"
        assert out[2] == "% -- A --" and out[4] == "% -- B --" and out[6] == "--- Return from Python Interface ---"
        assert (g.a != a).nnz == 0 and (g.b != b).nnz == 0 and g.name == key
    with pytest.raises(TypeError, match="Unsupported matrix type"), redirect_stdout(io.StringIO()):
        spada.load_pickled_gemms(fp, "bad")
    with redirect_stdout(io.StringIO()):
        g = spada.load_pickled_gemms(fp, "raw")          # canonicalised: duplicates summed, columns sorted
    assert g.a.indices.tolist() == [0, 2] and g.a.data.tolist() == [5.0, 1.0]


def test_storage_types(spada, cari):
    g = spada.GEMM.from_mat("cari", cari)
    sa, sb = spada.CsrMatStorage.init_with_gemm(g)
    assert sa.indptr.dtype == np.uint64 and sa.indices.dtype == np.uint64 and sa.data.dtype == np.float64
    assert sa.mat_shape == [1200, 400] and sb.mat_shape == [400, 1200]      # [cols, rows] (storage.rs:225)
    assert sa.row_num() == 400 and sa.get_ele_num(0, 2) == 764
    row = sa.read_row(0)
    assert str(row) == "rowptr: 0 indptr: [0, 1, 2, 3, 4] data: [0.052016, 0.038102, 0.024245, 0.01071, 0.003553]"
    assert str(spada.CsrRow.new_from_data(3, [], [])) == "rowptr: 3 indptr: [] data: []"
    remap = spada.sort_by_length(sa)
    assert sorted(remap.values()) == list(range(400))
    sq = spada.GEMM.from_mat("sq", random_csr(5, 5, density=0.5, seed=6))
    s1, s2 = spada.CsrMatStorage.init_with_gemm(sq)
    assert s2.indptr is s1.indptr                                           # shared buffers for A x A


def test_sort_by_length_is_stable(spada):
    lens = [3, 1, 3, 0, 1]
    m = sp.csr_matrix((np.ones(sum(lens)), np.concatenate([np.arange(l) for l in lens]),
                       np.concatenate([[0], np.cumsum(lens)])), shape=(5, 4))
    s, _ = spada.CsrMatStorage.init_with_gemm(spada.GEMM("x", m, m))
    assert spada.sort_by_length(s) == {0: 3, 1: 1, 2: 4, 3: 0, 4: 2}


def test_cli_grammar_and_config(spada, tmp_path):
    cli = spada.parse_args(["ACCURATESIMU", "spada", "ss", "cari", "cfg.json", "-p"])
    assert (cli.simulator, cli.accelerator, cli.category, cli.workload, cli.preprocess) == \
        ("AccurateSimu", "Spada", "SS", "cari", True)
    assert spada.parse_args(["accuratesimu", "multirow", "NN", "w", "c"]).accelerator == "MultiRow"
    with pytest.raises(SystemExit):
        spada.parse_args(["accuratesimu", "tpu", "ss", "cari", "cfg.json"])
    cfg = {"ss_filepath": "./matrices", "nn_filepath": "./matrices/nn_gemm.pkl", "pe_num": 2, "at_num": 16,
           "lane_num": 8, "cache_size": 1572864, "word_byte": 8, "block_shape": [1, 10000000], "mem_latency": 30,
           "cache_latency": 0, "freq": 1.0, "channel": 16, "bandwidth_per_channel": 8.0}
    fp = tmp_path / "config_1mb_row1.json"
    fp.write_text(json.dumps(cfg))
    buf = io.StringIO()
    with redirect_stdout(buf):
        c = spada.parse_config(str(fp))
    assert buf.getvalue() == f"{fp}\n"                   # frontend.rs:78 prints the path
    assert c.lane_num == 8 and c.block_shape == [1, 10000000] and c.ss_filepath == "./matrices"
    del cfg["lane_num"]
    fp.write_text(json.dumps(cfg))
    with pytest.raises(KeyError), redirect_stdout(io.StringIO()):
        spada.parse_config(str(fp))


def test_workload_generators_are_deterministic(spada):
    w = spada.workloads
    for name in ("poisson", "er", "rmat", "rect"):
        a1, b1 = w.build(name, 1 / 256)
        a2, _ = w.build(name, 1 / 256)
        assert a1.has_canonical_format and (a1 != a2).nnz == 0
        assert (b1 is a1) == (a1.shape[0] == a1.shape[1])
    # scaled-down known answers (generator regression guard); full-size counts are asserted in bench.py
    assert w.poisson2d(64).nnz == 5 * 64 * 64 - 4 * 64
    assert w.erdos_renyi(12).nnz == 65418 and w.rmat(12).nnz == 63906 and w.rect_powerlaw(12, 14).nnz == 118721
    assert w.algorithmic_bytes(20963328, 4194304, 20963328, 4194304, 54484996) == 1257_603_144


def test_balanced_bounds_rule():
    import importlib
    d = importlib.import_module("spada-sim_b200.distributed")
    w = np.array([5, 1, 1, 1, 10, 2, 2, 8, 1, 1])
    b = d.balanced_bounds(w, 4)
    assert b[0] == 0 and b[-1] == len(w) and np.all(np.diff(b) >= 0)
    assert d.balanced_bounds(np.ones(8, dtype=np.int64), 4).tolist() == [0, 2, 4, 6, 8]
    assert d.balanced_bounds(np.ones(3, dtype=np.int64), 8).tolist()[-1] == 3
