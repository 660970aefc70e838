"""Compiled host side (spada-sim_b200/host: C++ mirror of the reference interface + `spada-sim` CLI).

CPU: argv grammar, config parsing, native Matrix Market reader (bit-for-bit against
scipy.io.mmread(..).tocsr(), the reference's loader py2rust.rs:63-80), GEMM::from_mat transpose,
.pkl path, stdout transcript up to the engine, loud failure without a GPU.
GPU: the full transcript for cari."""
import json
import os
import pickle
import re
import subprocess

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

from conftest import ROOT, random_csr
from test_cli import CFG, HEAD

HOST = os.path.join(ROOT, "spada-sim_b200", "host")
BIN = os.path.join(ROOT, "spada-sim_b200", "bin", "spada-sim")


@pytest.fixture(scope="session")
def binary(spada):
    spada._abi.lib()   # the shared library must exist first
    subprocess.check_call(["make", "-C", HOST, "-s"])
    return BIN


def write_case(tmp_path, name, mat, **mm):
    (tmp_path / "matrices").mkdir(exist_ok=True)
    scipy.io.mmwrite(str(tmp_path / "matrices" / f"{name}.mtx"), mat, precision=17, **mm)
    (tmp_path / "cfg.json").write_text(json.dumps(CFG))


def run(binary, cwd, *args, dump=None):
    env = dict(os.environ)
    if dump:
        env["SPADA_B200_DUMP_OPERANDS"] = str(dump)
    return subprocess.run([binary, *args], cwd=cwd, env=env, capture_output=True, text=True, timeout=600)


def read_dump(path):
    raw = np.fromfile(path, dtype="<u8")
    out, o = [], 0
    for _ in range(2):
        rows, cols, nnz = (int(x) for x in raw[o:o + 3]); o += 3
        ip = raw[o:o + rows + 1].astype(np.int64); o += rows + 1
        ix = raw[o:o + nnz].astype(np.int64); o += nnz
        dx = raw[o:o + nnz].view("<f8"); o += nnz
        out.append(sp.csr_matrix((dx, ix, ip), shape=(rows, cols)))
    return out


def same_bits(x, ref):
    ref = ref.tocsr(); ref.sort_indices()
    return (x.shape == ref.shape and np.array_equal(x.indptr, ref.indptr) and np.array_equal(x.indices, ref.indices)
            and np.array_equal(x.data.view(np.uint64), ref.data.astype(np.float64).view(np.uint64)))


def test_transcript_and_loud_failure_without_gpu(binary, tmp_path, cari, spada):
    if spada.device_count() > 0:
        pytest.skip("a GPU is present")
    (tmp_path / "config").mkdir()
    (tmp_path / "config" / "config_1mb_row1.json").write_text(json.dumps(CFG))
    (tmp_path / "matrices").mkdir()
    scipy.io.mmwrite(str(tmp_path / "matrices" / "cari.mtx"), cari, precision=17)
    p = run(binary, tmp_path, "accuratesimu", "spada", "ss", "cari", "config/config_1mb_row1.json")
    assert p.returncode == 101                      # Rust's panic status
    assert p.stdout == HEAD                          # identical to the Python mirror and the reference sections
    assert "no CPU fallback" in p.stderr and "spada_b200 error 7" in p.stderr


@pytest.mark.parametrize("field,symmetry", [("real", "general"), ("real", "symmetric"), ("integer", "general"),
                                            ("pattern", "general"), ("real", "skew-symmetric")])
def test_native_mtx_reader_matches_scipy_bits(binary, tmp_path, field, symmetry):
    m = random_csr(37, 37, density=0.15, seed=5, values="int" if field == "integer" else "signed")
    if symmetry == "symmetric":
        m = (m + m.T).tocsr()
    if symmetry == "skew-symmetric":
        m = (m - m.T).tocsr()
    if field == "pattern":
        m.data[:] = 1.0
    write_case(tmp_path, "w", m.astype(np.int64) if field == "integer" else m, field=field, symmetry=symmetry)
    run(binary, tmp_path, "accuratesimu", "spada", "ss", "w", "cfg.json", dump=tmp_path / "d.bin")
    with open(tmp_path / "matrices" / "w.mtx") as f:
        ref = scipy.io.mmread(f).tocsr().astype(np.float64)
    a, b = read_dump(tmp_path / "d.bin")
    assert same_bits(a, ref) and same_bits(b, ref)    # square => B = A (gemm.rs:42-43)


def test_rectangular_transpose_and_cari_bits(binary, tmp_path, cari):
    write_case(tmp_path, "cari", cari)
    p = run(binary, tmp_path, "accuratesimu", "spada", "ss", "cari", "cfg.json", dump=tmp_path / "d.bin")
    assert p.stdout.startswith("cfg.json\n---- Python Interface ----\n% Load cari from ./matrices\nGet GEMM cari\n")
    a, b = read_dump(tmp_path / "d.bin")
    assert same_bits(a, cari) and same_bits(b, cari.T)  # 400 x 1200 is not square => B = A^T (gemm.rs:44-46)


def test_pickled_gemm_path(binary, tmp_path):
    a = random_csr(9, 7, density=0.4, seed=6)
    b = random_csr(7, 11, density=0.4, seed=7)
    cfg = dict(CFG, nn_filepath=str(tmp_path / "nn_gemm.pkl"))
    (tmp_path / "cfg.json").write_text(json.dumps(cfg))
    pickle.dump({"g": (a.tocsc(), b.toarray()), "bad": (a.tolil(), b)}, open(tmp_path / "nn_gemm.pkl", "wb"))
    p = run(binary, tmp_path, "accuratesimu", "spada", "nn", "g", "cfg.json", dump=tmp_path / "d.bin")
    out = p.stdout.splitlines()
    assert out[1] == "---- Python Interface ----" and out[2].startswith("% Load g from")
    assert "--- Return from Python Interface ---" in out and "Get GEMM g" in out
    x, y = read_dump(tmp_path / "d.bin")
    assert same_bits(x, a) and same_bits(y, b)
    p = run(binary, tmp_path, "accuratesimu", "spada", "nn", "bad", "cfg.json")
    assert p.returncode == 101 and "Unsupported matrix type" in p.stderr


def test_errors_like_the_reference(binary, tmp_path):
    (tmp_path / "matrices").mkdir()
    scipy.io.mmwrite(str(tmp_path / "matrices" / "dense.mtx"), np.eye(3))
    (tmp_path / "cfg.json").write_text(json.dumps(CFG))
    p = run(binary, tmp_path, "accuratesimu", "spada", "ss", "dense", "cfg.json")
    assert p.returncode == 101 and "has no attribute 'tocsr'" in p.stderr       # array format: mmread -> ndarray
    p = run(binary, tmp_path, "accuratesimu", "tpu", "ss", "dense", "cfg.json")
    assert p.returncode == 2 and "isn't a valid value" in p.stderr
    p = run(binary, tmp_path, "accuratesimu", "spada", "ss")
    assert p.returncode == 2
    bad = dict(CFG); del bad["lane_num"]
    (tmp_path / "bad.json").write_text(json.dumps(bad))
    p = run(binary, tmp_path, "accuratesimu", "spada", "ss", "dense", "bad.json")
    assert p.returncode == 101 and "missing field `lane_num`" in p.stderr
    p = run(binary, tmp_path, "trafficmodel", "spada", "ss", "nothere", "cfg.json")
    assert p.returncode == 101 and "No such file" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("acc,extra", [("spada", []), ("ip", ["-p"])])
def test_cpp_cli_cari_transcript(binary, tmp_path, oracle, cari, spada, acc, extra):
    (tmp_path / "config").mkdir()
    (tmp_path / "config" / "config_1mb_row1.json").write_text(json.dumps(CFG))
    (tmp_path / "matrices").mkdir()
    scipy.io.mmwrite(str(tmp_path / "matrices" / "cari.mtx"), cari, precision=17)
    p = run(binary, tmp_path, "accuratesimu", acc, "ss", "cari", "config/config_1mb_row1.json", *extra)
    assert p.returncode == 0, p.stderr
    assert p.stdout.startswith(HEAD)
    tail = p.stdout[len(HEAD):].splitlines()
    assert tail[:8] == ["-----Result-----", "-----Access count", "Execution count: 0",
                        "A matrix count: read 305600 write 0", "B matrix count: read 115521600 write 0",
                        "C matrix count: read 0 write 320400", "Cache count: read 0 write 0",
                        "-----Output product matrix"]
    rows = tail[8:]
    assert len(rows) == 10
    g = spada.GEMM.from_mat("cari", cari)
    cp, cj, cx = oracle.spgemm(g.a, g.b)
    for r, line in enumerate(rows):
        m = re.fullmatch(r"rowptr: (\d+) indptr: \[(.*)\] data: \[(.*)\]", line)
        assert m and int(m.group(1)) == r
        assert [int(x) for x in m.group(2).split(", ")] == cj[cp[r]:cp[r] + 5].tolist()
        vals = np.array([float(x) for x in m.group(3).split(", ")])
        assert np.allclose(vals, cx[cp[r]:cp[r] + 5], rtol=1e-12, atol=0)


SINK_DRIVER = r"""
#include "spada_host.hpp"
#include <cstdio>
#include <cstdlib>
using namespace spada_host;
// argv: out.mtx n_cols  ; stdin: rows, then per row: n, then n x (col value-bits-as-u64)
int main(int argc, char** argv) {
    size_t n_rows;
    if (std::scanf("%zu", &n_rows) != 1) return 3;
    std::vector<CsrRow> rows(n_rows);
    for (size_t r = 0; r < n_rows; ++r) {
        size_t n;
        if (std::scanf("%zu", &n) != 1) return 3;
        rows[r].rowptr = r;
        for (size_t j = 0; j < n; ++j) {
            unsigned long long c, bits;
            if (std::scanf("%llu %llu", &c, &bits) != 2) return 3;
            double v;
            std::memcpy(&v, &bits, 8);
            rows[r].indptr.push_back(c);
            rows[r].data.push_back(v);
        }
    }
    std::printf("%s\n", dump_result(argv[1], rows, std::strtoull(argv[2], nullptr, 10)).c_str());
    return 0;
}
"""


def test_cpp_result_sink_matches_python(tmp_path, spada):
    # SPADA_B200_DUMP_C in the compiled host: same file (every f64 round-trips) and same digest line as main.py
    main = __import__("importlib").import_module("spada-sim_b200.main")
    (tmp_path / "drv.cpp").write_text(SINK_DRIVER)
    exe = tmp_path / "drv"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", HOST, "-I", os.path.join(ROOT, "include"), "-o", str(exe),
                           str(tmp_path / "drv.cpp"), "-L", os.path.join(ROOT, "spada-sim_b200", "lib"), "-lspada_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "spada-sim_b200", "lib")])
    rng = np.random.default_rng(5)
    c = sp.random(41, 29, density=0.25, format="csr", random_state=rng); c.sort_indices()
    c.data = rng.uniform(-1, 1, c.nnz) * 10.0 ** rng.integers(-200, 200, c.nnz)
    text = [str(c.shape[0])]
    for r in range(c.shape[0]):
        s0, s1 = c.indptr[r], c.indptr[r + 1]
        text.append(str(s1 - s0))
        text += [f"{int(j)} {int(b)}" for j, b in zip(c.indices[s0:s1], c.data[s0:s1].view(np.uint64))]
    p = subprocess.run([str(exe), str(tmp_path / "c_cpp.mtx"), str(c.shape[1])], input="\n".join(text), text=True,
                       capture_output=True, timeout=60)
    assert p.returncode == 0, p.stderr
    line_py = main.dump_result(str(tmp_path / "c_py.mtx"), c.indptr, c.indices, c.data, c.shape[1])
    assert p.stdout.strip() == line_py
    a = scipy.io.mmread(str(tmp_path / "c_cpp.mtx")).tocsr(); a.sort_indices()
    assert a.shape == c.shape and np.array_equal(a.indptr, c.indptr) and np.array_equal(a.indices, c.indices)
    assert np.array_equal(a.data.view(np.uint64), c.data.view(np.uint64))
