"""CLI drop-in: `spada-sim <simulator> <accelerator> <category> <workload> <configuration> [-p]`
(frontend.rs:52-75, main.rs:30-121) reproduced by `python -m spada-sim_b200`."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import scipy.io

from conftest import GOLDEN, ROOT

CFG = {"ss_filepath": "./matrices", "nn_filepath": "./matrices/nn_gemm.pkl", "pe_num": 2, "at_num": 16, "lane_num": 8,
       "cache_size": 1572864, "word_byte": 8, "block_shape": [1, 10000000], "mem_latency": 30, "cache_latency": 0,
       "freq": 1.0, "channel": 16, "bandwidth_per_channel": 8.0}   # = the reference's config/config_1mb_row1.json


@pytest.fixture()
def workdir(tmp_path, cari):
    (tmp_path / "matrices").mkdir()
    scipy.io.mmwrite(str(tmp_path / "matrices" / "cari.mtx"), cari, precision=17)   # round-trips every f64
    (tmp_path / "config").mkdir()
    (tmp_path / "config" / "config_1mb_row1.json").write_text(json.dumps(CFG))
    return tmp_path


def run_cli(workdir, *args):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, "-W", "ignore", "-m", "spada-sim_b200", *args], cwd=workdir, env=env,
                          capture_output=True, text=True, timeout=600)


HEAD = """config/config_1mb_row1.json
---- Python Interface ----
% Load cari from ./matrices
Get GEMM cari
---- cari ----
--A: (400, 1200)
data: [0.052016, 0.038102, 0.024245, 0.01071, 0.003553] .. 
indices: [0, 1, 2, 3, 4] ...
indptr: [0, 382, 764, 1146, 1528] ...
--B: (1200, 400)
data: [0.052016, 0.038102, 0.024245, 0.01071, 0.003553] ...
indices: [0, 1, 2, 3, 4] ...
indptr: [0, 380, 760, 1140, 1520] ...

Avg row len of A: 382, Avg row len of B: 127
"""


def test_cli_loads_and_fails_loudly_without_gpu(workdir, spada):
    if spada.device_count() > 0:
        pytest.skip("a GPU is present")
    p = run_cli(workdir, "accuratesimu", "spada", "ss", "cari", "config/config_1mb_row1.json")
    assert p.returncode != 0
    assert p.stdout == HEAD                      # everything up to the engine matches the reference transcript
    assert "NO_DEVICE" in p.stderr and "no CPU fallback" in p.stderr


def test_cli_rejects_bad_enum(workdir):
    p = run_cli(workdir, "accuratesimu", "tpu", "ss", "cari", "config/config_1mb_row1.json")
    assert p.returncode == 2 and "isn't a valid value" in p.stderr


def test_rust_binding_lists_every_abi_symbol():
    hdr = open(os.path.join(ROOT, "include", "spada_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(spada_b200_[a-z0-9_]+)\s*\(", hdr))
    rs = open(os.path.join(ROOT, "spada-sim_b200", "rust", "spada-b200-sys", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (spada_b200_[a-z0-9_]+)\s*\(", rs))
    assert bound == declared


@pytest.mark.gpu
@pytest.mark.parametrize("acc,extra", [("spada", []), ("spada", ["-p"]), ("ip", []), ("op", []), ("MultiRow", [])])
def test_cli_cari_transcript(workdir, oracle, cari, spada, acc, extra):
    # the accelerator picks the window policy (rows per tile), never the result
    p = run_cli(workdir, "accuratesimu", acc, "ss", "cari", "config/config_1mb_row1.json", *extra)
    assert p.returncode == 0, p.stderr
    assert p.stdout.startswith(HEAD)
    tail = p.stdout[len(HEAD):].splitlines()
    assert tail[0] == "-----Result-----" and tail[1] == "-----Access count"
    assert tail[2] == "Execution count: 0"
    assert tail[3] == "A matrix count: read 305600 write 0"                 # 2 x nnz(A)
    assert tail[4] == "B matrix count: read 115521600 write 0"             # 2 x 57 760 800 products
    assert tail[5] == "C matrix count: read 0 write 320400"                # 2 x 160 000 + 400 rows
    assert tail[6] == "Cache count: read 0 write 0"
    assert tail[7] == "-----Output product matrix"
    rows = tail[8:]
    assert len(rows) == 10                                                  # min(result.len(), 10), main.rs:114
    g = spada.GEMM.from_mat("cari", cari)
    cp, cj, cx = oracle.spgemm(g.a, g.b)
    for r, line in enumerate(rows):
        m = re.fullmatch(r"rowptr: (\d+) indptr: \[(.*)\] data: \[(.*)\]", line)
        assert m and int(m.group(1)) == r
        assert [int(x) for x in m.group(2).split(", ")] == cj[cp[r]:cp[r] + 5].tolist()
        vals = np.array([float(x) for x in m.group(3).split(", ")])
        assert np.array_equal(vals, cx[cp[r]:cp[r] + 5])   # bit-exact; -p never changes C (simulator.rs:1039-1060)


@pytest.mark.gpu
def test_cli_on_all_gpus_of_the_process(workdir, oracle, cari, spada):
    # SPADA_B200_GPUS=2: the same transcript with every product sharded over two GPUs (spada_b200_group_*), both hosts
    if spada.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    g = spada.GEMM.from_mat("cari", cari)
    cp, cj, cx = oracle.spgemm(g.a, g.b)
    env = dict(os.environ, PYTHONPATH=ROOT, SPADA_B200_GPUS="2", SPADA_B200_DUMP_C=str(workdir / "c2.mtx"))
    cmds = [[sys.executable, "-W", "ignore", "-m", "spada-sim_b200"]]
    binary = os.path.join(ROOT, "spada-sim_b200", "bin", "spada-sim")
    if os.path.exists(binary):
        cmds.append([binary])
    for cmd in cmds:
        p = subprocess.run(cmd + ["accuratesimu", "spada", "ss", "cari", "config/config_1mb_row1.json"], cwd=workdir, env=env,
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr
        assert p.stdout.startswith(HEAD)
        c = scipy.io.mmread(str(workdir / "c2.mtx")).tocsr(); c.sort_indices()
        assert np.array_equal(c.indptr, cp) and np.array_equal(c.indices, cj)
        assert np.array_equal(c.data.view(np.uint64), cx.view(np.uint64))


def test_result_sink_roundtrip(tmp_path, spada):
    # SPADA_B200_DUMP_C: Matrix Market text that round-trips every f64, FNV-1a-64 digest
    main = __import__("importlib").import_module("spada-sim_b200.main")
    assert main.fnv1a64(b"") == 0xcbf29ce484222325 and main.fnv1a64(b"a") == 0xaf63dc4c8601ec8c   # published vectors
    rng = np.random.default_rng(3)
    import scipy.sparse as sp
    c = sp.random(37, 23, density=0.2, format="csr", random_state=rng); c.sort_indices()
    c.data = rng.uniform(-1, 1, c.nnz) * 10.0 ** rng.integers(-30, 30, c.nnz)
    line = main.dump_result(str(tmp_path / "c.mtx"), c.indptr, c.indices, c.data, c.shape[1])
    assert line.startswith(f"C dumped: nnz {c.nnz} sum ")
    back = scipy.io.mmread(str(tmp_path / "c.mtx")).tocsr(); back.sort_indices()
    assert back.shape == c.shape and np.array_equal(back.indptr, c.indptr) and np.array_equal(back.indices, c.indices)
    assert np.array_equal(back.data.view(np.uint64), c.data.view(np.uint64))


@pytest.mark.gpu
def test_cli_result_sink(workdir, oracle, cari, spada):
    env = dict(os.environ, PYTHONPATH=ROOT, SPADA_B200_DUMP_C=str(workdir / "c_py.mtx"))
    p = subprocess.run([sys.executable, "-W", "ignore", "-m", "spada-sim_b200", "accuratesimu", "spada", "ss", "cari",
                        "config/config_1mb_row1.json"], cwd=workdir, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    assert p.stdout.startswith(HEAD)                      # stdout stays the reference's transcript
    assert re.search(r"C dumped: nnz 160000 sum 7833\.70723\d+ fnv1a64 [0-9a-f]{16}", p.stderr)
    g = spada.GEMM.from_mat("cari", cari)
    cp, cj, cx = oracle.spgemm(g.a, g.b)
    outs = [workdir / "c_py.mtx"]
    binary = os.path.join(ROOT, "spada-sim_b200", "bin", "spada-sim")
    if os.path.exists(binary):
        env["SPADA_B200_DUMP_C"] = str(workdir / "c_cpp.mtx")
        q = subprocess.run([binary, "accuratesimu", "spada", "ss", "cari", "config/config_1mb_row1.json"], cwd=workdir,
                           env=env, capture_output=True, text=True, timeout=600)
        assert q.returncode == 0, q.stderr
        assert re.search(r"C dumped: nnz 160000 sum 7833\.70723\d+ fnv1a64 [0-9a-f]{16}", q.stderr)
        outs.append(workdir / "c_cpp.mtx")
    for path in outs:
        c = scipy.io.mmread(str(path)).tocsr(); c.sort_indices()
        assert c.shape == (400, 400) and np.array_equal(c.indptr, cp) and np.array_equal(c.indices, cj)
        assert np.array_equal(c.data.view(np.uint64), cx.view(np.uint64))
