"""The wire / file formats either side of the path: the reference's YAML vector tree
(<function>/<suite>/<case>/data.yaml, reference src/lib.rs:14-204 and src/test_formats/*.rs) through
tools/run_yaml_vectors.py, and the reference's criterion benchmark names (benches/kzg_benches.rs:46-126) through
tools/criterion_bench.py.  The YAML tree is rebuilt from the committed fixtures (tests/golden), because the reference
checkout does not exist on the GPU box."""
import json
import os
import subprocess
import sys

import pytest
import yaml

from golden_util import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = golden()


def _hex(v):
    """fixture value -> the hex string the YAML file holds (malformed hex strings are kept as they are)"""
    if isinstance(v, dict):
        return "0x" + G.get_bytes(v).hex()
    return v


def _write_tree(root, cases):
    for c in cases:
        d = os.path.join(root, c["fn"], "small", c["name"])
        os.makedirs(d, exist_ok=True)
        inp = {k: ([_hex(x) for x in v] if isinstance(v, list) else _hex(v)) for k, v in c["input"].items()}
        with open(os.path.join(d, "data.yaml"), "w") as fh:
            yaml.safe_dump({"input": inp, "output": c["output"]}, fh)


def _run(tree, setup, *flags):
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_yaml_vectors.py"), tree, setup, *flags],
                          capture_output=True, text=True, timeout=1800)


def test_yaml_runner_on_the_oracle(tmp_path):
    """CPU: a slice of every suite (valid and invalid cases) through the runner with --oracle."""
    cases = []
    for fn in ("blob_to_kzg_commitment", "compute_kzg_proof", "compute_blob_kzg_proof", "verify_kzg_proof",
               "verify_blob_kzg_proof", "verify_blob_kzg_proof_batch"):
        cs = G.by_fn(fn)
        cases += [c for c in cs if c["output"] is None][:2] + [c for c in cs if c["output"] is not None][:2]
    tree, setup = str(tmp_path / "tests"), str(tmp_path / "trusted_setup.txt")
    _write_tree(tree, cases)
    G.write_setup_text(setup)
    out = _run(tree, setup, "--oracle")
    assert out.returncode == 0, out.stdout + out.stderr
    assert "total %d, mismatches 0" % len(cases) in out.stdout
    # a wrong expected output is reported
    bad = dict(cases[2], name="tampered", output="0x" + "a1" * 48) if cases[2]["fn"] == "blob_to_kzg_commitment" else None
    if bad:
        _write_tree(tree, [bad])
        out = _run(tree, setup, "--oracle")
        assert out.returncode == 1 and "MISMATCH" in out.stdout


@pytest.mark.gpu
def test_yaml_runner_all_vectors_on_the_gpu(tmp_path):
    """All 208 vectors as a YAML tree + trusted_setup.txt through load_trusted_setup_file and the C ABI."""
    tree, setup = str(tmp_path / "tests"), str(tmp_path / "trusted_setup.txt")
    _write_tree(tree, G.cases)
    G.write_setup_text(setup)
    out = _run(tree, setup)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "total %d, mismatches 0" % len(G.cases) in out.stdout


@pytest.mark.gpu
def test_criterion_benchmark_names():
    """tools/criterion_bench.py reports the reference's benchmark names (benches/kzg_benches.rs:46-126)."""
    env = dict(os.environ, KZG_CRITERION_COMB_WIDTH="8", KZG_B200_CHUNK="128")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "criterion_bench.py"), "3"], capture_output=True, text=True,
                         timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    names = [json.loads(line)["bench"] for line in out.stdout.splitlines() if line.startswith("{")]
    want = ["blob_to_kzg_commitment", "compute_kzg_proof", "compute_blob_kzg_proof", "verify_kzg_proof", "verify_blob_kzg_proof"] + \
           ["verify_blob_kzg_proof_batch/%d" % c for c in (1, 2, 4, 8, 16, 32, 64)]
    assert names[:len(want)] == want
