"""GPU: the `rawhash2_b200` command line end to end — FASTA/`.ind` + BLOW5/SLOW5 in, PAF out — against the committed
golden PAFs (which the unmodified reference binary prints for the same files, tests/test_cli.py) and, for Rawsamble,
against the oracle.  Sorted last on purpose: the stage-by-stage parity tests run first."""
import os
import subprocess

import pytest

import _bind
from golden_util import CASES, GoldenCase
from test_cli import CLI, ava_expected, ava_world, golden_cli_options

pytestmark = pytest.mark.gpu


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900)


@pytest.mark.parametrize("case", CASES)
def test_cli_maps_golden_reads(built, tmp_path, case):
    from rawhash_b200 import api
    g = GoldenCase(case, str(tmp_path))
    opts = golden_cli_options(g)
    exp = _bind.strip_mt(g.paf)
    blow5, slow5 = str(tmp_path / "reads.blow5"), str(tmp_path / "reads.slow5")
    api.write_slow5(blow5, g.names, g.raws, *g.cal, float(g.sample_rate))
    api.write_slow5(slow5, g.names, g.raws, *g.cal, float(g.sample_rate))
    # 1. FASTA + pore model (index built on the GPU), BLOW5 reads, index dumped on the way
    ind = str(tmp_path / "t.ind")
    r = run([CLI] + opts + ["-t", "4", "-p", g.model, "-d", ind, g.fasta, blow5])
    assert r.returncode == 0, r.stderr
    assert _bind.strip_mt(r.stdout) == exp
    assert "GPU index build not possible" not in r.stderr
    # 2. the dumped `.ind`, ASCII SLOW5 reads, tiny mini-batches (-K) so that the three pipeline steps overlap, -o FILE
    out = str(tmp_path / "out.paf")
    r = run([CLI] + opts + ["-t", "4", "-K", "60k", "-o", out, ind, slow5])
    assert r.returncode == 0 and r.stdout == "", r.stderr
    assert _bind.strip_mt(open(out).read()) == exp
    # 3. a directory of signal files is scanned like find_sfiles does; reads of both files are mapped in file order
    d = tmp_path / "dir"
    d.mkdir()
    half = len(g.names) // 2
    api.write_slow5(str(d / "a.blow5"), g.names[:half], g.raws[:half], *g.cal, float(g.sample_rate), 0, 1)
    api.write_slow5(str(d / "b.blow5"), g.names[half:], g.raws[half:], *g.cal, float(g.sample_rate), 1, 0)
    r = run([CLI] + opts + [ind, str(d)])
    assert r.returncode == 0, r.stderr
    assert _bind.strip_mt(r.stdout) == exp


def test_cli_rawsamble(built, tmp_path):
    """-x ava: `-d` builds the index from the reads' own signals (event detection on the GPU), then all-vs-all."""
    from rawhash_b200 import api, synth
    w = ava_world()
    reads = str(tmp_path / "reads.blow5")
    api.write_slow5(reads, w.names, w.reads["raw"], synth.OFFSET, synth.RANGE, synth.DIGITISATION)
    ind = str(tmp_path / "ava.ind")
    r = run([CLI, "-x", "ava", "-p", w.model, "-d", ind, reads])
    assert r.returncode == 0, r.stderr
    r = run([CLI, "-x", "ava", ind, reads])
    assert r.returncode == 0, r.stderr
    exp = ava_expected(w)
    assert len(exp) > len(w.names)
    assert _bind.strip_mt(r.stdout).splitlines() == exp


def _n_gpus():
    import shutil
    if shutil.which("nvidia-smi") is None:
        return 0
    r = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True)
    return sum(1 for ln in r.stdout.splitlines() if ln.startswith("GPU ")) if r.returncode == 0 else 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_cli_two_gpus_same_paf(built, tmp_path):
    """--gpus 2: every mini-batch is cut into two sample-balanced read ranges mapped on their own GPU (index
    replicated), records concatenated in range order — the PAF must not change."""
    from rawhash_b200 import api, synth
    from common import World
    g = GoldenCase("r94_sensitive", str(tmp_path))
    reads = str(tmp_path / "reads.blow5")
    api.write_slow5(reads, g.names, g.raws, *g.cal)
    r = run([CLI, "-x", "sensitive", "--gpus", "2", "-p", g.model, g.fasta, reads])
    assert r.returncode == 0, r.stderr
    assert _bind.strip_mt(r.stdout) == _bind.strip_mt(g.paf)
    w = World(n_contigs=3, genome_len=900_000, n_reads=600, read_bp=3000, seed=31)
    big = str(tmp_path / "big.blow5")
    api.write_slow5(big, w.names, w.reads["raw"], synth.OFFSET, synth.RANGE, synth.DIGITISATION)
    ind = str(tmp_path / "w.ind")
    assert run([CLI, "-x", "sensitive", "-p", w.model, "-d", ind, w.fasta]).returncode == 0
    one = run([CLI, "-x", "sensitive", "-K", "8M", ind, big])
    two = run([CLI, "-x", "sensitive", "-K", "8M", "--gpus", "2", ind, big])
    assert one.returncode == 0 and two.returncode == 0, one.stderr + two.stderr
    assert len(one.stdout.splitlines()) >= 600 and _bind.strip_mt(one.stdout) == _bind.strip_mt(two.stdout)
