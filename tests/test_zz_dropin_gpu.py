"""GPU: the drop-in proper.  oracle/_ref/rawhash2_gpu is the reference's OWN rawhash2 — main.cpp option parsing,
slow5lib reader, kt_pipeline, PAF printer — with the one line `kt_for(p->n_threads, map_worker_for, in, s->n_sig)`
(src/rmap.cpp:700) replaced by rh_gpu_map_batch_raw through integration/rh_dropin.h.  Its output must be the PAF the
unmodified binary prints (the committed golden vectors; tests/test_cli.py shows the unmodified binary prints them)."""
import os
import subprocess

import pytest

import _bind
from golden_util import CASES, GoldenCase
from test_cli import CLI, DROPIN, ava_expected, ava_world, golden_cli_options

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.isfile(DROPIN), reason="oracle/_ref/rawhash2_gpu not built")]


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900)


@pytest.mark.parametrize("case", CASES)
def test_patched_reference_prints_the_golden_paf(built, tmp_path, case):
    from rawhash_b200 import api
    g = GoldenCase(case, str(tmp_path))
    opts = golden_cli_options(g)
    exp = _bind.strip_mt(g.paf)
    reads = str(tmp_path / "reads.blow5")
    api.write_slow5(reads, g.names, g.raws, *g.cal, float(g.sample_rate))
    # FASTA target: the reference builds its index in memory; the shim hands it over through ri_idx_dump
    r = run([DROPIN] + opts + ["-t", "4", "-p", g.model, g.fasta, reads])
    assert r.returncode == 0, r.stderr
    assert _bind.strip_mt(r.stdout) == exp
    # `.ind` target written by the reference itself (-d), tiny mini-batches: several step-1 calls through the pipeline
    ind = str(tmp_path / "t.ind")
    r = run([DROPIN] + opts + ["-t", "4", "-p", g.model, "-d", ind, g.fasta])
    assert r.returncode == 0, r.stderr
    r = run([DROPIN] + opts + ["-t", "4", "-K", "60k", ind, reads])
    assert r.returncode == 0, r.stderr
    assert _bind.strip_mt(r.stdout) == exp


def test_patched_reference_rawsamble(built, tmp_path):
    from rawhash_b200 import api, synth
    w = ava_world()
    reads = str(tmp_path / "reads.blow5")
    api.write_slow5(reads, w.names, w.reads["raw"], synth.OFFSET, synth.RANGE, synth.DIGITISATION)
    ind = str(tmp_path / "ava.ind")
    r = run([DROPIN, "-x", "ava", "-t", "4", "-p", w.model, "-d", ind, reads])  # the reference's own CPU index build
    assert r.returncode == 0, r.stderr
    r = run([DROPIN, "-x", "ava", "-t", "4", ind, reads])
    assert r.returncode == 0, r.stderr
    assert _bind.strip_mt(r.stdout).splitlines() == ava_expected(w)
