"""CPU: the `rawhash2_b200` command line (csrc/rh_main.cpp) next to the unmodified reference binary
(oracle/_ref/rawhash2 = src/main.cpp + slow5lib, built by oracle/Makefile): option handling, `-d` index files,
and the end-to-end file path that pins the GPU CLI test (tests/test_zz_cli_gpu.py): the reference binary, fed the
FASTA / BLOW5 / `.ind` files OUR writers produce, prints the committed golden PAF."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import _bind
from common import World
from golden_util import CASES, GoldenCase

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "rawhash_b200", "rawhash2_b200")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "rawhash2")
needs_ref_cli = pytest.mark.skipif(not os.path.isfile(REF_CLI), reason="oracle/_ref/rawhash2 not built")


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, **kw)


def have_gpu():
    return shutil.which("nvidia-smi") is not None and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0


def golden_cli_options(g):
    opts = ["-x", g.preset] + (["--r10"] if g.r10 else [])
    if g.non_default_sampling():
        opts += ["--sample-rate", str(g.sample_rate), "--bp-per-sec", str(g.bp_per_sec)]
    return opts


def test_cli_is_built_and_reports_usage(built):
    assert os.path.isfile(CLI), "rawhash_b200/rawhash2_b200 missing: python -m rawhash_b200.build"
    r = run([CLI])
    assert r.returncode == 1 and "Usage: rawhash2_b200" in r.stderr
    r = run([CLI, "-h"])
    assert r.returncode == 0 and "--max-chunks" in r.stdout and "[10]" in r.stdout
    assert run([CLI, "--version"]).stdout.startswith("2.1")
    r = run([CLI, "-x", "viral", "-h"])  # the preset is applied before the help text prints its defaults
    assert "--max-chunks INT [5]" in r.stdout


@pytest.mark.parametrize("args,msg", [
    (["-x", "nope", "t.fa"], "unknown preset 'nope'"),
    (["--no-such-option", "t.fa"], "unknown option"),
    (["-t"], "missing option argument"),
    (["--rmq", "t.fa"], "outside the mapping path"),
    (["--dtw-evaluate-chains", "t.fa"], "outside the mapping path"),
    (["--sequence-until", "t.fa"], "outside the mapping path"),
    (["-t", "1", "--io-thread", "2", "t.fa"], "must NOT be smaller"),
    (["-w", "3", "-n", "2", "t.fa"], "cannot be set together"),
    (["/no/such/file.fa", "q.blow5"], "failed to open file"),
])
def test_cli_option_errors(built, args, msg):
    r = run([CLI] + args)
    assert r.returncode == 1 and msg in r.stderr, r.stderr


def test_cli_missing_inputs(built, tmp_path):
    w = World(n_contigs=1, genome_len=20_000, n_reads=1, read_bp=500, seed=3)
    r = run([CLI, w.fasta])
    assert r.returncode == 1 and "missing input: please specify a query" in r.stderr
    r = run([CLI, "-d", str(tmp_path / "x.ind"), w.fasta])
    assert r.returncode == 1 and "pore model file with -p" in r.stderr


@needs_ref_cli
@pytest.mark.parametrize("opts", [["-x", "sensitive"], ["-x", "faster"], ["-x", "fast", "-e", "7", "--sig-diff", "0.3", "--fine-range", "0.5"]])
def test_cli_index_dump_equals_reference_binary(built, tmp_path, opts):
    """`rawhash2_b200 -d` (host builder here, GPU builder on a GPU box) vs `rawhash2 -d`: same bytes but the two
    stale pointers; options given AFTER the positional argument are honoured like ketopt's permutation does."""
    w = World(n_contigs=3, genome_len=240_000, n_reads=1, read_bp=500, seed=17)
    mine, theirs = str(tmp_path / "mine.ind"), str(tmp_path / "ref.ind")
    r = run([REF_CLI] + opts + ["-t", "4", "-p", w.model, "-d", theirs, w.fasta])
    assert r.returncode == 0, r.stderr
    r = run([CLI] + opts + ["-t", "4", "-p", w.model, w.fasta, "-d", mine])
    assert r.returncode == 0 and "Only the index is constructed" in r.stderr, r.stderr
    a, b = np.fromfile(mine, np.uint8), np.fromfile(theirs, np.uint8)
    assert a.size == b.size
    d = np.nonzero(a != b)[0]
    assert d.size <= 16 and (d.size == 0 or (d.min() >= 46 and d.max() < 62))


@needs_ref_cli
@pytest.mark.parametrize("case", CASES)
def test_reference_binary_on_our_files_prints_the_golden_paf(built, tmp_path, case):
    """The chain of custody for the GPU CLI test: golden reads written by rh_slow5_write, the index written by
    rh_index_dump, both consumed by the UNMODIFIED reference binary -> the committed golden PAF."""
    from rawhash_b200 import api
    g = GoldenCase(case, str(tmp_path))
    reads = str(tmp_path / "reads.blow5")
    api.write_slow5(reads, g.names, g.raws, *g.cal, float(g.sample_rate))
    opts = golden_cli_options(g)
    r = run([REF_CLI] + opts + ["-t", "2", "-p", g.model, g.fasta, reads])
    assert r.returncode == 0, r.stderr
    assert _bind.strip_mt(r.stdout) == _bind.strip_mt(g.paf)
    ind = str(tmp_path / "ours.ind")
    r = run([CLI] + opts + ["-t", "2", "-p", g.model, "-d", ind, g.fasta])
    assert r.returncode == 0, r.stderr
    r = run([REF_CLI] + opts + ["-t", "2", ind, reads])
    assert r.returncode == 0, r.stderr
    assert _bind.strip_mt(r.stdout) == _bind.strip_mt(g.paf)


@needs_ref_cli
def test_reference_binary_rawsamble_equals_oracle(built, tmp_path):
    """-x ava through files: the reference binary (index from the reads' own BLOW5, then all-vs-all) prints what the
    oracle's in-memory path prints — the expected output of the GPU CLI's Rawsamble test."""
    from rawhash_b200 import api, synth
    w = ava_world()
    reads = str(tmp_path / "reads.blow5")
    api.write_slow5(reads, w.names, w.reads["raw"], synth.OFFSET, synth.RANGE, synth.DIGITISATION)
    ind = str(tmp_path / "ava.ind")
    r = run([REF_CLI, "-x", "ava", "-t", "2", "-p", w.model, "-d", ind, reads])  # ri_idx_siggen refuses to run without a pore model
    assert r.returncode == 0, r.stderr
    r = run([REF_CLI, "-x", "ava", "-t", "2", ind, reads])
    assert r.returncode == 0, r.stderr
    exp = ava_expected(w)
    assert len(exp) > len(w.names)
    assert _bind.strip_mt(r.stdout).splitlines() == exp


def ava_world():
    return World(n_contigs=1, genome_len=60_000, n_reads=60, read_bp=4000, seed=23)


def ava_expected(w):
    orc = _bind.OracleLib().open("ava", False, w.model)
    sigs = [w.pa(i) for i in range(len(w.names))]
    orc.build_index_sig(sigs, w.names, 4)
    orc.mapopt_update()
    exp, _ = orc.map_paf(sigs, w.names, 4)
    return _bind.strip_mt(exp).splitlines()


@pytest.mark.skipif(have_gpu(), reason="a GPU is present: the CLI maps instead of refusing")
def test_cli_refuses_to_map_without_a_gpu(built, tmp_path):
    from rawhash_b200 import api
    g = GoldenCase("r94_sensitive", str(tmp_path))
    reads = str(tmp_path / "reads.blow5")
    api.write_slow5(reads, g.names, g.raws, *g.cal)
    r = run([CLI, "-x", "sensitive", "-p", g.model, g.fasta, reads])
    assert r.returncode == 1 and "there is no CPU mapping path" in r.stderr and r.stdout == ""


# ---- the drop-in itself: the reference's own binary with step 1 replaced (integration/build_dropin.py) ----------------
DROPIN = os.path.join(ROOT, "oracle", "_ref", "rawhash2_gpu")
needs_dropin = pytest.mark.skipif(not os.path.isfile(DROPIN), reason="oracle/_ref/rawhash2_gpu not built (integration/build_dropin.py)")


@needs_dropin
@pytest.mark.skipif(have_gpu(), reason="a GPU is present: the patched reference maps instead of refusing")
def test_patched_reference_fails_loudly_without_a_gpu(built, tmp_path):
    """The reference binary with `kt_for(map_worker_for)` replaced by rh_gpu_map_batch_raw has no CPU path left."""
    from rawhash_b200 import api
    g = GoldenCase("r94_sensitive", str(tmp_path))
    reads = str(tmp_path / "reads.blow5")
    api.write_slow5(reads, g.names, g.raws, *g.cal)
    r = run([DROPIN, "-x", "sensitive", "-t", "2", "-p", g.model, g.fasta, reads])
    assert r.returncode == 1 and "no usable CUDA device" in r.stderr and r.stdout == ""


@needs_dropin
@needs_ref_cli
def test_patched_reference_index_only_mode_is_untouched(built, tmp_path):
    w = World(n_contigs=2, genome_len=100_000, n_reads=1, read_bp=500, seed=21)
    a, b = str(tmp_path / "a.ind"), str(tmp_path / "b.ind")
    assert run([DROPIN, "-x", "sensitive", "-t", "2", "-p", w.model, "-d", a, w.fasta]).returncode == 0
    assert run([REF_CLI, "-x", "sensitive", "-t", "2", "-p", w.model, "-d", b, w.fasta]).returncode == 0
    x, y = np.fromfile(a, np.uint8), np.fromfile(b, np.uint8)
    d = np.nonzero(x != y)[0]
    assert x.size == y.size and d.size <= 16


# ---- option parsing next to the reference's: both programs print the values in effect in their help text -------------
_NUM = r"\[(-?[0-9][0-9.e+-]*(?:, ?-?[0-9][0-9.e+-]*)?)\]"
_SHOWN = ["-k", "-e", "-q", "-w", "-t", "--level_column", "--sig-diff", "--q-mid-occ", "--min-events", "--bw", "--max-target-gap", "--max-query-gap",
          "--min-anchors", "--best-chains", "--min-score", "--chain-gap-scale", "--chain-skip-scale", "--primary-ratio", "--primary-length",
          "--max-skips", "--max-iterations", "--max-chunks", "--min-mapq", "--bp-per-sec", "--sample-rate", "--chunk-size",
          "--seg-window-length1", "--seg-window-length2", "--seg-threshold1", "--seg-threshold2", "--seg-peak-height", "--io-thread"]


def _values_in_help(text):
    import re
    out = {}
    for name in _SHOWN:
        m = re.search(r"(?m)^\s+" + re.escape(name) + r" [A-Z][^\n]*", text) if not name.startswith("--") else re.search(re.escape(name) + r" [A-Z][^\n]*", text)
        assert m, name
        seg = m.group(0)
        nxt = re.search(r"\s--[a-z]", seg[len(name):])  # our help lists several options on one line
        if nxt:
            seg = seg[:len(name) + nxt.start()]
        v = re.search(_NUM, seg)
        assert v, (name, seg)
        out[name] = [float(x) for x in v.group(1).replace(" ", "").split(",")]
    return out


@needs_ref_cli
@pytest.mark.parametrize("args", [
    [],
    ["-x", "viral"], ["-x", "sensitive"], ["-x", "fast"], ["-x", "faster"], ["-x", "ava"], ["-x", "ava-sensitive"], ["-x", "ava-viral"], ["-x", "ava-large"],
    ["--r10"], ["--depletion"], ["-x", "fast", "--r10", "--depletion"],
    ["--bw", "77", "-x", "viral"],  # the preset is applied first wherever it stands
    ["-k", "9", "-e", "7", "-q", "3", "-w", "5", "-t", "12", "--io-thread", "4", "--level_column", "2", "--sig-diff", "0.3"],
    ["--q-mid-occ", "40,9000", "--min-events", "33", "--max-target-gap", "1234", "--max-query-gap", "4321", "--min-anchors", "4", "--best-chains", "3"],
    ["--q-mid-occ", "60", "--min-score", "21", "--chain-gap-scale", "0.75", "--chain-skip-scale", "0.125", "--primary-ratio", "0.25", "--primary-length", "5000"],
    ["--max-skips", "9", "--max-iterations", "150", "--max-chunks", "3", "--min-mapq", "7", "--chunk-size", "5000"],
    ["--bp-per-sec", "400", "--sample-rate", "5000"], ["--sample-rate", "5000", "--bp-per-sec", "400"],
    ["--seg-window-length1", "4", "--seg-window-length2", "8", "--seg-threshold1", "5.5", "--seg-threshold2", "2.25", "--seg-peak-height", "0.3"],
    ["-k9", "-e7", "--bw=99"],
])
def test_cli_options_take_effect_like_the_reference(built, args):
    """`<options> -h` prints the values in effect in both programs (the reference prints its help after parsing,
    src/main.cpp:421-520): every option shown by both must agree, for presets, macros and single options."""
    if "--bw=99" in args:  # ketopt also accepts --name=value; give the reference the spaced form of the same options
        ref_args = ["-k", "9", "-e", "7", "--bw", "99"]
    else:
        ref_args = args
    mine = run([CLI] + args + ["-h"])
    ref = run([REF_CLI] + ref_args + ["-h"])
    assert mine.returncode == 0 and ref.returncode == 0
    a, b = _values_in_help(mine.stdout), _values_in_help(ref.stdout)
    if "-t" not in args:  # deliberate difference: the reference defaults to 3 mapping threads, rawhash2_b200 to all hardware threads (host-side decode)
        assert b.pop("-t") == [3.0] and a.pop("-t")[0] >= 3
    assert a == b, {k: (a[k], b[k]) for k in a if a[k] != b[k]}
