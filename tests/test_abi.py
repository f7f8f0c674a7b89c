"""CPU: the C-ABI library loads, exports every symbol include/rawhash_b200.h declares, mirrors the
reference's option/preset/index semantics on the host side, and refuses to map without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _bind
from golden_util import CASES, GoldenCase, sha

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "rawhash_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rh_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(built):
    from rawhash_b200 import api
    lib = C.CDLL(api.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} is declared in include/rawhash_b200.h but not exported"
    assert set(api.EXPORTED_SYMBOLS) == set(syms), "api.py binds a different set than the header declares"


def test_struct_layouts_match_header(built):
    from rawhash_b200 import api
    assert C.sizeof(api.Params) == C.sizeof(_bind.Params) == 192
    assert C.sizeof(api.MapRec) == 60
    assert api.MAPREC_DTYPE.itemsize == 60


@pytest.mark.parametrize("case", CASES)
def test_params_presets_match_reference(built, case):
    """rh_params_init/preset/r10 against the bytes the reference's option code produced (golden)."""
    from rawhash_b200 import api
    g = GoldenCase(case)
    over = dict(sample_rate=g.sample_rate, bp_per_sec=g.bp_per_sec) if g.non_default_sampling() else {}
    P = api.make_params(g.preset, g.r10, **over)
    assert bytes(P) == g.z["params"].tobytes()


def test_unknown_preset_is_an_error(built):
    from rawhash_b200 import api
    with pytest.raises(api.RawHashError):
        api.make_params("no-such-preset")


@pytest.mark.parametrize("case", CASES)
def test_host_index_matches_reference(built, case):
    """rh_pore_load + rh_index_build + rh_index_update_mapopt + rh_index_get vs golden ri_idx_get answers."""
    from rawhash_b200 import api
    g = GoldenCase(case)
    over = dict(sample_rate=g.sample_rate, bp_per_sec=g.bp_per_sec) if g.non_default_sampling() else {}
    P = api.make_params(g.preset, g.r10, **over)
    pore = api.load_pore(g.model, g.k)
    assert sha(pore) == str(g.z["pore_vals_sha"])
    names, seqs = g.genome_strings()
    idx = api.Index.build(P, pore, names, seqs, 4)
    assert idx.n_seq == len(names) and [idx.seq_name(i) for i in range(idx.n_seq)] == names
    assert [idx.seq_len(i) for i in range(idx.n_seq)] == [len(s) for s in seqs]
    assert idx.update_mapopt(P) == g.mid_occ
    hs, off, pos = g.z["idx_hash"], g.z["idx_off"], g.z["idx_pos"]
    for j, h in enumerate(hs):
        assert np.array_equal(idx.get(int(h)), pos[int(off[j]):int(off[j + 1])])


def test_reference_ind_file_loads(built, tmp_path):
    """A `.ind` written by the reference format writer (oracle restatement of ri_idx_dump) loads through
    rh_index_load and answers lookups like the in-memory build."""
    from rawhash_b200 import api
    g = GoldenCase("r94_sensitive")
    lib = (_bind.RefLib if _bind.have_ref() else _bind.OracleLib)().open(g.preset, g.r10, g.model)
    ind = str(tmp_path / "ref.ind")
    lib.build_index(g.fasta, ind, 2)
    assert os.path.getsize(ind) > 0
    P = api.make_params(g.preset, g.r10)
    idx = api.Index.load(ind, P)
    assert idx.update_mapopt(P) == g.mid_occ
    hs, off, pos = g.z["idx_hash"], g.z["idx_off"], g.z["idx_pos"]
    for j, h in enumerate(hs):
        assert np.array_equal(idx.get(int(h)), pos[int(off[j]):int(off[j + 1])])
    names, _ = g.genome_strings()
    assert [idx.seq_name(i) for i in range(idx.n_seq)] == names


def test_paf_formatting_matches_reference_lines(built):
    """rh_format_paf on records rebuilt from the golden PAF text reproduces that text (mt:f: aside)."""
    from rawhash_b200 import api
    g = GoldenCase("r94_sensitive")
    P = api.make_params(g.preset, g.r10)
    names, seqs = g.genome_strings()
    idx = api.Index.build(P, api.load_pore(g.model, g.k), names, seqs, 2)
    lines = g.paf.splitlines()
    recs = np.zeros(len(lines), dtype=api.MAPREC_DTYPE)
    for i, ln in enumerate(lines):
        t = ln.split("\t")
        tags = {x[:2]: x[5:] for x in t[12:]}
        r = recs[i]
        r["read_idx"] = g.names.index(t[0]); r["read_length"] = int(t[1])
        r["ci"] = int(tags["ci"]); r["sl"] = int(tags["sl"]); r["cm"] = int(tags["cm"]); r["nc"] = int(tags["nc"]); r["s1"] = int(tags["s1"])
        if t[4] != "*":
            r["mapped"] = 1; r["read_start_position"] = int(t[2]); r["read_end_position"] = int(t[3]); r["rev"] = t[4] == "-"
            r["ref_id"] = names.index(t[5]); r["fragment_start_position"] = int(t[7]); r["fragment_length"] = int(t[8]) - int(t[7])
            r["mapq"] = int(t[11])
    out = _bind.strip_mt(idx.format_paf(recs, g.names))
    assert out == g.paf


def test_gpu_entry_points_fail_loudly_without_a_device(built):
    """No CPU fallback: on a box without CUDA rh_gpu_init returns NULL and says why."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from rawhash_b200 import api
    g = GoldenCase("r94_sensitive")
    P = api.make_params(g.preset, g.r10)
    names, seqs = g.genome_strings()
    idx = api.Index.build(P, api.load_pore(g.model, g.k), names, seqs, 2)
    idx.update_mapopt(P)
    with pytest.raises(api.RawHashError, match="CUDA|device"):
        api.Mapper(idx, P, 0, 1 << 20)


@pytest.mark.parametrize("case", CASES)
def test_format_paf_reprints_the_golden_paf(built, case):
    """rh_format_paf (the fprintf block of src/rmap.cpp:751-772) on CPU: the records parsed back out of the committed
    golden PAF — which the reference printed — must print to the same text, mapped and unmapped lines alike."""
    from rawhash_b200 import api
    g = GoldenCase(case)
    P = api.make_params(g.preset, g.r10)
    pore = api.load_pore(g.model, g.k)
    names, seqs = g.genome_strings()
    idx = api.Index.build(P, pore, names, seqs, 2)
    lines = _bind.strip_mt(g.paf).splitlines()
    recs = np.zeros(len(lines), dtype=api.MAPREC_DTYPE)
    order = {n: i for i, n in enumerate(g.names)}
    n_unmapped = 0
    for i, ln in enumerate(lines):
        f = ln.split("\t")
        tags = dict(t.split(":", 2)[::2] for t in f[12:])
        r = recs[i]
        r["read_idx"], r["read_length"] = order[f[0]], int(f[1])
        r["ci"], r["sl"], r["cm"], r["nc"], r["s1"] = (int(tags[k]) for k in ("ci", "sl", "cm", "nc", "s1"))
        if f[2] == "*":
            n_unmapped += 1
            r["mapq"] = int(f[11])
            continue
        r["mapped"], r["rev"], r["ref_id"] = 1, f[4] == "-", names.index(f[5])
        r["read_start_position"], r["read_end_position"] = int(f[2]), int(f[3])
        r["fragment_start_position"], r["fragment_length"], r["mapq"] = int(f[7]), int(f[10]), int(f[11])
        assert int(f[6]) == idx.seq_len(names.index(f[5])) and int(f[8]) == int(f[7]) + int(f[10]) and int(f[9]) == int(f[3]) - int(f[2]) - 1
    got = _bind.strip_mt(idx.format_paf(recs, g.names)).splitlines()
    assert got == lines
    assert len(lines) >= len(g.names)


def test_c99_example_compiles_links_and_fails_loudly_without_a_gpu(built, tmp_path):
    """examples/map_blow5.c: the boundary is usable from plain C99 (the reference's language) — compile with
    -pedantic, link against the library, and (on a box without a GPU) see rh_gpu_init refuse."""
    import shutil
    import subprocess
    from rawhash_b200 import api
    exe = str(tmp_path / "map_blow5")
    libdir = os.path.dirname(api.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-O2", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "map_blow5.c"), "-L" + libdir, "-lrawhash_b200", "-Wl,-rpath," + libdir, "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    g = GoldenCase("r94_sensitive", str(tmp_path))
    P = api.make_params(g.preset)
    pore = api.load_pore(g.model, g.k)
    names, seqs = g.genome_strings()
    ind, reads = str(tmp_path / "t.ind"), str(tmp_path / "reads.blow5")
    api.Index.build(P, pore, names, seqs, 2).dump(ind, pore)
    api.write_slow5(reads, g.names, g.raws, *g.cal)
    r = subprocess.run([exe, ind, reads], capture_output=True, text=True, timeout=300)
    has_gpu = shutil.which("nvidia-smi") is not None and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0
    if has_gpu:
        assert r.returncode == 0 and _bind.strip_mt(r.stdout) == _bind.strip_mt(g.paf)
    else:
        assert r.returncode == 1 and "no usable CUDA device" in r.stderr and r.stdout == ""
