"""CPU: the oracle against the compiled reference (oracle/_ref/libref_tap.so) run live on more
seeded worlds than the committed fixtures cover.  Skipped where the reference .so is absent."""
import numpy as np
import pytest

import _bind
from common import World, tap_equal

pytestmark = pytest.mark.skipif(not _bind.have_ref(), reason="oracle/_ref/libref_tap.so not built (needs /root/reference)")


def _pair(w, preset, r10=False, sampling=None):
    ref = _bind.RefLib().open(preset, r10, w.model)
    orc = _bind.OracleLib().open(preset, r10, w.model)
    if sampling:
        ref.set_sampling(*sampling); orc.set_sampling(*sampling)
    assert bytes(ref.params()) == bytes(orc.params())
    ref.build_index(w.fasta, "", 4); orc.build_index(w.fasta, "", 4)
    assert ref.mapopt_update() == orc.mapopt_update()
    return ref, orc


@pytest.mark.parametrize("preset,seed", [("sensitive", 31), ("fast", 32), ("viral", 33), ("faster", 34)])
def test_taps_and_paf(built, preset, seed):
    w = World(n_contigs=3, genome_len=600_000 if preset != "viral" else 50_000, n_reads=24, read_bp=3500, seed=seed)
    ref, orc = _pair(w, preset)
    sigs = [w.pa(i) for i in range(len(w.names))]
    for i in range(0, len(sigs), 3):
        assert tap_equal(orc.tap_read(sigs[i], w.names[i]), ref.tap_read(sigs[i], w.names[i])) == [], f"read {i}"
    a, _ = ref.map_paf(sigs, w.names, 2)
    b, _ = orc.map_paf(sigs, w.names, 2)
    assert _bind.strip_mt(a) == _bind.strip_mt(b)


def test_r10(built):
    w = World(n_contigs=2, genome_len=300_000, n_reads=12, read_bp=2500, seed=41, kind="r10.4.1", sample_rate=5000.0, bp_per_sec=400.0)
    ref, orc = _pair(w, "sensitive", True, (5000, 400))
    sigs = [w.pa(i) for i in range(len(w.names))]
    for i in range(0, len(sigs), 4):
        assert tap_equal(orc.tap_read(sigs[i], w.names[i]), ref.tap_read(sigs[i], w.names[i])) == []
    a, _ = ref.map_paf(sigs, w.names, 2)
    b, _ = orc.map_paf(sigs, w.names, 2)
    assert _bind.strip_mt(a) == _bind.strip_mt(b)


def test_events_and_sketch_functions(built):
    """detect_events with carried sums over consecutive chunks, ri_sketch on the result."""
    w = World(n_contigs=1, genome_len=100_000, n_reads=4, read_bp=3000, seed=51)
    ref = _bind.RefLib().open("sensitive", False, w.model)
    orc = _bind.OracleLib().open("sensitive", False, w.model)
    for i in range(4):
        pa = w.pa(i)
        sr, so = [0.0, 0.0, 0], [0.0, 0.0, 0]
        for c in range(0, len(pa), 4000):
            er = ref.detect_events(pa[c:c + 4000], sr)
            eo = orc.detect_events(pa[c:c + 4000], so)
            assert np.array_equal(er.view(np.uint32), eo.view(np.uint32))
            assert sr == so
            assert np.array_equal(ref.sketch(er, 3, 0), orc.sketch(eo, 3, 0))


def test_klib_sort_tie_order(built):
    ref, orc = _bind.RefLib(), _bind.OracleLib()
    rng = np.random.Generator(np.random.PCG64(5))
    for n, nk, sh in [(10, 3, 0), (65, 4, 0), (400, 5, 8), (9000, 70, 16), (30000, 500, 20), (30000, 3, 56)]:
        x = rng.integers(0, nk, n).astype(np.uint64) << np.uint64(sh)
        xy = np.stack([x, np.arange(n, dtype=np.uint64)], axis=1)
        assert np.array_equal(ref.radix_sort_128x(xy), orc.radix_sort_128x(xy))
        assert np.array_equal(ref.radix_sort_64(x), orc.radix_sort_64(x))


def test_rawsamble_ava(built):
    """-x ava: signal-target index built from the reads, all-vs-all overlaps (config 5 of BASELINE.json)."""
    w = World(n_contigs=1, genome_len=60_000, n_reads=60, read_bp=4000, seed=23)
    sigs = [w.pa(i) for i in range(len(w.names))]
    out = []
    for L in (_bind.RefLib, _bind.OracleLib):
        o = L().open("ava", False, w.model)
        o.build_index_sig(sigs, w.names, 4)
        mid = o.mapopt_update()
        paf, _ = o.map_paf(sigs, w.names, 2)
        out.append((mid, _bind.strip_mt(paf)))
    assert out[0] == out[1]
    assert len(out[0][1].splitlines()) > len(w.names)


def test_best_n_secondaries(built):
    """best_n > 0: mm_select_sub keeps secondaries (hit.c:338-367), which changes nc/mapq of the reported chain."""
    w = World(n_contigs=2, genome_len=500_000, n_reads=80, read_bp=4000, seed=43)
    ref, orc = _pair(w, "sensitive")
    ref.set_best_n(3); orc.set_best_n(3)
    sigs = [w.pa(i) for i in range(len(w.names))]
    for i in range(0, len(sigs), 8):
        assert tap_equal(orc.tap_read(sigs[i], w.names[i]), ref.tap_read(sigs[i], w.names[i])) == [], f"read {i}"
    a, _ = ref.map_paf(sigs, w.names, 2)
    b, _ = orc.map_paf(sigs, w.names, 2)
    assert _bind.strip_mt(a) == _bind.strip_mt(b)
    ref.set_best_n(0)
    c, _ = ref.map_paf(sigs, w.names, 2)
    assert _bind.strip_mt(a) != _bind.strip_mt(c), "best_n had no observable effect in this world"


def test_reference_index_filled_from_flat_table(built):
    """ref_index_from_flat (the reference's ri_idx_t filled from keys/off/pos, used for human-size worlds) serves the
    same ri_idx_get answers and the same PAF as the index the reference builds itself from the FASTA."""
    from rawhash_b200 import api
    w = World(n_contigs=3, genome_len=500_000, n_reads=16, read_bp=3000, seed=61)
    P = api.make_params("fast")
    names, seqs = w.genome_strings()
    idx = api.Index.build(P, api.load_pore(w.model, w.k), names, seqs, 4)
    keys, off, pos = idx.flat()
    a = _bind.RefLib().open("fast", False, w.model)
    a.build_index(w.fasta, "", 4)
    b = _bind.RefLib().open("fast", False, w.model)
    b.index_from_flat(names, [len(s) for s in seqs], keys, off, pos, 4)
    assert a.mapopt_update() == b.mapopt_update()
    rng = np.random.Generator(np.random.PCG64(3))
    probe = list(keys[rng.integers(0, len(keys), 3000)]) + list(rng.integers(0, 1 << 32, 500))
    for h in probe:
        assert np.array_equal(a.idx_get(int(h)), b.idx_get(int(h)))
    sigs = [w.pa(i) for i in range(len(w.names))]
    pa, _ = a.map_paf(sigs, w.names, 2)
    pb, _ = b.map_paf(sigs, w.names, 2)
    assert _bind.strip_mt(pa) == _bind.strip_mt(pb)
