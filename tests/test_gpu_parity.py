"""GPU parity: the CUDA path (through the C-ABI) against the oracle, bit for bit."""
import numpy as np
import pytest

from common import World, tap_equal

pytestmark = pytest.mark.gpu



def _same_records(a, b):
    """Record arrays equal in every field but mt_ms (a time)."""
    f = [n for n in a.dtype.names if n != "mt_ms"]
    return a.shape == b.shape and np.array_equal(a[f], b[f])

def _setup(world, preset="sensitive", r10=False, **over):
    from rawhash_b200 import api
    from _bind import OracleLib
    P = api.make_params(preset, r10, **over)
    pore = api.load_pore(world.model, world.k)
    names, seqs = world.genome_strings()
    idx = api.Index.build(P, pore, names, seqs, 8)
    idx.update_mapopt(P)
    orc = OracleLib().open(preset, r10, world.model)
    if "sample_rate" in over:
        orc.set_sampling(over["sample_rate"], over["bp_per_sec"])
    orc.build_index(world.fasta, "", 4)
    if "mid_occ" not in over:
        assert orc.mapopt_update() == P.mid_occ
    return api, P, idx, orc


def test_tap_stages_match_oracle(built):
    from rawhash_b200 import synth
    w = World(n_contigs=2, genome_len=400_000, n_reads=12, read_bp=3000, seed=1)
    api, P, idx, orc = _setup(w)
    m = api.Mapper(idx, P, 0, 1 << 30)
    for i in range(len(w.names)):
        got = m.tap_read(w.reads["raw"][i], synth.OFFSET, synth.RANGE, synth.DIGITISATION, w.names[i])
        exp = orc.tap_read(w.pa(i), w.names[i])
        assert tap_equal(got, exp) == [], f"read {i}"
        for g, e in zip(got, exp):
            assert g["cnt"][0] == e["cnt"][0] and g["cnt"][7] == e["cnt"][7]
    m.close()


def test_paf_matches_oracle(built):
    from rawhash_b200 import synth
    from _bind import strip_mt
    w = World(n_contigs=3, genome_len=1_000_000, n_reads=200, read_bp=5000, seed=2)
    api, P, idx, orc = _setup(w)
    m = api.Mapper(idx, P, 0, 2 << 30)
    recs = m.map_batch(w.reads["raw"], w.reads["offset"], w.reads["range"], w.reads["digitisation"], w.names)
    got = idx.format_paf(recs, w.names)
    exp, _ = orc.map_paf([w.pa(i) for i in range(len(w.names))], w.names, 4)
    assert strip_mt(got) == strip_mt(exp)
    st = m.stats()
    assert st["kernel_launches"] > 0 and st["n_reads"] == len(w.names)
    m.close()


def _check_paf(world, preset="sensitive", r10=False, extra_raw=None, arena=2 << 30, **over):
    from rawhash_b200 import synth
    from _bind import strip_mt
    api, P, idx, orc = _setup(world, preset, r10, **over)
    raws = list(world.reads["raw"]); names = list(world.names)
    if extra_raw:
        for k, r in enumerate(extra_raw):
            raws.append(r); names.append(f"extra_{k:03d}")
    n = len(raws)
    cal = (np.full(n, synth.OFFSET), np.full(n, synth.RANGE), np.full(n, synth.DIGITISATION))
    m = api.Mapper(idx, P, 0, arena)
    recs = m.map_batch(raws, *cal, names)
    got = strip_mt(idx.format_paf(recs, names)).splitlines()
    exp, _ = orc.map_paf([synth.raw_to_pa(r, synth.OFFSET, synth.RANGE, synth.DIGITISATION) for r in raws], names, 4)
    exp = strip_mt(exp).splitlines()
    m.close()
    assert len(got) == len(exp)
    bad = [(g, e) for g, e in zip(got, exp) if g != e]
    assert not bad, f"{len(bad)} of {len(exp)} PAF lines differ, first: {bad[0]}"
    return recs


def _repeat_genome(unit=3000, copies=40, flank=100_000, seed=9):
    """Tandem repeats: reads from them produce many anchors with identical target position
    (equal sort keys), exercising klib's tie order."""
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.integers(0, 4, unit, dtype=np.uint8)
    left = rng.integers(0, 4, flank, dtype=np.uint8); right = rng.integers(0, 4, flank, dtype=np.uint8)
    # a short internal tandem duplication inside the unit makes reads repeat their own hashes
    u[1000:1400] = u[600:1000]
    return [("rep1", np.concatenate([left] + [u] * copies + [right])), ("chr2", rng.integers(0, 4, 150_000, dtype=np.uint8))]


def test_tie_order_on_repeats(built):
    """Equal anchor keys and equal chain scores: tap every stage on reads from a tandem-repeat locus."""
    from rawhash_b200 import synth
    w = World(n_contigs=2, genome_len=200_000, n_reads=2, read_bp=1000, seed=3)
    w.genome = _repeat_genome()
    w.fasta = "/tmp/rh_world_repeat.fa"
    synth.write_fasta(w.fasta, w.genome)
    # reads sampled inside the repeat array
    import numpy as _np
    w.reads = synth.make_reads([("rep1", w.genome[0][1][100_000:220_000])], 10, 4000, w.k, w.means, w.stdv, seed=11)
    w.names = w.reads["names"]
    api, P, idx, orc = _setup(w, mid_occ=2000)
    orc.set_mid_occ(2000)
    m = api.Mapper(idx, P, 0, 2 << 30)
    n_ties = 0
    for i in range(len(w.names)):
        got = m.tap_read(w.reads["raw"][i], synth.OFFSET, synth.RANGE, synth.DIGITISATION, w.names[i])
        exp = orc.tap_read(w.pa(i), w.names[i])
        assert tap_equal(got, exp) == [], f"read {i}"
        for c in exp:
            x = c["anchors"][:, 0]
            n_ties += int((x[1:] == x[:-1]).sum())
    m.close()
    assert n_ties > 0, "test world produced no equal keys"


@pytest.mark.parametrize("preset", ["sensitive", "fast", "viral"])
def test_presets_paf(built, preset):
    w = World(n_contigs=4, genome_len=2_000_000 if preset != "viral" else 60_000, n_reads=150, read_bp=4000, seed=4)
    _check_paf(w, preset)


def test_unmappable_and_edge_reads(built):
    """Reads from another genome (all 10 chunks, carried anchors, unmapped output), empty reads,
    reads shorter than a chunk, reads whose samples are all outside (30,200) pA."""
    from rawhash_b200 import synth
    w = World(n_contigs=2, genome_len=800_000, n_reads=40, read_bp=6000, seed=6)
    other = synth.make_genome(1, 300_000, seed=77)
    alien = synth.make_reads(other, 40, 6000, w.k, w.means, w.stdv, seed=78)["raw"]
    extra = alien + [np.zeros(0, np.int16), np.full(5000, 2000, np.int16), w.reads["raw"][0][:1500], w.reads["raw"][1][:300],
                     w.reads["raw"][2][:4003]]
    recs = _check_paf(w, "sensitive", extra_raw=extra)
    assert (recs["mapped"] == 0).sum() > 0 and recs["ci"].max() == 10


def test_r10_paf(built):
    w = World(n_contigs=2, genome_len=600_000, n_reads=60, read_bp=3000, seed=8, kind="r10.4.1", sample_rate=5000.0, bp_per_sec=400.0)
    _check_paf(w, "sensitive", r10=True, sample_rate=5000, bp_per_sec=400)


def test_small_arena_groups(built):
    """A tiny anchor arena forces the chunk round to run in several groups; results must not change."""
    w = World(n_contigs=3, genome_len=1_000_000, n_reads=120, read_bp=5000, seed=2)
    _check_paf(w, "sensitive", arena=48 << 20)


def test_wide_sort_keys(built):
    """>32 varying key bits (many targets x long target): the 64-bit sort path + tie replay."""
    from rawhash_b200 import synth
    w = World(n_contigs=1, genome_len=300_000, n_reads=30, read_bp=3000, seed=12)
    rng = np.random.Generator(np.random.PCG64(99))
    small = [(f"s{i:05d}", rng.integers(0, 4, 120, dtype=np.uint8)) for i in range(40_000)]
    w.genome = [w.genome[0]] + small
    w.fasta = "/tmp/rh_world_wide.fa"
    synth.write_fasta(w.fasta, w.genome)
    api, P, idx, orc = _setup(w)
    m = api.Mapper(idx, P, 0, 2 << 30)
    for i in range(6):
        got = m.tap_read(w.reads["raw"][i], synth.OFFSET, synth.RANGE, synth.DIGITISATION, w.names[i])
        exp = orc.tap_read(w.pa(i), w.names[i])
        assert tap_equal(got, exp) == [], f"read {i}"
        x = exp[0]["anchors"][:, 0]
        assert len(x) and int(x.max() ^ x.min()).bit_length() > 32
    m.close()
    _check_paf(w, "sensitive")


def test_faster_preset_minimizers(built):
    """-x faster (w=3 minimizer seeding, ri_sketch_min): seeds and everything after, per chunk, then PAF."""
    from rawhash_b200 import synth
    w = World(n_contigs=3, genome_len=1_500_000, n_reads=80, read_bp=4000, seed=21)
    api, P, idx, orc = _setup(w, "faster")
    assert P.w > 0
    m = api.Mapper(idx, P, 0, 1 << 30)
    for i in range(8):
        got = m.tap_read(w.reads["raw"][i], synth.OFFSET, synth.RANGE, synth.DIGITISATION, w.names[i])
        exp = orc.tap_read(w.pa(i), w.names[i])
        assert tap_equal(got, exp) == [], f"read {i}"
    m.close()
    _check_paf(w, "faster")


def test_rawsamble_all_vs_all(built, monkeypatch):
    """-x ava: index built from the reads' own signals (event detection on the GPU), all-vs-all overlap,
    every chain reported (RI_M_ALL_CHAINS), hits filtered by read-name order (rmap.cpp:82-86).  Mapped a second
    time with a record arena far too small for one record per chain: the range is mapped again with the arena
    sized from the device's count, and the PAF is the same."""
    from rawhash_b200 import api, synth
    from _bind import OracleLib, strip_mt
    w = World(n_contigs=1, genome_len=60_000, n_reads=60, read_bp=4000, seed=23)
    P = api.make_params("ava")
    n = len(w.names)
    cal = (np.full(n, synth.OFFSET), np.full(n, synth.RANGE), np.full(n, synth.DIGITISATION))
    idx = api.Index.build_from_signals(P, w.names, w.reads["raw"], *cal)
    idx.update_mapopt(P)
    orc = OracleLib().open("ava", False, w.model)
    sigs = [w.pa(i) for i in range(n)]
    orc.build_index_sig(sigs, w.names, 4)
    assert orc.mapopt_update() == P.mid_occ
    m = api.Mapper(idx, P, 0, 1 << 30)
    recs = m.map_batch(w.reads["raw"], *cal, w.names)
    m.close()
    got = strip_mt(idx.format_paf(recs, w.names)).splitlines()
    exp, _ = orc.map_paf(sigs, w.names, 4)
    exp = strip_mt(exp).splitlines()
    assert len(exp) > n, "expected overlaps between reads"
    assert got == exp
    monkeypatch.setenv("RH_REC_CAP_TEST", "16")   # first guess: 16 records for 72
    m = api.Mapper(idx, P, 0, 1 << 30)
    recs = m.map_batch(w.reads["raw"], *cal, w.names)
    m.close()
    assert strip_mt(idx.format_paf(recs, w.names)).splitlines() == exp


def test_workers_and_device_resident_input(built, monkeypatch):
    """A batch cut into concurrent read ranges (RH_WORKERS=3, own stream and arenas each) and the same batch passed as
    one device-resident buffer give the records of the single-range host-buffer call, in input order."""
    import torch
    from rawhash_b200 import synth
    w = World(n_contigs=3, genome_len=900_000, n_reads=400, read_bp=3000, seed=31)
    api, P, idx, orc = _setup(w)
    n = len(w.names)
    cal = (np.full(n, synth.OFFSET), np.full(n, synth.RANGE), np.full(n, synth.DIGITISATION))
    m1 = api.Mapper(idx, P, 0, 1 << 30)
    ref = m1.map_batch(w.reads["raw"], *cal, w.names)
    # device-resident variant: one concatenated buffer + offsets (arbitrary, unaligned read starts)
    lens = np.array([len(r) for r in w.reads["raw"]], dtype=np.uint64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    flat = torch.from_numpy(np.concatenate(w.reads["raw"] + [np.zeros(8, np.int16)])).cuda()
    dev = m1.map_batch_device(flat.data_ptr(), off, *cal, names=w.names)
    m1.close()
    assert _same_records(ref, dev)
    monkeypatch.setenv("RH_WORKERS", "3")
    m3 = api.Mapper(idx, P, 0, 1 << 30)
    assert m3.set_workers(3) == 3
    got = m3.map_batch(w.reads["raw"], *cal, w.names)
    st = m3.stats()
    m3.close()
    assert _same_records(ref, got)
    assert st["n_reads"] == n and np.array_equal(got["read_idx"], np.sort(got["read_idx"]))


def test_scale_properties(built):
    """Size-independent properties on a batch the oracle would take minutes for: two runs agree record for record,
    host-buffer and device-resident inputs agree, every read comes back exactly once and in order, and the
    synthetic reads land on the locus they were sampled from."""
    import torch
    from rawhash_b200 import api, synth
    w = World(n_contigs=8, genome_len=6_000_000, n_reads=4000, read_bp=4000, seed=41)
    P = api.make_params("sensitive")
    names, seqs = w.genome_strings()
    idx = api.Index.build(P, api.load_pore(w.model, w.k), names, seqs, 8)
    idx.update_mapopt(P)
    n = len(w.names)
    cal = (np.full(n, synth.OFFSET), np.full(n, synth.RANGE), np.full(n, synth.DIGITISATION))
    m = api.Mapper(idx, P, 0, 4 << 30)
    a = m.map_batch(w.reads["raw"], *cal, w.names)
    b = m.map_batch(w.reads["raw"], *cal, w.names)
    lens = np.array([len(r) for r in w.reads["raw"]], dtype=np.uint64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    flat = torch.from_numpy(np.concatenate(w.reads["raw"] + [np.zeros(8, np.int16)])).cuda()
    c = m.map_batch_device(flat.data_ptr(), off, *cal, names=w.names)
    m.close()
    assert _same_records(a, b) and _same_records(a, c)
    assert (a["mt_ms"] > 0).all()                               # mt:f: = the read's share of the batch's stream time
    assert np.array_equal(a["read_idx"], np.arange(n))          # -x sensitive: one record per read, input order
    ok = 0
    for r, (ci, st, strand) in zip(a, w.reads["truth"]):
        if r["mapped"] and r["ref_id"] == ci and r["rev"] == strand and abs(int(r["fragment_start_position"]) - st) < 4000:
            ok += 1
    assert ok >= 0.97 * n, f"only {ok} of {n} reads mapped to their true locus"


def test_secondary_reporting_best_n(built):
    """best_n > 0 (mm_select_sub keeps secondaries, hit.c:338-367) — the general path of the decision kernel."""
    from rawhash_b200 import synth
    from _bind import strip_mt
    w = World(n_contigs=2, genome_len=500_000, n_reads=80, read_bp=4000, seed=43)
    api, P, idx, orc = _setup(w, "sensitive", best_n=3)
    orc.set_best_n(3)
    n = len(w.names)
    cal = (np.full(n, synth.OFFSET), np.full(n, synth.RANGE), np.full(n, synth.DIGITISATION))
    m = api.Mapper(idx, P, 0, 1 << 30)
    recs = m.map_batch(w.reads["raw"], *cal, w.names)
    m.close()
    got = strip_mt(idx.format_paf(recs, w.names)).splitlines()
    exp, _ = orc.map_paf([w.pa(i) for i in range(n)], w.names, 4)
    exp = strip_mt(exp).splitlines()
    assert got == exp


@pytest.mark.parametrize("kind,r10", [("r9.4", False), ("r10.4.1", True)])
def test_gpu_index_build_equals_host_build(built, kind, r10):
    """rh_index_build_gpu (events, diff filter, sketch of both strands, sort, key table on the device) against the
    host builder, which the CPU tests pin to the reference's ri_idx_get answers: same keys, same position lists."""
    from rawhash_b200 import api
    w = World(n_contigs=5, genome_len=1_200_000, n_reads=1, read_bp=1000, seed=51, kind=kind)
    P = api.make_params("sensitive", r10)
    pore = api.load_pore(w.model, w.k)
    names, seqs = w.genome_strings()
    seqs = [seqs[0].lower()] + seqs[1:] + ["ACGTACG"[: w.k - 1]]   # lower case accepted; a sequence shorter than k gives no events
    names = names + ["tiny"]
    host = api.Index.build(P, pore, names, seqs, 8)
    dev = api.Index.build_gpu(P, pore, names, seqs, 0)
    assert dev.n_seq == host.n_seq and dev.n_keys == host.n_keys and dev.n_pos == host.n_pos
    assert host.update_mapopt(api.make_params("sensitive", r10)) == dev.update_mapopt(api.make_params("sensitive", r10))
    nk = host.n_keys
    for i in list(range(0, nk, 37)) + [nk - 1]:
        h = host.key(i)
        assert dev.key(i) == h
        assert np.array_equal(dev.get(h), host.get(h)), f"key {h:#x}"
