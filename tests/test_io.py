"""CPU: the file surfaces either side of the path (csrc/rh_io.cpp) against the reference's own code —
`.ind` writer vs ri_idx_dump/ri_idx_load (src/rindex.c:545-776), SLOW5/BLOW5 reader and writer vs the reference's
vendored slow5lib (extern/slow5lib, compiled into oracle/_ref/libslow5_tap.so), FASTA reader vs the genome written."""
import ctypes as C
import glob
import gzip
import os

import numpy as np
import pytest

import _bind
from common import World

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
S5TAP = os.path.join(ROOT, "oracle", "_ref", "libslow5_tap.so")
S5LIB = "/root/reference/extern/slow5lib"
needs_s5 = pytest.mark.skipif(not os.path.isfile(S5TAP), reason="oracle/_ref/libslow5_tap.so not built")
needs_ref = pytest.mark.skipif(not _bind.have_ref(), reason="oracle/_ref/libref_tap.so not built")


class Slow5Lib:
    """slow5_open / slow5_get_next of the reference's slow5lib."""

    def __init__(self):
        T = C.CDLL(S5TAP)
        T.s5tap_open.restype = C.c_void_p
        T.s5tap_open.argtypes = [C.c_char_p]
        T.s5tap_next.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.POINTER(C.c_int16)), C.POINTER(C.c_uint64)] + [C.POINTER(C.c_double)] * 4
        T.s5tap_close.argtypes = [C.c_void_p]
        self.T = T

    def read(self, path):
        h = self.T.s5tap_open(path.encode())
        assert h, f"slow5lib cannot open {path}"
        out = []
        while True:
            rid, raw, n = C.c_char_p(), C.POINTER(C.c_int16)(), C.c_uint64()
            d = [C.c_double() for _ in range(4)]
            rc = self.T.s5tap_next(h, C.byref(rid), C.byref(raw), C.byref(n), *[C.byref(x) for x in d])
            assert rc >= 0, f"slow5lib error {rc} on {path}"
            if rc == 0:
                break
            sig = np.ctypeslib.as_array(raw, shape=(n.value,)).copy() if n.value else np.zeros(0, np.int16)
            out.append((rid.value.decode(), sig, d[0].value, d[1].value, d[2].value, d[3].value))  # id raw digitisation offset range rate
        self.T.s5tap_close(h)
        return out


def read_mine(path, threads=3, **kw):
    from rawhash_b200 import api
    f = api.SignalFile(path, threads)
    out, n_batches = [], 0
    while True:
        b = f.next_batch(**kw)
        if b is None:
            break
        n_batches += 1
        assert b["contiguous"], "a batch must hold its reads back to back in one arena"
        for i in range(len(b["names"])):
            out.append((b["names"][i], b["raw"][i], b["digitisation"][i], b["offset"][i], b["range"][i], b["sampling_rate"][i]))
    f.close()
    return out, n_batches


def same_records(a, b):
    return len(a) == len(b) and all(x[0] == y[0] and np.array_equal(x[1], y[1]) and tuple(x[2:]) == tuple(y[2:]) for x, y in zip(a, b))


def _edge_reads(seed=3):
    rng = np.random.default_rng(seed)
    raws = [rng.integers(-32768, 32767, size=n).astype(np.int16) for n in (0, 1, 3, 4, 5, 1000, 77777)]  # full-range deltas: 1-3 byte codes
    raws.append(np.cumsum(rng.integers(-300, 300, size=50000)).clip(-30000, 30000).astype(np.int16))
    raws.append(np.full(4097, -32768, np.int16))
    names = [f"read_{i}" for i in range(len(raws))]
    return names, raws, rng.uniform(-20, 20, len(raws)), np.full(len(raws), 1402.882), np.full(len(raws), 8192.0)


@needs_s5
@pytest.mark.skipif(not os.path.isdir(S5LIB), reason="reference tree not present")
def test_reader_equals_slow5lib_on_the_reference_files(built):
    """Every SLOW5/BLOW5 file shipped in the reference tree that its slow5lib can read: same ids, samples, calibration
    (ASCII; binary with none/zlib records and none/svb-zd signals; auxiliary fields; two read groups)."""
    files = [f"{S5LIB}/examples/example.slow5", f"{S5LIB}/examples/example2.slow5", f"{S5LIB}/examples/adv/example3.blow5"]
    for sub in ("aux_array", "one_fast5", "two_rg"):
        files += sorted(glob.glob(f"{S5LIB}/test/data/exp/{sub}/*.blow5"))
    files += [f"{S5LIB}/test/data/exp/aux_array/exp_lossless.slow5", f"{S5LIB}/test/data/exp/aux_array/exp_lossless_end_reason.slow5"]
    s5 = Slow5Lib()
    n = 0
    for p in files:
        exp = s5.read(p)
        got, _ = read_mine(p, max_samples=100_000)
        assert same_records(got, exp), p
        n += len(exp)
    assert len(files) >= 13 and n >= 30


@needs_s5
@pytest.mark.parametrize("fname,rec,sig", [("a.slow5", 0, 0), ("b.blow5", 0, 0), ("c.blow5", 1, 0), ("d.blow5", 0, 1), ("e.blow5", 1, 1),
                                           ("f.blow5", 2, 0), ("g.blow5", 2, 1)])  # 2 = zstd records (libzstd.so.1 bound at run time)
def test_writer_is_read_by_slow5lib_and_round_trips(built, tmp_path, fname, rec, sig):
    from rawhash_b200 import api
    names, raws, off, rg, dg = _edge_reads()
    p = str(tmp_path / fname)
    api.write_slow5(p, names, raws, off, rg, dg, 4000.0, rec, sig)
    want = [(n, r, g, o, q, 4000.0) for n, r, o, q, g in zip(names, raws, off, rg, dg)]
    assert same_records(Slow5Lib().read(p), want)
    got, nb = read_mine(p)
    assert same_records(got, want) and nb == 1


def test_batches_follow_the_minibatch_rule(built, tmp_path):
    """ri_sig_read_frag (src/rmap.cpp:600-660): reads are appended until the running sample count reaches the limit."""
    from rawhash_b200 import api
    rng = np.random.default_rng(1)
    raws = [rng.integers(200, 900, size=int(n)).astype(np.int16) for n in rng.integers(50, 400, size=300)]
    names = [f"r{i}" for i in range(300)]
    p = str(tmp_path / "m.blow5")
    api.write_slow5(p, names, raws, 10.0, 1400.0, 8192.0)
    f = api.SignalFile(p, 2)
    seen, sizes = [], []
    while (b := f.next_batch(max_samples=5000)) is not None:
        n = [len(r) for r in b["raw"]]
        assert sum(n) >= 5000 > sum(n[:-1]) or len(seen) + len(n) == 300
        seen += b["names"]
        sizes.append(len(n))
    assert seen == names and len(sizes) > 5
    f = api.SignalFile(p, 2)
    assert len(f.next_batch(max_reads=7)["names"]) == 7 and f.next_batch(max_reads=7)["names"] == names[7:14]
    f.close()


def test_reader_rejects_damaged_files(built, tmp_path):
    from rawhash_b200 import api
    names, raws, off, rg, dg = _edge_reads()
    p = str(tmp_path / "ok.blow5")
    api.write_slow5(p, names, raws, off, rg, dg)
    blob = open(p, "rb").read()
    cut = str(tmp_path / "cut.blow5")
    open(cut, "wb").write(blob[: len(blob) // 2])
    with pytest.raises(api.RawHashError):
        f = api.SignalFile(cut, 1)
        while f.next_batch() is not None:
            pass
    noeof = str(tmp_path / "noeof.blow5")
    open(noeof, "wb").write(blob[:-5])
    with pytest.raises(api.RawHashError):
        f = api.SignalFile(noeof, 1)
        while f.next_batch() is not None:
            pass
    bad = str(tmp_path / "bad.blow5")
    open(bad, "wb").write(b"NOTBLOW5" + blob[8:])
    with pytest.raises(api.RawHashError):
        api.SignalFile(bad, 1)
    with pytest.raises(api.RawHashError):
        api.SignalFile(str(tmp_path / "reads.fast5"), 1)
    with pytest.raises(api.RawHashError):
        api.SignalFile(str(tmp_path / "missing.blow5"), 1)


def test_find_signal_files(built, tmp_path):
    from rawhash_b200 import api
    (tmp_path / "sub" / "deeper").mkdir(parents=True)
    for rel in ("b.blow5", "a.slow5", "sub/c.blow5", "sub/deeper/d.slow5", "sub/notes.txt"):
        (tmp_path / rel).write_bytes(b"")
    got = [os.path.relpath(p, tmp_path) for p in api.find_signal_files(str(tmp_path))]
    assert got == ["a.slow5", "b.blow5", "sub/c.blow5", "sub/deeper/d.slow5"]
    assert api.find_signal_files(str(tmp_path / "b.blow5")) == [str(tmp_path / "b.blow5")]
    assert api.find_signal_files(str(tmp_path / "sub" / "notes.txt")) == []


def test_fasta_reader(built, tmp_path):
    from rawhash_b200 import api
    w = World(n_contigs=3, genome_len=30_000, n_reads=1, read_bp=500, seed=4)
    names, seqs = w.genome_strings()
    seqs = [s if isinstance(s, bytes) else s.encode() for s in seqs]
    n2, s2 = api.read_fasta(w.fasta)
    assert n2 == list(names) and s2 == seqs
    gz = str(tmp_path / "g.fa.gz")
    with gzip.open(gz, "wb") as f:  # comments after the name, CRLF, lower case, blank line, a FASTQ record
        f.write(b">chrA some comment\r\nACGT\r\nacgtn\r\n\r\n>chrB\tx\nGG\n@q1 d\nACGTA\n+\n>>>>>\n>chrC\n\n")
    n3, s3 = api.read_fasta(gz)
    assert n3 == ["chrA", "chrB", "q1", "chrC"] and s3 == [b"ACGTacgtn", b"GG", b"ACGTA", b""]
    with pytest.raises(api.RawHashError):
        api.read_fasta(str(tmp_path / "missing.fa"))


@needs_ref
@pytest.mark.parametrize("preset,seed,glen", [("sensitive", 11, 300_000), ("faster", 5, 600_000), ("viral", 2, 40_000)])
def test_ind_writer_equals_reference_dump(built, tmp_path, preset, seed, glen):
    """rh_index_dump vs ri_idx_dump: identical bytes except the two heap pointers the reference writes raw inside
    ri_pore_t (file offsets 46..61); the reference loads our file and answers ri_idx_get identically; so do we."""
    from rawhash_b200 import api
    w = World(n_contigs=3, genome_len=glen, n_reads=1, read_bp=500, seed=seed)
    P = api.make_params(preset)
    pore = api.load_pore(w.model, w.k)
    names, seqs = w.genome_strings()
    idx = api.Index.build(P, pore, names, seqs, 4)
    mine, theirs = str(tmp_path / "mine.ind"), str(tmp_path / "ref.ind")
    idx.dump(mine, pore)
    ref = _bind.RefLib().open(preset, False, w.model)
    ref.build_index(w.fasta, theirs, 2)
    a, b = np.fromfile(mine, np.uint8), np.fromfile(theirs, np.uint8)
    assert a.size == b.size
    diff = np.nonzero(a != b)[0]
    assert diff.size <= 16 and (diff.size == 0 or (diff.min() >= 46 and diff.max() < 62)), diff[:20]
    ref_on_mine = _bind.RefLib().open(preset, False, w.model)
    ref_on_mine.build_index(mine, "", 2)  # ri_idx_reader_open recognises the magic and calls ri_idx_load
    P2 = api.make_params(preset)
    back = api.Index.load(mine, P2)
    assert bytes(P2)[:44] == bytes(P)[:44] and back.n_keys == idx.n_keys and back.n_pos == idx.n_pos
    assert [back.seq_name(i) for i in range(back.n_seq)] == list(names)
    rng = np.random.default_rng(0)
    probes = [idx.key(int(i)) for i in rng.integers(0, idx.n_keys, 1500)] + [int(h) for h in rng.integers(0, 2**32, 500)]
    for h in probes:
        exp = idx.get(h)
        assert np.array_equal(ref_on_mine.idx_get(h), exp) and np.array_equal(back.get(h), exp)
    assert ref_on_mine.mapopt_update() == ref.mapopt_update()
    # the loader, exhaustively: the reference's own file read by rh_index_load and written back gives the same bytes
    again = str(tmp_path / "again.ind")
    api.Index.load(theirs, api.make_params(preset)).dump(again, pore)
    c = np.fromfile(again, np.uint8)
    d2 = np.nonzero(c != b)[0]
    assert c.size == b.size and d2.size <= 16 and (d2.size == 0 or (d2.min() >= 46 and d2.max() < 62))


@needs_ref
def test_ind_writer_signal_index_without_pore(built, tmp_path):
    """A hand-made index with an empty bucket run, a singleton and a long list; no pore table (signal index)."""
    from rawhash_b200 import api
    w = World(n_contigs=1, genome_len=20_000, n_reads=1, read_bp=500, seed=9)
    P = api.make_params("sensitive")
    pore = api.load_pore(w.model, w.k)
    names, seqs = w.genome_strings()
    idx = api.Index.build(P, pore, names, seqs, 1)
    p = str(tmp_path / "nopore.ind")
    idx.dump(p, None)
    back = api.Index.load(p, api.make_params("sensitive"))
    assert back.n_keys == idx.n_keys and back.n_pos == idx.n_pos
    for i in range(0, idx.n_keys, 97):
        assert np.array_equal(back.get(idx.key(i)), idx.get(idx.key(i)))


def test_own_inflate_equals_zlib(built):
    """csrc/rh_inflate.h (the BLOW5 record decompressor) against Python's zlib: stored, fixed and dynamic blocks, long
    codes (sub-tables), every level/strategy, small windows, multi-block streams; damaged streams are errors."""
    import random
    import zlib
    from rawhash_b200 import api
    rng = np.random.default_rng(5)
    random.seed(5)
    skew = np.array([1.6 ** -i for i in range(1, 256)])
    skew /= skew.sum()
    datas = [b"", b"a", b"hello hello hello hello hello", bytes(70000), bytes(rng.integers(0, 256, 200000, dtype=np.uint8)),
             bytes(rng.integers(0, 4, 300000, dtype=np.uint8)), open(__file__, "rb").read() * 3,
             np.cumsum(rng.integers(-40, 40, 200000)).astype(np.int16).tobytes(), (b"abc" * 7 + b"Z") * 9000,
             bytes(rng.choice(255, size=600000, p=skew).astype(np.uint8))]
    datas += [bytes(rng.integers(0, 256, n, dtype=np.uint8)) for n in (2, 3, 257, 258, 259, 65535, 65536, 65537)]
    n = 0
    for data in datas:
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_HUFFMAN_ONLY, zlib.Z_FIXED, zlib.Z_RLE):
                for wbits in (15, 9):
                    c = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
                    assert api.zlib_inflate(c.compress(data) + c.flush()) == data, (len(data), level, strategy, wbits)
                    n += 1
    assert n >= 500
    c, parts, blob = zlib.compressobj(6), [], b""
    for i in range(40):
        d = bytes(rng.integers(0, 256, int(rng.integers(1, 5000)), dtype=np.uint8)) if i % 2 else b"run" * int(rng.integers(1, 3000))
        blob += d
        parts += [c.compress(d), c.flush(random.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_NO_FLUSH, zlib.Z_BLOCK]))]
    assert api.zlib_inflate(b"".join(parts) + c.flush()) == blob
    z = zlib.compress(open(__file__, "rb").read(), 6)
    for k in range(400):
        b = bytearray(z)
        b[random.randrange(len(b))] ^= 1 << random.randrange(8)
        try:  # same verdict as zlib: a flipped padding bit before the Adler-32 trailer is harmless, anything else is an error
            want = zlib.decompress(bytes(b))
        except zlib.error:
            want = None
        if want is None:
            with pytest.raises(api.RawHashError):
                api.zlib_inflate(bytes(b))
        else:
            assert api.zlib_inflate(bytes(b)) == want
    for cut in range(0, len(z), 211):
        with pytest.raises(api.RawHashError):
            api.zlib_inflate(z[:cut])


@needs_s5
@pytest.mark.skipif(not os.path.isdir(S5LIB), reason="reference tree not present")
def test_reader_takes_zstd_files_written_by_slow5lib(built, tmp_path):
    """The other direction for zstd: a file written by the reference's slow5lib (zstd records + svb-zd signals, the
    combination slow5tools calls `-c zstd -s svb-zd`) read by rh_sigfile_*."""
    src = os.path.join(ROOT, "oracle", "slow5_tap.c")
    if "s5tap_convert" not in open(src).read():
        pytest.skip("slow5_tap.c has no converter")
    T = C.CDLL(S5TAP)
    T.s5tap_convert.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    out = str(tmp_path / "z.blow5")
    rc = T.s5tap_convert(f"{S5LIB}/examples/example.slow5".encode(), out.encode(), 2, 1)
    if rc == -100:
        pytest.skip("slow5lib built without zstd")
    assert rc == 5
    exp = Slow5Lib().read(f"{S5LIB}/examples/example.slow5")
    got, _ = read_mine(out)
    assert same_records(got, exp)
    assert open(out, "rb").read()[9] == 2  # record compression byte: zstd


@pytest.mark.parametrize("preset,group", [("sensitive", 120_000), ("sensitive", 250_000), ("faster", 120_000), ("sensitive", 1)])
def test_index_built_over_contig_groups_equals_one_pass(built, tmp_path, monkeypatch, preset, group):
    """rh_index_build_grouped (how references beyond one GPU pass are built: contig groups, host merge with rebased
    sequence ids) gives the very same index as a single pass — compared through the `.ind` bytes."""
    from rawhash_b200 import api
    w = World(n_contigs=5, genome_len=500_000, n_reads=1, read_bp=500, seed=29)
    P = api.make_params(preset)
    pore = api.load_pore(w.model, w.k)
    names, seqs = w.genome_strings()
    monkeypatch.delenv("RH_INDEX_GROUP_BASES", raising=False)
    whole = api.Index.build(P, pore, names, seqs, 4)
    monkeypatch.setenv("RH_INDEX_GROUP_BASES", str(group))
    parts = api.Index.build(P, pore, names, seqs, 4)
    assert parts.n_keys == whole.n_keys and parts.n_pos == whole.n_pos and parts.n_seq == whole.n_seq == 5
    a, b = str(tmp_path / "whole.ind"), str(tmp_path / "parts.ind")
    whole.dump(a, pore)
    parts.dump(b, pore)
    assert open(a, "rb").read() == open(b, "rb").read()


@needs_s5
def test_random_files_round_trip_and_match_slow5lib(built, tmp_path):
    """40 seeded random files: read count, lengths (incl. empty reads), sample statistics (noise, steps, saturated),
    calibration, every compression pair, random mini-batch limits — our reader == slow5lib == what was written."""
    from rawhash_b200 import api
    rng = np.random.default_rng(77)
    s5 = Slow5Lib()
    for t in range(40):
        n = int(rng.integers(0, 12))
        raws = []
        for _ in range(n):
            ln = int(rng.choice([0, 1, 2, 7, 100, 5000, 40000]))
            kind = int(rng.integers(0, 4))
            if kind == 0:
                r = rng.integers(-32768, 32768, ln)
            elif kind == 1:
                r = np.cumsum(rng.integers(-30, 31, ln)) + 500
            elif kind == 2:
                r = np.repeat(rng.integers(200, 900, ln // 9 + 1), 9)[:ln] + rng.integers(-3, 4, ln)
            else:
                r = np.full(ln, int(rng.choice([-32768, 32767, 0])))
            raws.append(np.clip(r, -32768, 32767).astype(np.int16))
        names = [f"r{t}_{i}_{'x' * int(rng.integers(0, 40))}" for i in range(n)]
        off, rg, dg = rng.uniform(-300, 300, n), rng.uniform(200, 3000, n), rng.choice([2048.0, 8192.0], n)
        rec, sig = int(rng.integers(0, 3)), int(rng.integers(0, 2))
        ext = "slow5" if (rec == 0 and sig == 0 and t % 3 == 0) else "blow5"
        p = str(tmp_path / f"f{t}.{ext}")
        api.write_slow5(p, names, raws, off, rg, dg, 5000.0, rec, sig)
        want = [(nm, r, g, o, q, 5000.0) for nm, r, o, q, g in zip(names, raws, off, rg, dg)]
        assert same_records(s5.read(p), want), (t, ext, rec, sig)
        got, _ = read_mine(p, threads=int(rng.integers(1, 5)), max_samples=int(rng.choice([1, 1000, 10**9])), max_reads=int(rng.choice([0, 1, 3])))
        assert same_records(got, want), (t, ext, rec, sig)


def test_reader_rejects_a_crafted_signal_length(built, tmp_path):
    """An uncompressed record whose length field is 2^63 + 2: twice that wraps to 4, which would pass a bounds check made
    after the multiplication and hand a batch with an 18-byte arena and 2^63 + 2 samples to the mapper."""
    import struct
    from rawhash_b200 import api
    raws = [np.arange(64, dtype=np.int16)]
    p = str(tmp_path / "one.blow5")
    api.write_slow5(p, ["r0"], raws, [0.0], [1402.882], [8192.0], 4000.0, 0, 0)
    blob = bytearray(open(p, "rb").read())
    at = blob.find(struct.pack("<Q", 64) + raws[0].tobytes())
    assert at > 0, "length field followed by the samples"
    for crafted in ((1 << 63) + 2, (1 << 64) - 1, 1 << 40):
        blob[at:at + 8] = struct.pack("<Q", crafted)
        bad = str(tmp_path / "crafted.blow5")
        open(bad, "wb").write(bytes(blob))
        with pytest.raises(api.RawHashError):
            f = api.SignalFile(bad, 1)
            while f.next_batch(1 << 62) is not None:   # a mini-batch limit that does not stop the record by itself
                pass
