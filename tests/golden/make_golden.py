#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref_tap.so).

Run in the build container (where /root/reference exists and `make -C oracle` has compiled it):

    python tests/golden/make_golden.py

The reference ships no golden vectors for the mapping path, so these fixtures are
outputs of the compiled reference itself on seeded synthetic inputs.  Everything needed to replay
a case is stored in the fixture (seed of the synthetic pore-model table, genome, raw int16 reads,
calibration), so the tests never read /root/reference.  Stored per case:

    params         rh_params_t bytes as the reference's option code leaves them (preset [+ --r10])
    pore_vals      load_pore output (z-normalised k-mer levels; k=9: SHA-256 only, the table is 1 MB)
    mid_occ        ri_mapopt_update result
    idx_hash/idx_off/idx_pos   ri_idx_get answers for a sample of hashes (present and absent)
    per read + chunk: events (float bits), seeds, u[], chain anchors, carried anchors, regions,
                   counts, and a SHA-256 of the sorted anchor array (the arrays are large)
    paf            the reference's PAF text (mt:f: removed)
    quant_in/out   dynamic_quantize over a grid
    sort_in/out    radix_sort_128x on an array with heavy key ties (klib's unstable order)
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from rawhash_b200 import synth  # noqa: E402
import _bind  # noqa: E402

CASES = {
    # name: (preset, r10, k, genome_len, n_contigs, n_reads, read_bp, sample_rate, bp_per_sec, extra)
    "r94_sensitive": dict(preset="sensitive", r10=False, k=6, genome_len=240_000, n_contigs=2, n_reads=6, read_bp=2200, sr=4000, bps=450),
    "r94_fast": dict(preset="fast", r10=False, k=6, genome_len=240_000, n_contigs=3, n_reads=4, read_bp=1800, sr=4000, bps=450),
    "r10_sensitive": dict(preset="sensitive", r10=True, k=9, genome_len=160_000, n_contigs=2, n_reads=3, read_bp=1500, sr=5000, bps=400),
}


def model_text(k: int, seed: int = 7) -> str:
    """Seeded synthetic k-mer table (the ONT tables are third-party data and are not committed)."""
    p = f"/tmp/rh_golden_model_{k}.tsv"
    synth.write_synthetic_model(p, k, seed)
    return open(p).read()


def pack2(seq: np.ndarray) -> np.ndarray:
    pad = (-len(seq)) % 4
    s = np.concatenate([seq, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return (s[:, 0] | (s[:, 1] << 2) | (s[:, 2] << 4) | (s[:, 3] << 6)).astype(np.uint8)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_case(name: str, c: dict) -> dict:
    k = c["k"]
    txt = model_text(k)
    mp = f"/tmp/rh_golden_model_{k}.tsv"
    means, stdv = synth.load_model_pa(mp, k)
    seed = sum(name.encode()) + 1000
    genome = synth.make_genome(c["n_contigs"], c["genome_len"], seed=seed)
    fa = f"/tmp/rh_golden_{name}.fa"
    synth.write_fasta(fa, genome)
    rd = synth.make_reads(genome, c["n_reads"], c["read_bp"], k, means, stdv, sample_rate=c["sr"], bp_per_sec=c["bps"], seed=seed + 1)
    # one read from a foreign genome (walks all chunks, ends unmapped) and one short read
    alien = synth.make_reads(synth.make_genome(1, 50_000, seed=seed + 2), 1, c["read_bp"] * 2, k, means, stdv,
                             sample_rate=c["sr"], bp_per_sec=c["bps"], seed=seed + 3)["raw"][0]
    raws = rd["raw"] + [alien, rd["raw"][0][:900], np.zeros(0, np.int16)]
    names = rd["names"] + ["alien_0", "short_0", "empty_0"]

    ref = _bind.RefLib().open(c["preset"], c["r10"], mp)
    if c["sr"] != 4000 or c["bps"] != 450:
        ref.set_sampling(c["sr"], c["bps"])
    P = ref.params()  # before ri_mapopt_update: mid_occ still 0, stored separately below
    ref.build_index(fa, "", 4)
    mid_occ = ref.mapopt_update()
    out = {
        "model_seed": np.array(7),
        "preset": np.array(c["preset"]), "r10": np.array(int(c["r10"])), "k": np.array(k),
        "sample_rate": np.array(c["sr"]), "bp_per_sec": np.array(c["bps"]),
        "params": np.frombuffer(bytes(P), dtype=np.uint8).copy(),
        "pore_vals": ref.pore_vals() if k <= 6 else np.zeros(0, np.float32),
        "pore_vals_sha": np.array(sha(ref.pore_vals())),
        "mid_occ": np.array(mid_occ),
        "n_contigs": np.array(len(genome)),
        "names": np.array(names),
        "offset": np.array(synth.OFFSET), "range": np.array(synth.RANGE), "digitisation": np.array(synth.DIGITISATION),
    }
    for i, (nm, s) in enumerate(genome):
        out[f"contig{i}_name"] = np.array(nm)
        out[f"contig{i}_len"] = np.array(len(s))
        out[f"contig{i}_2bit"] = pack2(s)
    sigs = []
    for i, r in enumerate(raws):
        out[f"raw{i}"] = r
        pa = ref.raw_to_pa(r, synth.OFFSET, synth.RANGE, synth.DIGITISATION)
        sigs.append(pa)
        out[f"lsig{i}"] = np.array(len(pa))
        chunks = ref.tap_read(pa, names[i]) if len(pa) else []
        out[f"nchunks{i}"] = np.array(len(chunks))
        for ci, ch in enumerate(chunks):
            pre = f"r{i}c{ci}_"
            out[pre + "cnt"] = ch["cnt"]
            out[pre + "events"] = ch["events"].view(np.uint32)
            out[pre + "seeds"] = ch["seeds"]
            out[pre + "anchors_sha"] = np.array(sha(ch["anchors"]))
            out[pre + "u"] = ch["u"]
            out[pre + "chain_a"] = ch["chain_a"]
            out[pre + "prev_a"] = ch["prev_a"]
            out[pre + "regs"] = ch["regs"]
    paf, _ = ref.map_paf(sigs, names, 1)
    out["paf"] = np.array(_bind.strip_mt(paf))
    # index answers: every 97th seed hash seen in the taps + a few absent keys
    hs = []
    for i in range(len(raws)):
        for ci in range(int(out[f"nchunks{i}"])):
            hs += [int(x) >> 6 for x in out[f"r{i}c{ci}_seeds"][::97, 0]]
    hs = sorted(set(hs))[:64] + [1, 2, 0xFFFFFFFE]
    off, pos = [0], []
    for h in hs:
        p = ref.idx_get(h)
        pos.append(p); off.append(off[-1] + len(p))
    out["idx_hash"] = np.array(hs, dtype=np.uint64)
    out["idx_off"] = np.array(off, dtype=np.uint64)
    out["idx_pos"] = np.concatenate(pos) if pos else np.zeros(0, np.uint64)
    return out


def shared_vectors() -> dict:
    ref = _bind.RefLib()
    grid = np.concatenate([np.linspace(-4, 4, 801), [-2.0, 2.0, -3.0, 3.0, 0.0]]).astype(np.float32)
    q = np.array([ref.dynamic_quantize(float(v)) for v in grid], dtype=np.uint32)
    rng = np.random.Generator(np.random.PCG64(4242))
    sorts = {}
    for j, (n, nkeys, shift) in enumerate([(50, 7, 0), (300, 9, 0), (5000, 40, 0), (5000, 300, 8), (20000, 1000, 24), (3000, 50, 40)]):
        x = (rng.integers(0, nkeys, n).astype(np.uint64) << np.uint64(shift)) | (rng.integers(0, 3, n).astype(np.uint64) << np.uint64(56)) * np.uint64(shift >= 24)
        xy = np.stack([x, np.arange(n, dtype=np.uint64)], axis=1)
        sorts[f"sort{j}_in"] = xy
        sorts[f"sort{j}_out_y"] = ref.radix_sort_128x(xy)[:, 1].astype(np.uint32)
        k64 = (rng.integers(0, nkeys, n).astype(np.uint64) << np.uint64(32)) | rng.integers(0, 4, n).astype(np.uint64)
        sorts[f"sort64_{j}_in"] = k64
        sorts[f"sort64_{j}_out"] = ref.radix_sort_64(k64)
    return {"quant_in": grid, "quant_out": q, "n_sorts": np.array(6), **sorts}


def main():
    assert _bind.have_ref(), "oracle/_ref/libref_tap.so is missing: run `make -C oracle` in the build container first"
    flags = open(os.path.join(ROOT, "oracle", "_ref", "BUILD_FLAGS.txt")).read().strip()
    for name, c in CASES.items():
        d = make_case(name, c)
        d["ref_build_flags"] = np.array(flags)
        p = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(p, **d)
        print(name, os.path.getsize(p) // 1024, "KiB")
    d = shared_vectors()
    d["ref_build_flags"] = np.array(flags)
    p = os.path.join(HERE, "klib_quant_vectors.npz")
    np.savez_compressed(p, **d)
    print("klib_quant_vectors", os.path.getsize(p) // 1024, "KiB")


if __name__ == "__main__":
    main()
