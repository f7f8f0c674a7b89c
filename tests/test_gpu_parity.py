"""GPU parity: the CUDA path (through the C-ABI) against the oracle, bit for bit."""
import numpy as np
import pytest

from common import World, tap_equal

pytestmark = pytest.mark.gpu


def _setup(world, preset="sensitive", r10=False, **over):
    from rawhash_b200 import api
    from _bind import OracleLib
    P = api.make_params(preset, r10, **over)
    pore = api.load_pore(world.model, world.k)
    names, seqs = world.genome_strings()
    idx = api.Index.build(P, pore, names, seqs, 8)
    idx.update_mapopt(P)
    orc = OracleLib().open(preset, r10, world.model)
    orc.build_index(world.fasta, "", 4)
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
