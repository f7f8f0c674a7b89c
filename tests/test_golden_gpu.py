"""GPU: the CUDA path (through the C-ABI) against the golden vectors generated from the compiled
reference — every stage of every chunk bit for bit, and the PAF text."""
import numpy as np
import pytest

from golden_util import CASES, GoldenCase, compare_tap

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES)
def test_gpu_matches_reference_golden(built, case):
    from rawhash_b200 import api
    from _bind import strip_mt
    g = GoldenCase(case)
    over = dict(sample_rate=g.sample_rate, bp_per_sec=g.bp_per_sec) if g.non_default_sampling() else {}
    P = api.make_params(g.preset, g.r10, **over)
    names, seqs = g.genome_strings()
    idx = api.Index.build(P, api.load_pore(g.model, g.k), names, seqs, 4)
    assert idx.update_mapopt(P) == g.mid_occ
    m = api.Mapper(idx, P, 0, 1 << 30)
    for i, raw in enumerate(g.raws):
        if int(g.z[f"lsig{i}"]) == 0:
            continue
        got = m.tap_read(raw, *g.cal, g.names[i])
        assert compare_tap(got, g.chunks(i)) == [], f"read {i}"
    n = len(g.raws)
    recs = m.map_batch(g.raws, np.full(n, g.cal[0]), np.full(n, g.cal[1]), np.full(n, g.cal[2]), g.names)
    st = m.stats()
    m.close()
    assert strip_mt(idx.format_paf(recs, g.names)) == g.paf
    assert st["kernel_launches"] > 0
