"""CPU: the oracle (oracle/rh_oracle.cpp) against the golden vectors generated from the compiled,
unmodified reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

import _bind
from golden_util import CASES, GoldenCase, compare_tap, sha


@pytest.fixture(scope="module")
def orc_lib(built):
    return _bind.OracleLib


def _open(orc_lib, g):
    o = orc_lib().open(g.preset, g.r10, g.model)
    if g.non_default_sampling():
        o.set_sampling(g.sample_rate, g.bp_per_sec)
    return o


@pytest.mark.parametrize("case", CASES)
def test_params_and_pore(orc_lib, case):
    g = GoldenCase(case)
    o = _open(orc_lib, g)
    assert bytes(o.params()) == g.z["params"].tobytes()
    pv = o.pore_vals()
    assert sha(pv) == str(g.z["pore_vals_sha"])
    if len(g.z["pore_vals"]):
        assert np.array_equal(pv.view(np.uint32), g.z["pore_vals"].view(np.uint32))


@pytest.mark.parametrize("case", CASES)
def test_index_and_mapping(orc_lib, case):
    g = GoldenCase(case)
    o = _open(orc_lib, g)
    o.build_index(g.fasta, "", 4)
    assert o.mapopt_update() == g.mid_occ
    hs, off, pos = g.z["idx_hash"], g.z["idx_off"], g.z["idx_pos"]
    for j, h in enumerate(hs):
        assert np.array_equal(o.idx_get(int(h)), pos[int(off[j]):int(off[j + 1])]), f"hash {int(h):#x}"
    sigs = []
    for i, raw in enumerate(g.raws):
        pa = o.raw_to_pa(raw, *g.cal)
        assert len(pa) == int(g.z[f"lsig{i}"])
        sigs.append(pa)
        got = o.tap_read(pa, g.names[i]) if len(pa) else []
        assert compare_tap(got, g.chunks(i)) == [], f"read {i}"
    paf, _ = o.map_paf(sigs, g.names, 2)
    assert _bind.strip_mt(paf) == g.paf


def test_quantize_and_klib_sorts(orc_lib):
    import os
    from golden_util import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "klib_quant_vectors.npz"))
    o = orc_lib()
    q = np.array([o.dynamic_quantize(float(v)) for v in z["quant_in"]], dtype=np.uint32)
    assert np.array_equal(q, z["quant_out"])
    for j in range(int(z["n_sorts"])):
        out = o.radix_sort_128x(z[f"sort{j}_in"])
        assert np.array_equal(out[:, 1].astype(np.uint32), z[f"sort{j}_out_y"]), f"128x case {j}: tie order differs from klib"
        assert np.array_equal(o.radix_sort_64(z[f"sort64_{j}_in"]), z[f"sort64_{j}_out"])
