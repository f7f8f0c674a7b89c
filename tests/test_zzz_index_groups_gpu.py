"""GPU: the device-resident index builder (rh_index_build_dev) on a reference of tens of Mb — many filter segments per
strand, so the speculative incoming states and their fix-up rounds are exercised — against the host builder, through
the byte-exact `.ind` writer; and the device-side mid_occ (radix select over the CSR offsets) against the host one."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("preset,r10", [("sensitive", False), ("fast", False), ("fast", True)])
def test_device_index_equals_host_build(built, tmp_path, preset, r10):
    import torch
    from rawhash_b200 import api, synth
    from common import model_for
    kind, k = ("r10.4.1", 9) if r10 else ("r9.4", 6)
    lens = [5_000_017, 3_000_001, 1_999_999, 777, 8, 4, 2_500_000]   # includes contigs shorter than k and than one window
    G = synth.DeviceGenome([f"c{i}" for i in range(len(lens))], lens, device=torch.device("cuda", 0), seed=13)
    P = api.make_params(preset, r10)
    pore = api.load_pore(model_for(kind), k)
    dev = api.Index.build_dev(P, pore, G.names, G.codes.data_ptr(), G.lens, 0)
    assert dev.on_device == 0
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    hc = G.host_contigs()
    host = api.Index.build(P, pore, [n for n, _ in hc], [lut[c].tobytes() for _, c in hc], 16)
    P2 = api.make_params(preset, r10)
    assert dev.update_mapopt(P) == host.update_mapopt(P2)          # device radix select vs host nth_element
    assert (dev.n_keys, dev.n_pos) == (host.n_keys, host.n_pos)
    a, b = str(tmp_path / "host.ind"), str(tmp_path / "dev.ind")
    host.dump(a, pore)
    dev.dump(b, pore)
    assert open(a, "rb").read() == open(b, "rb").read()
