"""GPU: references larger than one device pass are indexed contig group by contig group (rh_index_build_grouped);
forced here with a small RH_INDEX_GROUP_BASES; the CPU twin through the host builder is
tests/test_io.py::test_index_built_over_contig_groups_equals_one_pass."""
import pytest

from common import World

pytestmark = pytest.mark.gpu


def test_gpu_index_over_contig_groups_equals_host_build(built, tmp_path, monkeypatch):
    from rawhash_b200 import api
    w = World(n_contigs=5, genome_len=1_200_000, n_reads=1, read_bp=1000, seed=51)
    P = api.make_params("sensitive")
    pore = api.load_pore(w.model, w.k)
    names, seqs = w.genome_strings()
    monkeypatch.delenv("RH_INDEX_GROUP_BASES", raising=False)
    host = api.Index.build(P, pore, names, seqs, 8)
    monkeypatch.setenv("RH_INDEX_GROUP_BASES", "500000")   # contigs of 240 kb: groups of two, two, one
    dev = api.Index.build_gpu(P, pore, names, seqs, 0)
    a, b = str(tmp_path / "host.ind"), str(tmp_path / "dev.ind")
    host.dump(a, pore)
    dev.dump(b, pore)
    assert open(a, "rb").read() == open(b, "rb").read()
