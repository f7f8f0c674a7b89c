"""GPU: PAF parity against the compiled reference at the sizes of BASELINE.json's configs — the paths small worlds never
reach: chunks of 10^5..10^6 anchors (global-memory anchor sort, klib tie replay through cta_big_*, exact score sort of
every anchor with `-x fast`), several anchor-arena groups per round, carry-arena regrowth, read ranges split on
device-memory exhaustion.

The genome is generated in device memory and indexed there (rh_index_build_dev); the reference maps the same reads with
its ri_idx_t filled from the same flattened table (ref_index_from_flat, checked against ri_idx_gen in
tests/test_oracle_vs_ref.py).  Skipped where oracle/_ref/libref_tap.so is absent."""
import os

import numpy as np
import pytest

import _bind

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not _bind.have_ref(), reason="oracle/_ref/libref_tap.so not built")]


def _world(names, lens, preset, r10=False, seed=3):
    import torch
    from rawhash_b200 import api, synth
    from common import model_for
    kind, k = ("r10.4.1", 9) if r10 else ("r9.4", 6)
    model = model_for(kind)
    means, stdv = synth.load_model_pa(model, k)
    dev = torch.device("cuda", 0)
    G = synth.DeviceGenome(names, lens, device=dev, seed=seed)
    over = dict(sample_rate=5000, bp_per_sec=400) if r10 else {}
    P = api.make_params(preset, r10, **over)
    idx = api.Index.build_dev(P, api.load_pore(model, k), G.names, G.codes.data_ptr(), G.lens, 0)
    idx.update_mapopt(P)
    return dict(api=api, synth=synth, G=G, P=P, idx=idx, model=model, means=means, stdv=stdv, k=k, preset=preset, r10=r10, dev=dev)


def _reads(W, n, seed):
    synth = W["synth"]
    rate, bps = (5000.0, 400.0) if W["r10"] else (4000.0, 450.0)
    raw_dev, raw_off, lens, truth = synth.make_reads_torch(W["G"], n, 5000, W["k"], W["means"], W["stdv"], device=W["dev"],
                                                          sample_rate=rate, bp_per_sec=bps, seed=seed)
    return raw_dev, raw_off


def _reference_paf(W, raw_dev, raw_off, n, names):
    synth = W["synth"]
    ref = _bind.RefLib().open(W["preset"], W["r10"], W["model"])
    if W["r10"]:
        ref.set_sampling(5000, 400)
    keys, off, pos = W["idx"].flat()
    ref.index_from_flat(W["G"].names, W["G"].lens, keys, off, pos, os.cpu_count() or 8)
    assert ref.mapopt_update() == W["P"].mid_occ
    host = raw_dev[: int(raw_off[n])].cpu().numpy()
    sigs = [synth.raw_to_pa(host[int(raw_off[i]):int(raw_off[i + 1])], synth.OFFSET, synth.RANGE, synth.DIGITISATION) for i in range(n)]
    paf, _ = ref.map_paf(sigs, names, os.cpu_count() or 8)
    return _bind.strip_mt(paf).splitlines()


def _gpu_paf(W, mapper, raw_dev, raw_off, n, names):
    synth = W["synth"]
    cal = (np.full(n, synth.OFFSET), np.full(n, synth.RANGE), np.full(n, synth.DIGITISATION))
    recs = mapper.map_batch_device(raw_dev.data_ptr(), raw_off[: n + 1], *cal, names=None)
    return _bind.strip_mt(W["idx"].format_paf(recs, names)).splitlines(), mapper.stats()


def _assert_same(got, exp):
    assert len(got) == len(exp)
    bad = [(a, b) for a, b in zip(got, exp) if a != b]
    assert not bad, f"{len(bad)} of {len(exp)} PAF lines differ, first:\nGPU: {bad[0][0]}\nREF: {bad[0][1]}"


def test_yeast_size_20k_reads(built):
    """BASELINE configs[1]: 12 Mb / 16 contigs, -x sensitive, 20 000 reads — every line of the PAF against the reference."""
    W = _world([f"chr{i + 1}" for i in range(16)], [750_000] * 16, "sensitive")
    n = 20_000
    raw_dev, raw_off = _reads(W, n, 11)
    names = [f"read_{i:07d}" for i in range(n)]
    m = W["api"].Mapper(W["idx"], W["P"], 0, 8 << 30)
    got, st = _gpu_paf(W, m, raw_dev, raw_off, n, names)
    m.close()
    _assert_same(got, _reference_paf(W, raw_dev, raw_off, n, names))
    assert st["n_chunks"] > n


@pytest.mark.parametrize("preset,r10", [("sensitive", False), ("fast", False), ("fast", True)])
def test_150mb_600_reads_small_arena(built, preset, r10, monkeypatch):
    """150 Mb, chunks of 10^4..10^5 anchors in an arena that holds a few dozen of them: global-memory sort, tie replay and
    score sort through the long-sub-array path, many arena groups per round, carry arenas that have to grow.  By its own
    estimate the library would map a batch this small in plain chunk rounds; the `fast` runs force the streaming scheduler."""
    if preset == "fast":
        monkeypatch.setenv("RH_SCHED_STREAM", "1")
    lens = [40_000_000, 35_000_000, 30_000_000, 25_000_000, 15_000_000, 5_000_000]
    W = _world([f"c{i}" for i in range(len(lens))], lens, preset, r10, seed=7)
    n = 600
    raw_dev, raw_off = _reads(W, n, 21)
    names = [f"read_{i:07d}" for i in range(n)]
    m = W["api"].Mapper(W["idx"], W["P"], 0, 1 << 30)
    got, st = _gpu_paf(W, m, raw_dev, raw_off, n, names)
    m.close()
    _assert_same(got, _reference_paf(W, raw_dev, raw_off, n, names))


def test_human_size_200_reads(built):
    """BASELINE configs[2]: GRCh38-shaped 3.09 Gb genome, -x fast: 200 reads, every PAF line against the reference
    (its index: the same 5.1 G positions, ~80 GB of host memory for the two copies)."""
    import psutil
    if psutil.virtual_memory().available < 110 * 2**30:
        pytest.skip("needs ~100 GB of host memory for the reference's copy of the human-size index")
    from rawhash_b200 import synth
    W = _world(synth.GRCH38_NAMES, synth.GRCH38_LENS, "fast", seed=1)
    assert W["idx"].n_pos > 4_000_000_000
    n = 200
    raw_dev, raw_off = _reads(W, n, 31)
    names = [f"read_{i:07d}" for i in range(n)]
    import torch
    free_b, _ = torch.cuda.mem_get_info()
    m = W["api"].Mapper(W["idx"], W["P"], 0, int(free_b * 0.5))
    got, st = _gpu_paf(W, m, raw_dev, raw_off, n, names)
    m.close()
    assert st["n_anchors"] / max(st["n_chunks"], 1) > 200_000
    _assert_same(got, _reference_paf(W, raw_dev, raw_off, n, names))


def test_streaming_scheduler_limits(built, monkeypatch):
    """The streaming scheduler under tight limits: 128-read cohorts, at most 300 reads in flight and a 96 MB arena, so that
    waiting reads are deferred again and again, the in-flight cap binds, every iteration runs several arena groups and the
    heavy lane has work; the same batch through the plain chunk-round loop (RH_SCHED_ROUNDS) and through the reference."""
    W = _world([f"chr{i + 1}" for i in range(8)], [750_000] * 8, "fast", seed=13)
    n = 3000
    raw_dev, raw_off = _reads(W, n, 17)
    names = [f"read_{i:07d}" for i in range(n)]
    exp = _reference_paf(W, raw_dev, raw_off, n, names)
    monkeypatch.setenv("RH_MAX_INFLIGHT", "300")
    monkeypatch.setenv("RH_COHORT", "128")
    m = W["api"].Mapper(W["idx"], W["P"], 0, 96 << 20)
    got, st = _gpu_paf(W, m, raw_dev, raw_off, n, names)
    assert st["n_rounds"] > 12   # iterations: far more than the chunk rounds of one read
    _assert_same(got, exp)
    monkeypatch.setenv("RH_SCHED_ROUNDS", "1")
    got2, st2 = _gpu_paf(W, m, raw_dev, raw_off, n, names)
    m.close()
    assert st2["n_rounds"] <= W["P"].max_num_chunk
    _assert_same(got2, exp)
