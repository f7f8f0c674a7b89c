#!/usr/bin/env python
"""Larger-genome sanity run (not a bench line): 100 Mb synthetic genome, GPU index build, a small batch mapped
on the GPU and compared with the oracle, then a timed batch.  Exercises the paths the 12 Mb workload does not:
chunks of >100 k anchors (global-memory sort, tie replay beyond the packed 16-bit tables), long DP segments.

    python profiles/tools/scale_check.py [genome_mb] [n_parity_reads] [n_timed_reads]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from rawhash_b200 import api, synth  # noqa: E402
import _bind  # noqa: E402

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_par = int(sys.argv[2]) if len(sys.argv) > 2 else 48
n_timed = int(sys.argv[3]) if len(sys.argv) > 3 else 2000

mp = synth.model_path("r9.4")
means, stdv = synth.load_model_pa(mp, 6)
t0 = time.time()
genome = synth.make_genome(8, mb * 1_000_000, seed=5)
gs = synth.genome_to_strings(genome)
names, seqs = [n for n, _ in gs], [s for _, s in gs]
P = api.make_params("sensitive")
pore = api.load_pore(mp, 6)
t1 = time.time()
idx = api.Index.build_gpu(P, pore, names, seqs, 0)
t_idx = time.time() - t1
idx.update_mapopt(P)
out = {"genome_mb": mb, "index_keys": int(idx.n_keys), "index_positions": int(idx.n_pos), "index_build_gpu_s": round(t_idx, 2), "mid_occ": int(P.mid_occ)}

rd = synth.make_reads(genome, max(n_par, n_timed), 5000, 6, means, stdv, seed=9)
cal = lambda n: (np.full(n, synth.OFFSET), np.full(n, synth.RANGE), np.full(n, synth.DIGITISATION))
m = api.Mapper(idx, P, 0, 0)
if n_par > 0:
    recs = m.map_batch(rd["raw"][:n_par], *cal(n_par), rd["names"][:n_par])
    got = _bind.strip_mt(idx.format_paf(recs, rd["names"][:n_par])).splitlines()
    st = m.stats()
    out["anchors_per_chunk_parity_batch"] = st["n_anchors"] / max(st["n_chunks"], 1)
    fa = "/tmp/rh_scale_check.fa"
    synth.write_fasta(fa, genome)
    orc = _bind.OracleLib().open("sensitive", False, mp)
    t1 = time.time()
    orc.build_index(fa, "", os.cpu_count() or 8)
    out["index_build_oracle_s"] = round(time.time() - t1, 2)
    assert orc.mapopt_update() == P.mid_occ
    exp, secs = orc.map_paf([synth.raw_to_pa(r, synth.OFFSET, synth.RANGE, synth.DIGITISATION) for r in rd["raw"][:n_par]], rd["names"][:n_par], os.cpu_count() or 8)
    exp = _bind.strip_mt(exp).splitlines()
    out["parity_reads"] = n_par
    out["paf_lines_equal"] = sum(1 for a, b in zip(got, exp) if a == b)
    out["paf_identical"] = got == exp
    out["oracle_reads_per_s"] = n_par / secs

t1 = time.time()
recs = m.map_batch(rd["raw"][:n_timed], *cal(n_timed), rd["names"][:n_timed])
dt = time.time() - t1
st = m.stats()
out["timed_reads"] = n_timed
out["gpu_reads_per_s_host_buffers"] = n_timed / dt
out["anchors_per_chunk"] = st["n_anchors"] / max(st["n_chunks"], 1)
out["chunks_per_read"] = st["n_chunks"] / n_timed
out["mapped_fraction"] = float((recs["mapped"] == 1).mean())
out["stage_ms"] = {k: st[k] for k in ("ms_event_kernel", "ms_seed", "ms_sort", "ms_sort_ties", "ms_chain", "ms_post", "ms_total")}
m.close()
print(json.dumps(out))
