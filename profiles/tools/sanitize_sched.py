#!/usr/bin/env python
"""Small world for compute-sanitizer: a 20 Mb genome (chunks of ~3·10^4 anchors: global-memory sort, long-sub-array replay),
`-x fast`, 240 reads through the streaming scheduler under tight limits (cohorts of 48, 96 reads in flight, 160 MB arena: heavy
lane, deferred admissions, several groups per iteration), then the same batch through the chunk-round loop; the two record
sets must be equal.

    compute-sanitizer --tool memcheck --error-exitcode 1 python profiles/tools/sanitize_sched.py [reads] [genome Mb] [arena MB]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rawhash_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 240
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 20          # genome size: 400 gives chunks of ~7·10^4 anchors (sub-arrays >= 4096: cta_big_*)
arena = int(sys.argv[3]) << 20 if len(sys.argv) > 3 else 160 << 20
dev = torch.device("cuda", 0)
mp = synth.model_path("r9.4")
means, stdv = synth.load_model_pa(mp, 6)
G = synth.DeviceGenome([f"c{i}" for i in range(4)], [mb * 250_000] * 4, device=dev, seed=3)
P = api.make_params("fast")
idx = api.Index.build_dev(P, api.load_pore(mp, 6), G.names, G.codes.data_ptr(), G.lens, 0)
idx.update_mapopt(P)
raw_dev, raw_off, lens, truth = synth.make_reads_torch(G, n, 5000, 6, means, stdv, device=dev, seed=5)
cal = (np.full(n, synth.OFFSET), np.full(n, synth.RANGE), np.full(n, synth.DIGITISATION))
m = api.Mapper(idx, P, 0, arena)
os.environ.update(RH_SCHED_STREAM="1", RH_MAX_INFLIGHT="96", RH_COHORT="48")
a = m.map_batch_device(raw_dev.data_ptr(), raw_off, *cal, names=None)
st = m.stats()
os.environ.pop("RH_SCHED_STREAM"); os.environ["RH_SCHED_ROUNDS"] = "1"
b = m.map_batch_device(raw_dev.data_ptr(), raw_off, *cal, names=None)
m.close()
f = [x for x in a.dtype.names if x != "mt_ms"]
same = bool(np.array_equal(a[f], b[f]))
print(f"{n} reads, {st['n_chunks']} chunks, {st['n_anchors'] / max(st['n_chunks'], 1):.0f} anchors per chunk, {st['n_rounds']} iterations, "
      f"{int((a['mapped'] == 1).sum())} mapped records; streaming == chunk rounds: {same}")
sys.exit(0 if same else 2)
