#!/usr/bin/env python
"""Human-size sanity run (not a bench line): GRCh38-shaped synthetic genome (24 contigs, lengths scaled by `scale`)
generated on the device, device-resident index build, a small batch mapped on the GPU and compared line by line with
the compiled reference (whose ri_idx_t is filled from the same flattened index), then a timed batch.

    python profiles/tools/human_check.py [scale] [n_parity_reads] [n_timed_reads] [preset]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from rawhash_b200 import api, synth  # noqa: E402
import _bind  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
n_par = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n_timed = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
preset = sys.argv[4] if len(sys.argv) > 4 else "fast"

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
mp = synth.model_path("r9.4")
means, stdv = synth.load_model_pa(mp, 6)
lens = [max(10000, int(l * scale)) for l in synth.GRCH38_LENS]
out = {"genome_bases": int(sum(lens)), "preset": preset}
t0 = time.time()
G = synth.DeviceGenome(synth.GRCH38_NAMES, lens, device=dev, seed=5)
torch.cuda.synchronize()
out["genome_gen_s"] = round(time.time() - t0, 2)
P = api.make_params(preset)
pore = api.load_pore(mp, 6)
t0 = time.time()
idx = api.Index.build_dev(P, pore, G.names, G.codes.data_ptr(), G.lens, 0)
torch.cuda.synchronize()
out["index_build_s"] = round(time.time() - t0, 2)
t0 = time.time()
idx.update_mapopt(P)
out["mid_occ_s"] = round(time.time() - t0, 3)
out.update(index_keys=int(idx.n_keys), index_positions=int(idx.n_pos), mid_occ=int(P.mid_occ))
print(json.dumps(out), file=sys.stderr, flush=True)

n_all = max(n_par, n_timed)
raw_dev, raw_off, rlens, truth = synth.make_reads_torch(G, n_all, 5000, 6, means, stdv, device=dev, seed=9)
torch.cuda.synchronize()
free_b, tot_b = torch.cuda.mem_get_info()
out["free_gb_before_mapper"] = round(free_b / 2**30, 1)
t0 = time.time()
m = api.Mapper(idx, P, 0, int(float(os.environ['RH_ARENA_GB']) * 2**30) if 'RH_ARENA_GB' in os.environ else int(free_b * 0.6))
out["mapper_init_s"] = round(time.time() - t0, 2)
cal = lambda n: (np.full(n, synth.OFFSET), np.full(n, synth.RANGE), np.full(n, synth.DIGITISATION))
names = [f"read_{i:07d}" for i in range(n_all)]

if n_par > 0:
    t0 = time.time()
    recs = m.map_batch_device(raw_dev.data_ptr(), raw_off[:n_par + 1], *cal(n_par), names=None)
    out["parity_gpu_s"] = round(time.time() - t0, 2)
    st = m.stats()
    out["anchors_per_chunk_parity_batch"] = st["n_anchors"] / max(st["n_chunks"], 1)
    got = _bind.strip_mt(idx.format_paf(recs, names[:n_par])).splitlines()
    t0 = time.time()
    keys, off, pos = idx.flat()
    out["index_download_s"] = round(time.time() - t0, 2)
    ref = _bind.RefLib().open(preset, False, mp)
    t0 = time.time()
    ref.index_from_flat(G.names, G.lens, keys, off, pos, os.cpu_count() or 8)
    out["reference_index_fill_s"] = round(time.time() - t0, 2)
    assert ref.mapopt_update() == P.mid_occ, (ref.mapopt_update(), P.mid_occ)
    host = raw_dev[: int(raw_off[n_par])].cpu().numpy()
    sigs = [synth.raw_to_pa(host[int(raw_off[i]):int(raw_off[i + 1])], synth.OFFSET, synth.RANGE, synth.DIGITISATION) for i in range(n_par)]
    exp, secs = ref.map_paf(sigs, names[:n_par], os.cpu_count() or 8)
    exp = _bind.strip_mt(exp).splitlines()
    out["parity_reads"] = n_par
    out["paf_lines_equal"] = sum(1 for a, b in zip(got, exp) if a == b)
    out["paf_identical"] = got == exp
    out["reference_reads_per_s"] = n_par / secs
    out["reference_threads"] = os.cpu_count()
    if got != exp:
        for a, b in zip(got, exp):
            if a != b:
                print("GPU:", a, "\nREF:", b, file=sys.stderr)
                break
    print(json.dumps(out), file=sys.stderr, flush=True)

if n_timed > 0:
    m.map_batch_device(raw_dev.data_ptr(), raw_off[: min(n_timed, 256) + 1], *cal(min(n_timed, 256)), names=None)
    t0 = time.time()
    recs = m.map_batch_device(raw_dev.data_ptr(), raw_off[:n_timed + 1], *cal(n_timed), names=None)
    dt = time.time() - t0
    st = m.stats()
    out["timed_reads"] = n_timed
    out["gpu_reads_per_s"] = n_timed / dt
    out["anchors_per_chunk"] = st["n_anchors"] / max(st["n_chunks"], 1)
    out["chunks_per_read"] = st["n_chunks"] / n_timed
    out["mapped_fraction"] = float((recs["mapped"] == 1).mean())
    tr = np.array(truth[:n_timed])
    first = {}
    for r in recs:
        first.setdefault(int(r["read_idx"]), r)
    ok = sum(1 for i, r in first.items() if r["mapped"] and int(r["ref_id"]) == tr[i][0] and abs(int(r["fragment_start_position"]) - tr[i][1]) < 6000)
    out["true_locus_fraction"] = ok / n_timed
    out["stage_ms"] = {k: round(st[k], 1) for k in ("ms_event_kernel", "ms_seed", "ms_sort", "ms_sort_ties", "ms_chain", "ms_post", "ms_total")}
m.close()
print(json.dumps(out))
