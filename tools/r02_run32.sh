#!/bin/bash
# feature tracks on the GPU: tests + timing at the dataset's scale
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_frontend.py -q -m gpu -x 2>&1 | tail -6
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import pba_b200 as pb, oracle_ffi as of
rng = np.random.default_rng(3)
n_img, n_feat = 164, 1500
counts = [n_feat] * n_img
# a shared pool of 3-D points: image i sees a random 40 % of 6,000 points as its features; matches = shared points
pool = 6000
seen = [rng.choice(pool, n_feat, replace=False) for _ in range(n_img)]
inv = []
for s in seen:
    m = np.full(pool, -1, np.int32); m[s] = np.arange(n_feat); inv.append(m)
pairs, matches = [], []
for a in range(n_img):
    for b in range(a + 1, min(n_img, a + 12)):
        both = np.nonzero((inv[a] >= 0) & (inv[b] >= 0))[0]
        both = both[rng.random(len(both)) < 0.5]
        pairs.append((a, b)); matches.append(np.stack([inv[a][both], inv[b][both]], 1).astype(np.int32))
n_edges = sum(len(m) for m in matches)
pb.build_tracks(counts[:2], [(0, 1)], [matches[0]], 3)
t0 = time.perf_counter(); tg, ng = pb.build_tracks(counts, pairs, matches, 3); tgpu = time.perf_counter() - t0
t0 = time.perf_counter(); to, no = of.build_tracks("oracle", counts, pairs, matches, 3); tor = time.perf_counter() - t0
same = ng == no and all(np.array_equal(a, b) for a, b in zip(tg, to))
line = "tracks: %d images x %d features, %d pairs, %d matches -> %d tracks; GPU (host buffers in/out) %.2f ms, oracle union-find %.2f ms, identical %s" % (n_img, n_feat, len(pairs), n_edges, ng, 1e3 * tgpu, 1e3 * tor, same)
if of.have_ref_frontend():
    t0 = time.perf_counter(); tr, nr = of.build_tracks("ref", counts, pairs, matches, 3); tref = time.perf_counter() - t0
    line += "; reference TrackBuilder %.1f ms, identical %s" % (1e3 * tref, nr == ng and all(np.array_equal(a, b) for a, b in zip(tg, tr)))
print(line)
open('gpurun_out/r02d_tracks_bench.txt', 'w').write(line + "\n")
PY
