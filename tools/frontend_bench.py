"""Front-end throughput: match_all of the reference's SfM pipeline (src/sfm.cpp:1286-1330: every
image pair of the dataset through matchDescriptors, keypoints.h:282-300) at the bundled dataset's scale —
164 images x 1,500 descriptors (src/sfm.cpp:197), 13,366 pairs — through pba_match_descriptors (host buffers in,
host buffers out), against the reference's own matchDescriptors timed on a bounded sample of the same pairs
(all host threads, one pair per thread, as the reference's tbb::parallel_for does).  Prints one JSON line.

    python tools/frontend_bench.py [--images 164] [--features 1500] [--ref-pairs 32]
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pba_b200 as pb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=164)
    ap.add_argument("--features", type=int, default=1500)
    ap.add_argument("--ref-pairs", type=int, default=32)
    ap.add_argument("--repeat", type=int, default=3)
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    # descriptors with structure: every image sees noisy copies of a shared pool, so real matches exist
    pool = rng.integers(0, 256, (4 * a.features, 32), dtype=np.uint8)
    sets = []
    for i in range(a.images):
        d = pool[rng.choice(len(pool), a.features, replace=False)].copy()
        flips = rng.integers(0, 256, (a.features, 24))
        for k in range(24):
            d[np.arange(a.features), flips[:, k] // 8] ^= (1 << (flips[:, k] % 8)).astype(np.uint8)
        sets.append(d)
    pairs = np.array([(i, j) for i in range(a.images) for j in range(i + 1, a.images)], np.int32)
    pb.match_descriptors(sets[:2], [(0, 1)])  # context + module load
    best = None
    for _ in range(a.repeat):
        t0 = time.perf_counter()
        ms = pb.match_descriptors(sets, pairs, 70, 1.2)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    n_matches = int(sum(len(m) for m in ms))
    comparisons = 2.0 * len(pairs) * a.features * a.features
    line = {"metric": "descriptor_pairs_matched_per_s", "workload": "match_all: %d images x %d descriptors, %d image pairs"
            % (a.images, a.features, len(pairs)), "value": len(pairs) / best, "unit": "image pairs/s",
            "wall_s": best, "hamming_distances_per_s": comparisons / best, "matches": n_matches,
            "call": "pba_match_descriptors on host buffers (H2D of the descriptors, D2H of the matches inside the timed region)"}
    try:
        import oracle_ffi as of
        if of.have_ref_frontend():
            sample = pairs[rng.choice(len(pairs), min(a.ref_pairs, len(pairs)), replace=False)]
            threads = os.cpu_count() or 1

            def one(p):
                return of.match_descriptors("ref", sets[p[0]], sets[p[1]], 70, 1.2)
            of.match_descriptors("ref", sets[0][:8], sets[1][:8], 70, 1.2)
            t0 = time.perf_counter()
            with ThreadPoolExecutor(threads) as ex:
                ref = list(ex.map(one, sample))
            dt = time.perf_counter() - t0
            same = all(np.array_equal(ref[k], ms[int(np.nonzero((pairs == sample[k]).all(1))[0][0])]) for k in range(len(sample)))
            line["cpu_baseline"] = {"value": len(sample) / dt, "unit": "image pairs/s", "cores": threads, "kind": "reference",
                                    "sample": "%d of the %d pairs through the reference's matchDescriptors (ctypes releases the GIL)"
                                    % (len(sample), len(pairs)), "identical_matches": bool(same)}
    except ImportError:
        pass
    print(json.dumps(line))


if __name__ == "__main__":
    main()
