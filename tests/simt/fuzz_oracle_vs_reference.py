"""TEST INFRASTRUCTURE ONLY — the same random scenes as fuzz.py, but oracle/axr_oracle.c against the UNMODIFIED reference
(oracle/_ref, needs /root/reference at build time): strengthens the pin of the oracle beyond the golden fixtures. Nearest sampling
only (the reference has no other). usage: python tests/simt/fuzz_oracle_vs_reference.py [--seconds 60] [--seed 0]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import pyoracle as po  # noqa: E402
from fuzz import random_scene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    if not po.ref_available():
        print("reference build not available")
        return
    rng = np.random.default_rng(a.seed)
    t_end = time.time() + a.seconds
    n = covered = composites = 0
    prev = {}
    while time.time() < t_end:
        sc = random_scene(rng, n)
        sc.sampler = 0
        onto = prev.get((sc.height, sc.width)) if rng.random() < 0.5 else None
        kw = {} if onto is None else {"color": onto[0], "depth": onto[1]}
        c0, d0, _ = po.oracle_render(sc, threads=2, **kw)
        c1, d1, _ = po.ref_render(sc, threads=int(rng.choice([1, 2, 4])), **kw)
        m = po.compare(c1, d1, c0, d0)
        if m["coverage_mismatch"] or m["depth_bit_mismatch"] or m["color_max_diff"]:
            print(f"MISMATCH seed={a.seed} scene #{n} {sc.name} {sc.width}x{sc.height} composite={onto is not None} faces={sc.n_faces}: {m}", flush=True)
            sys.exit(1)
        prev[(sc.height, sc.width)] = (c0, d0)
        n += 1
        covered += m["covered"]
        composites += onto is not None
    print(f"FUZZ OK seed={a.seed}: {n} scenes ({composites} composites), {covered} covered pixels, oracle == unmodified reference bit for bit", flush=True)


if __name__ == "__main__":
    main()
