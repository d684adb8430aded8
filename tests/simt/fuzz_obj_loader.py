"""TEST INFRASTRUCTURE ONLY — differential fuzzing of the OBJ ingestion mirror (axiomr_b200/obj.py) against the reference's own loader
(AR::Mesh(path) through oracle/_ref): random OBJ text with the irregularities real files have — missing vt / vn, polygons, indices
out of range or zero, forward references, duplicate vertices, several usemtl groups, blank lines, comments, leading blanks, tabs,
CRLF line ends, numbers in several spellings. Vertices (incl. generated tangents / bitangents), faces and order must match bit for bit.
With AXR_B200_LIB set (e.g. the SIMT interpreter build) the tangents come from the CUDA kernels (axr_generate_tangents).
usage: python tests/simt/fuzz_obj_loader.py [--seconds 60] [--seed 0]"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from axiomr_b200 import obj  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def fnum(rng, x):
    k = rng.integers(0, 6)
    if k == 0:
        return f"{x:.6f}"
    if k == 1:
        return f"{x:.9g}"
    if k == 2:
        return f"{x:.3e}"
    if k == 3:
        return repr(float(np.float32(x)))
    if k == 4:
        return f"{x:.0f}"
    return f"{x:+.4f}" if rng.random() < 0.5 else f"{x:.17g}"


def random_obj(rng) -> str:
    nv, nt, nn = int(rng.integers(1, 40)), int(rng.integers(0, 20)), int(rng.integers(0, 20))
    eol = "\r\n" if rng.random() < 0.2 else "\n"
    lines = []
    pool = rng.normal(0, 2, (8, 3))  # a few repeated coordinates -> value-equal vertices from different `v` lines
    decl = []
    for i in range(nv):
        p = pool[rng.integers(0, 8)] if rng.random() < 0.3 else rng.normal(0, 2, 3)
        decl.append("v " + " ".join(fnum(rng, c) for c in p) + (" 1.0" if rng.random() < 0.1 else ""))
    for i in range(nt):
        uv = rng.uniform(-0.5, 1.5, 3)
        decl.append("vt " + " ".join(fnum(rng, c) for c in uv[: (3 if rng.random() < 0.2 else 2)]))
    for i in range(nn):
        n = rng.normal(0, 1, 3)
        n /= np.linalg.norm(n) + 1e-9
        decl.append("vn " + " ".join(fnum(rng, c) for c in n))
    faces = []
    for i in range(int(rng.integers(0, 60))):
        if rng.random() < 0.12:
            faces.append(f"usemtl m{int(rng.integers(0, 3))}")
        k = int(rng.choice([3, 3, 3, 4, 5, 2, 1]))
        toks = []
        for _ in range(k):
            vi = int(rng.integers(-1, nv + 3))  # 0, negative and > nv occur
            form = rng.integers(0, 4)
            ti = int(rng.integers(0, nt + 2)) if nt else 1
            ni = int(rng.integers(0, nn + 2)) if nn else 1
            toks.append([f"{vi}", f"{vi}/{ti}", f"{vi}//{ni}", f"{vi}/{ti}/{ni}"][form])
        faces.append(("f " if rng.random() < 0.9 else "f\t") + " ".join(toks))
    body = decl + faces
    if rng.random() < 0.5:  # interleave declarations and faces: forward references become invalid indices
        rng.shuffle(body)
    for ln in body:
        if rng.random() < 0.05:
            lines.append("")
        if rng.random() < 0.05:
            lines.append("# comment " + ln)
        if rng.random() < 0.05:
            lines.append("o thing" if rng.random() < 0.5 else "s off")
        lines.append(("  " if rng.random() < 0.05 else "") + ln + ("  " if rng.random() < 0.05 else ""))
    return eol.join(lines) + (eol if rng.random() < 0.8 else "")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--native", action="store_true", help="check axr_load_obj_file (the C++ loader behind the C ABI) instead of axiomr_b200/obj.py; needs AXR_B200_LIB")
    a = ap.parse_args()
    if not po.ref_available():
        print("reference build not available")
        return
    rng = np.random.default_rng(a.seed)
    d = tempfile.mkdtemp(prefix="axr_objfuzz_")
    path = os.path.join(d, "m.obj")
    with open(os.path.join(d, "m.mtl"), "w") as f:
        f.write("newmtl m0\nNs 0.25\nnewmtl m1\nNs 0.5\nnewmtl m2\nNs 1\n")
    dev = None
    if os.environ.get("AXR_B200_LIB"):  # tangents / bitangents by the CUDA kernels (axr_generate_tangents) instead of obj.tangents
        from axiomr_b200 import api
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import use_simt
        use_simt.install()
        dev = api.Device(16, 16)
    t_end = time.time() + a.seconds
    n = nfaces = 0
    while time.time() < t_end:
        text = random_obj(rng)
        with open(path, "w", newline="") as f:
            f.write(text)
        try:
            want_v, want_f = po.ref_load_obj(path)
        except Exception as e:  # the reference throws on some inputs (std::stoi on junk); those are not parity cases
            continue
        if a.native:
            try:
                mh, got_v, got_f, _ = dev.load_obj(path)
            except api.AxrError as e:  # the reference accepted the file: the native loader has to as well
                print(f"MISMATCH seed={a.seed} file #{n}: axr_load_obj_file refused what the reference loaded: {e}", flush=True)
                sys.exit(1)
            dev.free_mesh(mh)
        else:
            m = obj.load_obj(path, texture_loader=lambda p: None, device=dev)
            got_v, got_f = m.getVertices(), m.getFaces()
        ok = got_v.shape == want_v.shape and got_f.shape == want_f.shape and np.array_equal(got_f, want_f)
        if ok:
            same = (got_v.view(np.uint32) == want_v.view(np.uint32)) | (np.isnan(got_v) & np.isnan(want_v))
            ok = bool(same.all())
        if not ok:
            keep = os.path.join(tempfile.gettempdir(), f"axr_objfuzz_fail_{a.seed}_{n}.obj")
            with open(keep, "w", newline="") as f:
                f.write(text)
            print(f"MISMATCH seed={a.seed} file #{n}: shapes {got_v.shape} {want_v.shape} {got_f.shape} {want_f.shape}; kept as {keep}", flush=True)
            sys.exit(1)
        n += 1
        nfaces += int(want_f.shape[0])
    print(f"FUZZ OK seed={a.seed}: {n} OBJ files, {nfaces} faces, {'axr_load_obj_file' if a.native else 'obj.py'}{' + CUDA tangent kernels' if dev else ''} == reference loader bit for bit", flush=True)


if __name__ == "__main__":
    main()
