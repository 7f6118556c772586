"""Pin the per-case parity outcome against the reference's golden streams -> tests/golden/expected_parity.json.

    python tests/golden/pin_expected_parity.py --oracle          # CPU oracle part (runs anywhere)
    python tests/golden/pin_expected_parity.py --gpu [--out F]   # CUDA-path part (needs a B200); merges into the same file

For every golden case (tests/golden/golden*.npz, produced by the real reference) it records whether the stream produced
by the implementation under test is byte-identical to the reference's and, per layer, how many edge pixels / quantised
coefficients differ.  The tests then require every NAMED case to be no worse than what is pinned here."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in ("tests", "oracle", "adaptive-edge-aware-jpeg_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))

import oracle as O  # noqa: E402
import parity_report as PR  # noqa: E402
from conftest import Golden  # noqa: E402


def ref_edges(g, name, shapes):
    return [np.unpackbits(g.get(name, f"edge{i}"))[: shapes[i][0] * shapes[i][1]].reshape(shapes[i]) for i in range(3)]


def oracle_case(g, name):
    c = g.case(name)
    rgb = g.input_f32(name)
    H, W, _ = rgb.shape
    ref_bytes = g.get(name, "ajpg").tobytes()
    _, ref = PR.parse_ajpg(ref_bytes)
    q, b = tuple(c["quality"]), tuple(c["blocks"])
    got = O.encode_hot(rgb, c["space"], q, b)
    edges = ref_edges(g, name, O.layer_shapes(H, W, c["space"]))
    layers = [PR.layer_report(ref[i], edges[i], got[i]["states"], got[i]["root"], O.zigzag_stream(got[i]["coef"], got[i]["leaves"]), got[i]["edge"])
              for i in range(3)]
    return {"byte_identical": O.compress(rgb, c["space"], q, b, ".png") == ref_bytes, "layers": layers}


def gpu_case(g, name, codec=None):
    """the CUDA path through the reference-facing shim (Jpeg.compress) + the fused encoder's edge taps"""
    import torch
    from aeaj.codec import get_codec
    from image import Image
    from jpeg import Jpeg, JpegCompressionSettings
    codec = codec or get_codec(0)
    c = g.case(name)
    rgb = g.input_f32(name)
    H, W, _ = rgb.shape
    ref_bytes = g.get(name, "ajpg").tobytes()
    _, ref = PR.parse_ajpg(ref_bytes)
    q, b = tuple(c["quality"]), tuple(c["blocks"])
    mine = Jpeg(JpegCompressionSettings(c["space"], q, b)).compress(Image.from_array(rgb.copy(), None, ".png"))
    _, got = PR.parse_ajpg(mine)
    enc = codec.encode(torch.from_numpy(rgb).cuda(), c["space"], q, b, taps=True)
    got_edges = [e[0].cpu().numpy() for e in enc.edges]
    edges = ref_edges(g, name, O.layer_shapes(H, W, c["space"]))
    layers = [PR.layer_report(ref[i], edges[i], got[i]["states"], got[i]["root"], got[i]["coef"], got_edges[i]) for i in range(3)]
    return {"byte_identical": mine == ref_bytes, "layers": layers}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--out", default=PR.EXPECTED_PATH)
    a = ap.parse_args()
    g = Golden()
    exp = PR.load_expected()
    exp["_doc"] = ("per named golden case: parity of the CPU oracle ('oracle') and of the CUDA path ('gpu') against the reference's "
                   ".ajpg stream; written by tests/golden/pin_expected_parity.py, enforced by test_oracle_golden.py / test_gpu_parity.py")
    names = sorted(n for n, c in g.meta["cases"].items() if c.get("mode") != "D")
    for kind, fn in (("oracle", oracle_case), ("gpu", gpu_case)):
        if not getattr(a, kind):
            continue
        exp[kind] = {}
        for n in names:
            exp[kind][n] = fn(g, n)
            r = exp[kind][n]
            print(kind, n, "identical" if r["byte_identical"] else "differs", [(l["edge_px"], l["tree_equal"], l["coef_diffs"]) for l in r["layers"]], flush=True)
        print(kind, sum(r["byte_identical"] for r in exp[kind].values()), "of", len(names), "streams byte-identical to the reference's")
    with open(a.out, "w") as f:
        json.dump(exp, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
