"""Per-case, per-layer parity accounting against the reference's golden .ajpg streams (test infrastructure).

`layer_report` compares one layer of an implementation under test (the CPU oracle or the CUDA path) with the same layer of
the reference's stream: edge-map mismatches (pixels), quadtree equality (root + state stream), quantised-coefficient
differences (count, max |d|).  tests/golden/expected_parity.json pins the outcome PER NAMED CASE: which streams are
byte-identical to the reference's and, for the others, how many coefficients / edge pixels differ (tie classes T-DCT /
T-POW / T-BIL of DESIGN.md) -- so a regression in any single case fails, not just a drop of the overall rate."""
from __future__ import annotations

import io
import json
import os
import zlib

import numpy as np

EXPECTED_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "expected_parity.json")


def parse_ajpg(data: bytes):
    """-> (meta, [dict(states uint8, root int, coef int32 zigzag-ordered stream)])   (jpeg.py:609-672 container)"""
    s = io.BytesIO(data)
    meta = json.loads(s.read(int.from_bytes(s.read(4), "big")).decode())
    layers = []
    for _ in range(meta["num_layers"]):
        nbits = int.from_bytes(s.read(4), "big")
        root = int.from_bytes(s.read(4), "big")
        raw = np.frombuffer(s.read((nbits + 7) // 8), dtype=np.uint8)
        states = np.stack([(raw >> 6) & 3, (raw >> 4) & 3, (raw >> 2) & 3, raw & 3], axis=1).reshape(-1)[: nbits // 2]
        coef = np.frombuffer(zlib.decompress(s.read(int.from_bytes(s.read(4), "big"))), dtype=np.int32)
        layers.append(dict(states=states.astype(np.uint8), root=root, coef=coef))
    return meta, layers


def layer_report(ref_layer: dict, ref_edge, got_states, got_root, got_coef_zz, got_edge) -> dict:
    rep = {"edge_px": None if ref_edge is None or got_edge is None else int((np.asarray(got_edge) != np.asarray(ref_edge)).sum())}
    rep["tree_equal"] = bool(got_root == ref_layer["root"] and np.array_equal(got_states, ref_layer["states"]))
    if rep["tree_equal"] and len(got_coef_zz) == len(ref_layer["coef"]):
        d = np.abs(np.asarray(got_coef_zz, dtype=np.int64) - ref_layer["coef"].astype(np.int64))
        rep["coef_diffs"] = int((d != 0).sum())
        rep["coef_max_abs"] = int(d.max()) if d.size else 0
    else:
        rep["coef_diffs"] = None
        rep["coef_max_abs"] = None
    return rep


def load_expected() -> dict:
    if not os.path.exists(EXPECTED_PATH):
        return {}
    with open(EXPECTED_PATH) as f:
        return json.load(f)


def check_against_expected(kind: str, name: str, got: dict, expected: dict):
    """got = {"byte_identical": bool, "layers": [layer_report...]}.  A case may not get worse than what is pinned:
    byte identity must hold where pinned, trees must match where pinned, and no more edge pixels / coefficients may differ."""
    exp = expected.get(kind, {}).get(name)
    assert exp is not None, f"{name}: no pinned expectation for '{kind}' in tests/golden/expected_parity.json (regenerate it)"
    if exp["byte_identical"]:
        assert got["byte_identical"], f"{name}: stream is no longer byte-identical to the reference's"
    for i, (g, e) in enumerate(zip(got["layers"], exp["layers"])):
        if e["edge_px"] is not None:
            assert g["edge_px"] is not None and g["edge_px"] <= e["edge_px"], (name, i, "edge map", g["edge_px"], e["edge_px"])
        if e["tree_equal"]:
            assert g["tree_equal"], (name, i, "quadtree differs from the reference's")
            assert g["coef_diffs"] <= e["coef_diffs"] and g["coef_max_abs"] <= max(e["coef_max_abs"], 0), (name, i, "coefficients", g, e)
