"""Randomised differential check of the fused CUDA path against the oracle.  Collected by pytest through
tests/test_gpu_fuzz.py (30 cases, -m gpu); also a command-line tool for longer runs:

    python tests/fuzz_vs_oracle.py [n_cases] [seed]

Random shapes (1 .. 700 px per side, incl. degenerate ones), spaces, block ranges (2 .. 256) and quality ranges; batch of 2.
Linear spaces: edge maps, states, leaves bit-exact, coefficients within the T-DCT budget; all spaces: decode <= 1 LSB
against the oracle's decode of the same coefficients (T-NAN pixels excluded as in test_gpu_parity.py)."""
import os
import sys
import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in ("adaptive-edge-aware-jpeg_b200", "tests", "oracle"):
    if os.path.join(_ROOT, _p) not in sys.path:
        sys.path.insert(0, os.path.join(_ROOT, _p))
import oracle as O
from aeaj.codec import get_codec
from synth import synth

SPACES = ["YCbCr", "YCoCg", "YCoCg-R", "OKLAB", "ICaCb", "ICtCp", "JzAzBz"]


def run_cases(n: int, seed: int = 0, verbose: bool = True):
    """-> list of problem descriptions (empty = clean)"""
    rng = np.random.default_rng(seed)
    codec = get_codec(0)
    problems = []

    def report(*a):
        problems.append(" ".join(str(x) for x in a))
        if verbose:
            print(problems[-1])
    for case in range(n):
        space = SPACES[rng.integers(len(SPACES))]
        H = int(rng.choice([1, 2, 3, 5, 17, 64, 100, 129, 255, 256, 300, 511, 640, 700]))
        W = int(rng.choice([1, 2, 4, 7, 33, 64, 96, 130, 256, 257, 320, 512, 520, 700]))
        if min(min(sh) for sh in O.layer_shapes(H, W, space)) < 1:   # a chroma layer would be empty: the reference's cv.resize raises too
            continue
        lo = int(2 ** rng.integers(1, 6)); hi = int(lo * 2 ** rng.integers(0, 8))
        hi = min(hi, 256)
        if O.root_size(max(H // 2, 1), max(W // 4, 1)) < lo:      # smallest layer must hold one minimum block
            lo = 2
        if min(hi, O.root_size(H, W)) // lo > 128:
            hi = lo * 128
        q0 = int(rng.integers(1, 99)); q = (q0, int(rng.integers(q0, 100)))
        b = (lo, max(lo, hi))
        kind = rng.integers(3)
        imgs = []
        for s in range(2):
            if kind == 0:
                im = synth(H, W, seed=int(rng.integers(1 << 20)))
            elif kind == 1:
                im = (rng.integers(0, 256, (H, W, 3)).astype(np.float32) / 255.0).astype(np.float32)
            else:
                im = np.full((H, W, 3), float(rng.random()), np.float32)
            imgs.append(im.astype(np.float32))
        batch = np.stack(imgs)
        tag = f"case {case}: {space} {H}x{W} q{q} b{b} kind {kind}"
        try:
            enc = codec.encode(torch.from_numpy(batch).cuda(), space, q, b, taps=True)
            got = codec.download(enc)
            edges = [e.cpu().numpy() for e in enc.edges]
            dec = codec.decode_encoded(enc, space, q, b).cpu().numpy()
        except Exception as ex:                                   # settings the library rejects must be rejected by the oracle too
            report(tag, "-> library raised", type(ex).__name__, str(ex)[:80])
            continue
        exact = space in ("YCbCr", "YCoCg", "YCoCg-R")
        # the device-side .ajpg stream layout (zigzag blocks) must be the natural layout permuted, and decode back identically
        from aeaj import tables
        encz = codec.encode(torch.from_numpy(batch).cuda(), space, q, b, stream=True)
        gotz = codec.download(encz)
        decz = codec.decode_encoded(encz, space, q, b).cpu().numpy()
        if not np.array_equal(decz, dec):
            report(tag, "zigzag-layout decode differs")
        for k in range(2):
            for i in range(3):
                sizes = got[k][i]["leaves"][:, 2].astype(np.int64)
                offs = np.concatenate([[0], np.cumsum(sizes * sizes)])
                nat, zz = got[k][i]["coef"], gotz[k][i]["coef"]
                okz = len(nat) == len(zz)
                for j in np.random.default_rng(case).choice(len(sizes), size=min(len(sizes), 40), replace=False) if okz and len(sizes) else []:
                    sz = int(sizes[j]); blk = nat[offs[j]:offs[j + 1]]
                    if not np.array_equal(zz[offs[j]:offs[j + 1]], blk[tables.zigzag_ordering(sz)]):
                        okz = False; break
                if not okz:
                    report(tag, f"zigzag stream mismatch (image {k}, layer {i})")
        for k in range(2):
            ref = O.encode_hot(batch[k], space, q, b)
            same_edges = all(np.array_equal(edges[i][k], ref[i]["edge"].astype(np.uint8)) for i in range(3))
            if exact and not same_edges:
                report(tag, "edge maps differ")
            if same_edges:
                for i in range(3):
                    if not (np.array_equal(got[k][i]["states"], ref[i]["states"]) and np.array_equal(got[k][i]["leaves"][:, :3], ref[i]["leaves"])):
                        report(tag, f"quadtree differs (layer {i})")
                        break
                    d = np.abs(got[k][i]["coef"].astype(np.int64) - ref[i]["coef"].astype(np.int64))
                    if d.size and (d.max() > 1 or (d != 0).sum() > (4 if exact else 200)):
                        report(tag, f"coefficients: max {d.max()} n {(d != 0).sum()} (layer {i})")
            ref_dec = O.decode_hot([dict(leaves=got[k][i]["leaves"][:, :3], coef=got[k][i]["coef"]) for i in range(3)], H, W, space, q, b)
            lsb = np.abs((dec[k] * 255).astype(np.uint8).astype(int) - (ref_dec * 255).astype(np.uint8).astype(int))
            if space in ("ICaCb", "ICtCp", "JzAzBz"):
                nan_class = np.all(dec[k] == 1.0, axis=-1) | np.all(ref_dec == 1.0, axis=-1)
                lsb[nan_class] = 0
            if lsb.max() > 1:
                report(tag, "decode LSB", lsb.max())
    if verbose:
        print(f"{n} cases, {len(problems)} problems")
    return problems


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    return 1 if run_cases(n, int(sys.argv[2]) if len(sys.argv) > 2 else 0) else 0


if __name__ == "__main__":
    sys.exit(main())
