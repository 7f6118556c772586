"""Host logic of the multi-GPU halo-split (aeaj/tiled.py) on the CPU: the band partition, and -- with 2 gloo ranks and a
stub in place of libaeaj.so -- the exchange of the shared-workspace handles (every rank must end up with a table that
holds its own allocation at its own rank and the mapping of rank r's exported handle at r)."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-edge-aware-jpeg_b200"))


def test_bands_partition_the_image():
    from aeaj.tiled import band_of
    for world in (1, 2, 4, 8):
        H = 8192
        bands = [band_of(r, world, H) for r in range(world)]
        assert bands[0][0] == 0 and bands[-1][1] == H
        for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
            assert a1 == b0 and a0 < a1
        assert all(lo % 256 == 0 for lo, _ in bands)          # no chroma cell, filter tile or 128-leaf straddles two ranks
    assert band_of(1, 2, 1024, block_max=256) == (512, 1024)    # 256-blocks: bands are multiples of 512 rows
    with pytest.raises(ValueError):
        band_of(0, 4, 2160)                                      # 2160 rows do not split into 4 aligned bands
    with pytest.raises(ValueError):
        band_of(0, 4, 1024, block_max=256)


WORKER = textwrap.dedent("""
    import ctypes as C, json, os, sys
    import torch.distributed as dist
    sys.path.insert(0, %r)
    from aeaj import tiled

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()

    class Lib:                                   # stands in for libaeaj.so: fake device addresses, handles that name their owner
        def __init__(self): self.n = 0; self.calls = []
        def aeaj_peer_alloc(self, nbytes, out):
            self.n += 1; out._obj.value = 0x1000 * (rank + 1) + self.n; return 0
        def aeaj_peer_export(self, ptr, handle):
            v = ptr.value
            for i in range(8): handle[i] = (v >> (8 * i)) & 0xff
            handle[8] = rank; return 0
        def aeaj_peer_open(self, buf, out):
            raw = buf.raw
            out._obj.value = 0x100000 * (raw[8] + 1) + int.from_bytes(raw[:8], "little"); return 0       # "mapping of rank raw[8]'s block"
        def aeaj_peer_close(self, ptr): self.calls.append(("close", ptr.value)); return 0
        def aeaj_peer_free(self, ptr): self.calls.append(("free", ptr.value)); return 0
        def aeaj_plan_set_peers(self, plan, r, w, ws, fl):
            self.calls.append(("set", r, w, [ws[i] for i in range(w)], [fl[i] for i in range(w)])); return 0

    class Info: workspace_bytes = 4096
    class Plan: ptr = C.c_void_p(1); info = Info()

    lib = Lib()
    pe = tiled._Peers(lib, Plan(), rank, world, None)
    setc = [c for c in lib.calls if c[0] == "set"][0]
    pe.close()
    every = [None] * world
    dist.all_gather_object(every, {"set": setc[1:], "own": [pe.ws.value, pe.flags.value], "closed": [c for c in lib.calls if c[0] != "set"]})
    if rank == 0:
        print(json.dumps(every))
    dist.destroy_process_group()
""") % os.path.join(ROOT, "adaptive-edge-aware-jpeg_b200")


def test_two_ranks_exchange_workspace_handles(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29537", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    every = json.loads([l for l in out.stdout.splitlines() if l.startswith("[")][-1])
    for r, e in enumerate(every):
        rr, world, ws, fl = e["set"]
        assert (rr, world) == (r, 2)
        o = 1 - r
        assert ws[r] == e["own"][0] and fl[r] == e["own"][1]                         # own allocation at own rank
        assert ws[o] == 0x100000 * (o + 1) + every[o]["own"][0]                      # the mapping of the OTHER rank's workspace handle
        assert fl[o] == 0x100000 * (o + 1) + every[o]["own"][1]                      # ... and of its barrier flags
        kinds = [c[0] for c in e["closed"]]
        assert kinds.count("close") == 2 and kinds.count("free") == 2                # mappings closed, own blocks freed
