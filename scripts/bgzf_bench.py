"""scripts/bgzf_bench.py -- measurement of the BGZF inflater (first half) on one GPU, with
zlib on the box's host cores beside it (zlib is what htslib calls; one block per task, all cores).

    python scripts/bgzf_bench.py [--mb N] [--level L]

Payload: BAM-like records (integers, names, 4-bit sequence, qualities with runs), compressed block by block like a BAM
writer does.  Prints one JSON line: inflated GB/s of the kernel (CUDA events), end to end from host memory (H2D of the
compressed file + kernel, stream left on the device; and with the D2H of the inflated stream), and of zlib."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from pindel_b200 import api, bgzf
    from pindel_b200.synth import make_bam_like_bytes as bam_like
    # 16 MB of distinct records, compressed once per worker seed, tiled up to the requested size with different seeds
    cores = os.cpu_count() or 1
    unit = 16 << 20
    n_units = max(1, (a.mb << 20) // unit)
    with ThreadPoolExecutor(cores) as ex:
        payloads = list(ex.map(lambda k: bam_like(unit, 1000 + k % 8), range(min(n_units, 8))))
        comps = list(ex.map(lambda p: bgzf.compress(p, level=a.level, eof=False), payloads))
    comp = b"".join(comps[k % len(comps)] for k in range(n_units)) + bgzf.EOF_MARKER
    total = sum(len(payloads[k % len(payloads)]) for k in range(n_units))
    buf = np.frombuffer(comp, dtype=np.uint8)
    pg = api.PatternGrowth()
    out = np.empty(total, dtype=np.uint8)
    got = pg.bgzf_inflate(buf, out=out)
    assert got.size == total and got[:unit].tobytes() == payloads[0] and got[-unit:].tobytes() == payloads[(n_units - 1) % len(payloads)]
    res = {"stage": "BGZF inflate", "inflated_bytes": int(total), "compressed_bytes": int(buf.size), "ratio": total / buf.size,
           "blocks": int(np.ceil(unit / 0xff00)) * n_units, "level": a.level}
    for name, kw in (("crc", dict(check_crc=True)), ("nocrc", dict(check_crc=False))):
        kern, dev, host = [], [], []
        for _ in range(a.reps):
            t0 = time.perf_counter(); pg.bgzf_inflate(buf, to_host=False, **kw); dev.append(time.perf_counter() - t0)
            kern.append(pg.bgzf_last_kernel_ms)
            t0 = time.perf_counter(); pg.bgzf_inflate(buf, out=out, **kw); host.append(time.perf_counter() - t0)
        res[name] = {"kernel_ms": float(np.median(kern)), "kernel_GBps_out": total / (float(np.median(kern)) * 1e-3) / 1e9,
                     "kernel_GBps_in_plus_out": (total + buf.size) / (float(np.median(kern)) * 1e-3) / 1e9,
                     "e2e_device_resident_ms": 1e3 * float(np.median(dev)), "e2e_device_resident_GBps_out": total / float(np.median(dev)) / 1e9,
                     "e2e_to_host_ms": 1e3 * float(np.median(host)), "e2e_to_host_GBps_out": total / float(np.median(host)) / 1e9}
    # the same with the file and the output in pinned host memory (pg_host_alloc): what a reader that owns its buffers gets
    pbuf = api.pinned_empty(buf.size, np.uint8)
    pbuf[:] = buf
    pout = api.pinned_empty(total, np.uint8)
    dev, host = [], []
    for _ in range(a.reps):
        t0 = time.perf_counter(); pg.bgzf_inflate(pbuf, to_host=False); dev.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); pg.bgzf_inflate(pbuf, out=pout); host.append(time.perf_counter() - t0)
    assert pout[-unit:].tobytes() == payloads[(n_units - 1) % len(payloads)]
    res["pinned_crc"] = {"e2e_device_resident_ms": 1e3 * float(np.median(dev)), "e2e_device_resident_GBps_out": total / float(np.median(dev)) / 1e9,
                         "e2e_to_host_ms": 1e3 * float(np.median(host)), "e2e_to_host_GBps_out": total / float(np.median(host)) / 1e9}
    # zlib on the host cores: the same blocks, inflate + crc32 per block
    blocks = []
    pos = 0
    mv = memoryview(comp)
    while pos < len(comp):
        bsize = int.from_bytes(comp[pos + 16:pos + 18], "little") + 1
        blocks.append(mv[pos + 18:pos + bsize - 8])
        pos += bsize
    sample = blocks[:max(cores * 64, len(blocks) // 4)]

    def one(b):
        d = zlib.decompress(b, -15)
        zlib.crc32(d)
        return len(d)
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(one, sample[:cores * 4]))
        t0 = time.perf_counter()
        nbytes = sum(ex.map(one, sample, chunksize=16))
        dt = time.perf_counter() - t0
    res["cpu_zlib"] = {"GBps_out": nbytes / dt / 1e9, "threads": cores, "sample_bytes": int(nbytes),
                       "note": "zlib inflate + crc32 per block through Python's zlib (GIL released), ThreadPoolExecutor"}
    t0 = time.perf_counter()
    nb1 = sum(one(b) for b in sample[:2000])
    res["cpu_zlib_1thread_GBps_out"] = nb1 / (time.perf_counter() - t0) / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
