"""bench.py's host-side pieces that need no GPU: the `config` object is the same in both arms, chromosomes are dealt to ranks
without loss, the per-rank shared-memory result segments round-trip (minus their CUDA registration), and the reference arm's
line carries the contract's keys."""
import importlib.util
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def args(bench, *extra):
    old = sys.argv
    sys.argv = ["bench.py"] + list(extra)
    try:
        return bench.parse_args()
    finally:
        sys.argv = old


def test_config_is_identical_in_both_arms_and_names_the_workload(bench):
    a = args(bench)
    for world in (1, 2, 8):
        spec = bench.spec_c3(a, world)
        cfg = bench.config_of(spec, world)
        assert cfg == bench.config_of(bench.spec_c3(args(bench, "--impl", "reference"), world), world)
        assert "workload" in cfg and "150bp" in cfg["workload"] and "model" not in cfg
        assert cfg["reads_per_gpu_per_step"] == 10_000_000 and cfg["max_range_index"] == 2
        json.dumps(cfg)
    assert bench.spec_c3(a, 8, reads=50_000_000 // 8, scaling="strong")["reads_per_gpu_per_step"] == 6_250_000


def test_every_contig_has_one_owner_and_the_load_is_balanced(bench):
    for world in (1, 2, 4, 8):
        owner = bench.contig_owner(world)
        assert len(owner) == len(bench.CONTIGS_MBP) and set(owner) == set(range(world))
        load = [sum(m for m, o in zip(bench.CONTIGS_MBP, owner) if o == r) for r in range(world)]
        assert max(load) - min(load) <= max(bench.CONTIGS_MBP)


def test_shared_record_segments_round_trip(bench, monkeypatch):
    """SharedRecords without a GPU: the CUDA registration is stubbed, the file mapping, the views and the clean-up are real"""
    from pindel_b200 import runs as R
    calls = []
    rt = types.SimpleNamespace(cudaHostRegister=lambda p, n, f: calls.append(("reg", n)) or 0, cudaHostUnregister=lambda p: calls.append(("unreg", 0)) or 0)
    torch_stub = types.SimpleNamespace(cuda=types.SimpleNamespace(cudart=lambda: rt))
    monkeypatch.setenv("MASTER_PORT", "45991")
    segs = [bench.SharedRecords(r, 1000, 5000, torch_stub) for r in range(2)]
    for r, s in enumerate(segs):
        assert s.out.dtype == R.OUT_DT and s.out.size == 1000 and s.runs.dtype == R.RUN_DT and s.runs.size == 5000
        s.out["n_close_runs"][:] = r + 1
        s.runs["len0"][:] = 7 + r
    for r in range(2):      # what rank 0 does: map the other rank's file
        mm = np.memmap("%s%d" % (segs[0].tag, r), dtype=np.uint8, mode="r")
        assert np.all(mm[:1000 * R.OUT_DT.itemsize].view(R.OUT_DT)["n_close_runs"] == r + 1)
        assert np.all(mm[1000 * R.OUT_DT.itemsize:].view(R.RUN_DT)["len0"] == 7 + r)
        del mm
    paths = [s.path for s in segs]
    for s in segs:
        s.close()
    assert not any(os.path.exists(p) for p in paths)
    assert [c[0] for c in calls] == ["reg", "reg", "unreg", "unreg"] and calls[0][1] == segs[0].nbytes


def test_no_device_is_an_error_and_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout) and not p.stdout.strip().startswith("{")
