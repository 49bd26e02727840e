"""Multi-rank path on CPU (gloo, world_size 2): newline-aligned byte-range shards, per-rank scan, host merge with a
prefix sum of shard line counts — no collective on the data path.  The per-rank scan runs through the host-logic
build (CPU mock of the CUDA engine); the same merge is exercised on the GPU in test_gpu_parity.py."""

import ctypes
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port() -> int:
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]


def _worker(rank: int, world: int, port: int, lib_path: str, data: bytes, patterns, expected, queue) -> None:
    from gpu_api import scan_buffer

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = ctypes.CDLL(lib_path)
    lib.gpugrep_shard_begin.restype = ctypes.c_size_t
    lib.gpugrep_shard_begin.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint, ctypes.c_uint]
    lo = lib.gpugrep_shard_begin(data, len(data), rank, world)
    hi = lib.gpugrep_shard_begin(data, len(data), rank + 1, world)
    shard = data[lo:hi]
    buf = ctypes.create_string_buffer(shard, len(shard))
    rc, records, stats = scan_buffer(lib, ctypes.addressof(buf), len(shard), 0, patterns)
    assert rc == 0
    # the only exchange: one integer per rank (shard line counts) -> exclusive prefix sum
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([stats.lines], dtype=torch.int64))
    base = int(sum(int(c.item()) for c in counts[:rank]))
    mine = [(i, ln + base, text) for (i, ln, text) in records]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        merged = [r for part in gathered for r in part]
        queue.put(merged == expected)
    dist.destroy_process_group()


def test_two_rank_sharded_scan_equals_single_scan(hostmock_lib, oracle_lib):
    from hypergrep_b200 import synth
    from oracle_api import scan_bytes

    data = synth.syslog_bytes(1 << 20, seed=31, lib=hostmock_lib)
    patterns = synth.C2_PATTERNS
    rc, expected, _ = scan_bytes(oracle_lib, data, patterns)
    assert rc == 0 and len(expected) > 100
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, hostmock_lib._name, data, patterns, expected, queue)) for r in range(2)]
    for p in procs:
        p.start()
    ok = queue.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
