"""N>1 host logic on CPU: world_size-2 (and 3) gloo process groups.  Each rank takes its
byte-range shard (swiftover_b200.shard.shard_bounds — the rule of the device sharder), lifts
it with the CPU oracle (test infrastructure), and the rank-ordered concatenation must equal
the single-process result; the max-over-ranks reduction used by bench.py is checked too."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

from conftest import resource_bytes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, bed_name, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_ffi import Oracle, build_oracle
    from swiftover_b200 import shard
    import torch.distributed as dist
    r, _, w = shard.init("gloo")
    assert (r, w) == (rank, world)
    oc = Oracle().chain_parse(resource_bytes("hg19ToHg38.over.chain"))
    data = resource_bytes(bed_name)
    a, b = shard.shard_bounds(data, rank, world)
    lifted, unm, st = oc.lift_bed(data[a:b])
    shard.barrier()
    off, tot = shard.piece_offsets([len(lifted), len(unm)])
    parts = shard.gather_in_rank_order((a, b, lifted, unm, st["n_records"], off, tot))
    mx = shard.max_over_ranks([float(rank + 1), 10.0 - rank])
    if rank == 0:
        q.put((parts, mx))
    shard.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,bed", [(2, "hg19_col123.bed"), (2, "chr1_hg19.bed"), (3, "test.bed")])
def test_rank_sharding_matches_single_process(oracle, world, bed):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, bed, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts, mx = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = resource_bytes(bed)
    # piece offsets (one all_gather of lengths): every rank's piece lands at the sum of the lower ranks' lengths
    run_l = run_u = 0
    for p in parts:
        assert p[5] == [run_l, run_u]
        run_l, run_u = run_l + len(p[2]), run_u + len(p[3])
        assert p[6] == [sum(len(x[2]) for x in parts), sum(len(x[3]) for x in parts)]
    # shards tile the input exactly, on line boundaries
    assert parts[0][0] == 0 and parts[-1][1] == len(data)
    for (a0, b0, *_), (a1, b1, *_) in zip(parts[:-1], parts[1:]):
        assert b0 == a1 and (b0 == 0 or b0 == len(data) or data[b0 - 1] == 0x0A)
    oc = oracle.chain_parse(resource_bytes("hg19ToHg38.over.chain"))
    o_l, o_u, st = oc.lift_bed(data)
    assert b"".join(p[2] for p in parts) == o_l and b"".join(p[3] for p in parts) == o_u
    assert sum(p[4] for p in parts) == st["n_records"]
    assert mx == [float(world), 10.0]


def test_shard_bounds_edges():
    from swiftover_b200.shard import shard_bounds
    for data in (b"", b"a\n", b"a\nb", b"\n\n\n", b"abc", b"a\nbb\nccc\ndddd\n"):
        for world in (1, 2, 3, 5, 8):
            cuts = [shard_bounds(data, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == len(data)
            assert all(a <= b for a, b in cuts)
            assert all(c0[1] == c1[0] for c0, c1 in zip(cuts[:-1], cuts[1:]))
            assert b"".join(data[a:b] for a, b in cuts) == data
            for a, b in cuts:  # every shard starts at a line start
                assert a == 0 or a == len(data) or data[a - 1] == 0x0A
