"""N > 1 host logic on CPU: two gloo ranks shard a chunk by read range, each runs the chain operator on
its shard (through the test-only emulator here; on the GPU box each rank drives its own B200), and the
gathered records, merged in read order, equal the oracle's for the whole chunk."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, emu_lib, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lordfast_b200 import api, shard, sim
    w = sim.make_workload(80_000, 7, 1500, 0.12, 0.15, seed=31, sv_frac=0.3)  # same chunk on every rank
    seeds, chains = api.workload_chains(w)
    bounds = shard.shard_bounds(np.diff(w.read_off), world)
    mine = shard.shard_chains(chains["read_id"], bounds, rank)
    g = api.LfGpu(w.pac, len(w.ref), lib_path=emu_lib)
    recs, text, st = g.align_chains(w.reads, w.read_off.astype(np.uint64), w.contig_off, w.contig_len, seeds, chains[mine])
    part = []
    for d in api.records_to_dicts(recs, text):
        cid = int(mine[d.pop("chain")])
        part.append((int(chains["read_id"][cid]), d))
    gathered = [None] * world
    dist.all_gather_object(gathered, part)
    if rank == 0:
        q.put((shard.merge_in_read_order(gathered), [int(b) for b in bounds]))
    g.close()
    dist.destroy_process_group()


def test_two_ranks_shard_and_merge():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    from _common import build_emu
    from lordfast_b200 import sim
    emu_lib = build_emu()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, emu_lib, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, bounds = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    assert bounds[0] == 0 and bounds[-1] == 7 and 0 < bounds[1] < 7
    w = sim.make_workload(80_000, 7, 1500, 0.12, 0.15, seed=31, sv_frac=0.3)
    idx = O.RefIndex(w.ref.tobytes())
    exp = []
    for i in range(w.n_reads):
        a, _ = O.oracle_align_chain(idx, [tuple(int(x) for x in s) for s in w.chain(i)], w.oriented(i).tobytes(), int(w.is_rev[i]))
        exp.extend((i, s) for s in a)
    assert merged == exp


def _seed_worker(rank, world, port, emu_lib, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import test_seed
    from lordfast_b200 import api, fmindex, shard, sim
    ref, reads, off = test_seed.small_case(seed=7, ref_len=10_000, n_reads=5, read_len=600)   # same chunk on every rank
    bounds = shard.shard_bounds(np.diff(off.astype(np.int64)), world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    g = api.LfGpu(sim.pack_pac(ref), len(ref), lib_path=emu_lib)
    g.seed_init(fmindex.build(test_seed.CODE[ref], k_cache=6))        # every rank holds the whole index
    my_off = (off[lo:hi + 1] - off[lo]).astype(np.uint64)
    fwd, fo, rev, ro = g.seed_batch(reads[int(off[lo]):int(off[hi])], my_off, sampling_count=40)
    part = [(lo + i, ([tuple(int(v) for v in s) for s in fwd[int(fo[i]):int(fo[i + 1])]],
                      [tuple(int(v) for v in s) for s in rev[int(ro[i]):int(ro[i + 1])]])) for i in range(hi - lo)]
    gathered = [None] * world
    dist.all_gather_object(gathered, part)
    if rank == 0:
        q.put((shard.merge_in_read_order(gathered), [int(b) for b in bounds]))
    g.close()
    dist.destroy_process_group()


def test_two_ranks_shard_seeding_and_merge():
    """FM-index seeding shards the same way: reads by contiguous ranges, the index replicated, no collective but the gather."""
    sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fm_oracle
    import test_seed
    from _common import build_emu
    emu_lib = build_emu()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_seed_worker, args=(r, 2, port, emu_lib, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, bounds = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    ref, reads, off = test_seed.small_case(seed=7, ref_len=10_000, n_reads=5, read_len=600)
    n = len(off) - 1
    assert bounds[0] == 0 and bounds[-1] == n and 0 < bounds[1] < n
    idx = fm_oracle.TextIndex(test_seed.CODE[ref])
    rb = reads.tobytes()
    exp = [(i, fm_oracle.seed_read(idx, rb[int(off[i]):int(off[i + 1])], sampling_count=40)) for i in range(n)]
    assert [(i, (list(f), list(r))) for i, (f, r) in exp] == [(i, (f, r)) for i, (f, r) in merged]
    assert sum(len(f) + len(r) for _, (f, r) in merged) >= 40
