"""World-size-2 test of the N > 1 path on CPU (gloo): the candidate batch is cut
into the contiguous shards `lite_b200.shard_range` gives (the partition
lite_candidates_add_splits_philox uses on each rank), every rank evaluates its
shard with the CPU port of the oracle, one all_reduce(SUM) combines the packed
[lpl, grad, hess] buffer — the role ncclAllReduce plays inside lite_pl_eval — and
the result must equal the single-rank evaluation of the whole batch, with the
GLOBAL number of splits in the normaliser (reference lib/ttessel.cpp:3396)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ACS = ["edge_length", "face_area_2", "face_sum_of_angles", "is_segment_internal"]
N, SEED, OFF = 4001, 0x5EED, 17   # odd on purpose: uneven shards


def _worker(rank, world, port, out_path):
    for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import lite_b200
    import oracle_cpu as oc
    from util import load_golden
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    soa, gold = load_golden("t100")
    t = oc.Tess(soa)
    th = gold["pl_acs_1_theta"]
    d = len(ACS)
    j0, cnt = lite_b200.shard_range(N, rank, world)
    he, x, ang = oc.sample_splits(t, N, SEED, OFF, j0=j0, n=cnt)
    ss = oc.split_stats(t, ACS, he, x, ang)["split_stat"]
    if rank == 0:  # merges and flips are not sharded: one rank contributes them
        ms = oc.merge_stats(t, ACS)["merge_stat"]
        fs = oc.flip_stats(t, ACS)["flip_stat"]
    else:
        ms, fs = np.zeros((0, d)), np.zeros((0, d))
    v, g, H = oc.pl(th, ss, fs, ms.sum(0) if len(ms) else np.zeros(d), fs.sum(0) if len(fs) else np.zeros(d),
                    float(soa["int_length"]), n_global=N)
    buf = torch.tensor(np.concatenate([[v], g, H.ravel()]), dtype=torch.float64)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out_path, np.concatenate([[float(cnt)], buf.numpy()]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_allreduce_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lite_b200
    import oracle_cpu as oc
    from util import load_golden
    # shards tile the batch exactly
    for n in (0, 1, 7, N):
        for w in (1, 2, 3, 8):
            parts = [lite_b200.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(w - 1))
    out = str(tmp_path / "r0.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    assert got[0] == N // 2  # rank 0 holds [0, N/2)
    soa, gold = load_golden("t100")
    t = oc.Tess(soa)
    th = gold["pl_acs_1_theta"]
    he, x, ang = oc.sample_splits(t, N, SEED, OFF)
    ss = oc.split_stats(t, ACS, he, x, ang)["split_stat"]
    ms = oc.merge_stats(t, ACS)["merge_stat"]
    fs = oc.flip_stats(t, ACS)["flip_stat"]
    v, g, H = oc.pl(th, ss, fs, ms.sum(0), fs.sum(0), float(soa["int_length"]))
    d = len(ACS)
    want = np.concatenate([[v], g, H.ravel()])
    scale = np.maximum(1.0, np.abs(want).max())
    np.testing.assert_allclose(got[1:], want, rtol=1e-9, atol=1e-12 * scale)
