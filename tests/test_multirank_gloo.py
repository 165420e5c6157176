"""World-size-2 test of the host logic of the N > 1 path, on CPU (gloo).

What it covers: the product's own shard arithmetic (`lite_shard_range`, the partition
`lite_candidates_add_splits_philox` applies on each rank) and the product's own final
arithmetic (`lite_pl_combine`, the tail of `lite_pl_eval`: global |S| in the normaliser,
reference lib/ttessel.cpp:3396; merges and flips replicated on every rank and NOT summed
across ranks), with the per-candidate statistics supplied by the CPU port of the oracle and
gloo's all_reduce standing in for the cross-GPU sum of the 1+d+d(d+1)/2 split sums.

What it does not cover: the kernels, the NCCL communicator and the peer-memory exchange in
k_reduce — those need GPUs; tests/test_gpu_multirank.py is the gate for them (two ranks equal
one rank, an empty shard, both exchange paths) wherever two devices exist."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ACS = ["edge_length", "face_area_2", "face_sum_of_angles", "is_segment_internal"]
SEED, OFF = 0x5EED, 17


def packed_sums(theta, stat):
    """[S0, S1_k, S2_kl (k<=l)] of a statistic block, the layout k_reduce produces."""
    d = len(theta)
    e = np.exp(stat @ theta) if len(stat) else np.zeros(0)
    out = [e.sum()]
    out += [(e * stat[:, k]).sum() for k in range(d)]
    out += [(e * stat[:, k] * stat[:, l]).sum() for k in range(d) for l in range(k, d)]
    return np.array(out)


def _worker(rank, world, port, out_dir, n_total):
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
    th = np.asarray(gold["pl_acs_1_theta"], dtype=np.float64)
    d = len(ACS)
    j0, cnt = lite_b200.shard_range(n_total, rank, world)          # the product's partition
    if cnt:
        he, x, ang = oc.sample_splits(t, n_total, SEED, OFF, j0=j0, n=cnt)
        ss = oc.split_stats(t, ACS, he, x, ang)["split_stat"]
    else:
        ss = np.zeros((0, d))
    # merges and flips: evaluated on EVERY rank, as lite_candidates_merges_flips does
    ms = oc.merge_stats(t, ACS)["merge_stat"]
    fs = oc.flip_stats(t, ACS)["flip_stat"]
    S = torch.tensor(packed_sums(th, ss), dtype=torch.float64)
    dist.all_reduce(S, op=dist.ReduceOp.SUM)                         # the only exchange of the path
    lpl, g, H = lite_b200.pl_combine(th, n_total, float(soa["int_length"]), S.numpy(), packed_sums(th, fs),
                                     ms.sum(0), fs.sum(0))
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), cnt=cnt, lpl=lpl, g=g, H=H)
    dist.barrier()
    dist.destroy_process_group()


def _single(n_total):
    import oracle_cpu as oc
    from util import load_golden
    soa, gold = load_golden("t100")
    t = oc.Tess(soa)
    th = gold["pl_acs_1_theta"]
    he, x, ang = oc.sample_splits(t, n_total, SEED, OFF)
    ss = oc.split_stats(t, ACS, he, x, ang)["split_stat"]
    ms = oc.merge_stats(t, ACS)["merge_stat"]
    fs = oc.flip_stats(t, ACS)["flip_stat"]
    return oc.pl(th, ss, fs, ms.sum(0), fs.sum(0), float(soa["int_length"]))


def _run(tmp_path, n_total, salt):
    import torch.multiprocessing as mp
    port = 29500 + ((os.getpid() + salt) % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path), n_total), nprocs=2, join=True)
    return [np.load(str(tmp_path / ("r%d.npz" % r))) for r in (0, 1)]


def test_shards_tile_the_batch():
    import lite_b200
    for n in (0, 1, 7, 4001, 10 ** 8 + 1):
        for w in (1, 2, 3, 8):
            parts = [lite_b200.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(w - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_two_rank_shards_equal_single_rank(tmp_path):
    n_total = 4001                                # odd on purpose: uneven shards
    r0, r1 = _run(tmp_path, n_total, 0)
    assert int(r0["cnt"]) == n_total // 2 and int(r0["cnt"]) + int(r1["cnt"]) == n_total
    v, g, H = _single(n_total)
    for r in (r0, r1):                            # every rank holds the combined result
        assert abs(float(r["lpl"]) - v) <= 1e-9 * abs(v)
        np.testing.assert_allclose(r["g"], g, rtol=1e-9, atol=1e-12 * max(1.0, np.abs(g).max()))
        np.testing.assert_allclose(r["H"], H, rtol=1e-9, atol=1e-12 * max(1.0, np.abs(H).max()))
    assert float(r0["lpl"]) == float(r1["lpl"])


def test_empty_shard_on_one_rank(tmp_path):
    r0, r1 = _run(tmp_path, 1, 1)                 # one split in all: rank 0 holds none
    assert int(r0["cnt"]) == 0 and int(r1["cnt"]) == 1
    v, g, H = _single(1)
    assert abs(float(r0["lpl"]) - v) <= 1e-9 * abs(v)
    np.testing.assert_allclose(r0["H"], H, rtol=1e-9, atol=1e-12 * max(1.0, np.abs(H).max()))
