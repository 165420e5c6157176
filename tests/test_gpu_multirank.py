"""Two ranks on two GPUs against one rank: the split batch is sharded, the tessellation
replicated, the shards' sums are added across the GPUs inside lite_pl_eval —
over NVLink peer memory in the last block of k_reduce by default, by ncclAllReduce when
LITE_B200_ALLREDUCE=nccl — and chains of an ensemble are sharded with no collective.  Needs two
devices: skipped on a one-GPU box (tests/test_multirank_gloo.py covers the host logic at
world size 2 on the CPU; profiles/ keeps the log of this file passing on a two-GPU lease)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
pytestmark = pytest.mark.gpu

ACS = ["edge_length", "face_area_2", "face_sum_of_angles", "is_segment_internal"]
THETA = np.array([11.0 / np.pi, 12.0 ** 4 * 0.05, 1.0, -np.log(12.0)])
THETA2 = THETA * np.array([0.9, 1.05, 0.5, 1.1])
N_SPLITS, SEED, OFF = 2_000_001, 0x5EED, 9      # odd: uneven shards
N_CHAINS = 1001


def _two_gpus():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


def _evaluate(ctx, lite_b200, with_chains):
    """The sequence both the two-rank workers and the single-rank check run."""
    import math
    out = {}
    t, _ = lite_b200.generate_crtt(3000, seed=5)
    ctx.tess_upload(t.export())
    ctx.energy_set(ACS)
    ctx.merges_flips()
    ctx.add_splits_philox(N_SPLITS, SEED, OFF)
    out["n_local"], out["n_global"] = ctx.count()
    # three evaluations in a row: both parities of the mailbox and its reuse
    for k, th in enumerate((THETA, THETA2, THETA)):
        lpl, g, H = ctx.pl_eval(th)
        out["lpl%d" % k], out["g%d" % k], out["H%d" % k] = lpl, g, H
    # NOIS growth by an odd count, then a batch of ONE split: rank 0's share of it is empty
    ctx.add_splits_philox(777, SEED, OFF + N_SPLITS)
    out["lpl3"], out["g3"], out["H3"] = ctx.pl_eval(THETA)
    ctx.clear_splits()
    ctx.add_splits_philox(1, SEED, 5)
    out["n_local_one"] = ctx.count()[0]
    out["lpl4"], out["g4"], out["H4"] = ctx.pl_eval(THETA)
    out["mode"] = ctx.comm_info()[0]
    if with_chains:
        ctx.smf_create(N_CHAINS, 200, seed=3)
        ctx.smf_set_model(["is_segment_internal"], [-math.log(3.0)])
        out["rows"] = ctx.smf_run(2, 300)
        out["first"] = ctx.smf_count()[2]
    return out


def _worker(rank, world, port, out_dir, force_nccl):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    if force_nccl:
        os.environ["LITE_B200_ALLREDUCE"] = "nccl"
    import torch
    import torch.distributed as dist
    import lite_b200
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(lite_b200.nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    ctx = lite_b200.Context(rank, rank, world, bytes(idt.cpu().tolist()))
    out = _evaluate(ctx, lite_b200, with_chains=not force_nccl)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **out)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.fixture(scope="module")
def single():
    import lite_b200
    ctx = lite_b200.Context(0)
    out = _evaluate(ctx, lite_b200, with_chains=True)
    ctx.close()
    return out


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs")
@pytest.mark.parametrize("force_nccl", [False, True], ids=["peer_memory", "nccl"])
def test_two_ranks_equal_one_rank(tmp_path, single, force_nccl):
    import torch.multiprocessing as mp
    port = 29600 + ((os.getpid() + int(force_nccl)) % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path), force_nccl), nprocs=2, join=True)
    r0, r1 = (np.load(str(tmp_path / ("rank%d.npz" % r))) for r in (0, 1))
    assert int(r0["mode"]) == int(r1["mode"]) == (1 if force_nccl else 2)
    assert int(r0["n_local"]) + int(r1["n_local"]) == N_SPLITS == int(r0["n_global"])
    assert int(r0["n_local"]) == N_SPLITS // 2
    assert int(r0["n_local_one"]) == 0 and int(r1["n_local_one"]) == 1
    for k in range(5):
        # every rank holds the same combined result, bit for bit
        assert float(r0["lpl%d" % k]) == float(r1["lpl%d" % k])
        np.testing.assert_array_equal(r0["g%d" % k], r1["g%d" % k])
        np.testing.assert_array_equal(r0["H%d" % k], r1["H%d" % k])
        lpl, g, H = single["lpl%d" % k], single["g%d" % k], single["H%d" % k]
        assert float(r0["lpl%d" % k]) == pytest.approx(lpl, rel=1e-12)
        np.testing.assert_allclose(r0["g%d" % k], g, rtol=1e-11, atol=1e-12 * np.abs(g).max())
        np.testing.assert_allclose(r0["H%d" % k], H, rtol=1e-11, atol=1e-12 * np.abs(H).max())
    assert single["mode"] == 0
    if not force_nccl:
        # chains: rank r holds [r*n/2, (r+1)*n/2) and each chain is the one a single rank runs
        assert int(r0["first"]) == 0 and int(r1["first"]) == N_CHAINS // 2
        both = np.concatenate([r0["rows"], r1["rows"]], axis=2)
        np.testing.assert_array_equal(both, single["rows"])
