"""Two ranks on two GPUs (NCCL) against one rank: the split batch is sharded, the
tessellation replicated, one ncclAllReduce inside lite_pl_eval combines the shards
(SURVEY.md §8e); chains of an ensemble are sharded with no collective.  Needs two
devices: skipped on a one-GPU box (tests/test_multirank_gloo.py covers the host logic
at world size 2 on the CPU)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
pytestmark = pytest.mark.gpu

ACS = ["edge_length", "face_area_2", "face_sum_of_angles", "is_segment_internal"]
THETA = np.array([11.0 / np.pi, 12.0 ** 4 * 0.05, 1.0, -np.log(12.0)])
N_SPLITS, SEED, OFF = 2_000_001, 0x5EED, 9      # odd: uneven shards
N_CHAINS = 1001


def _two_gpus():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


def _worker(rank, world, port, out_dir):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import math
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
    t, _ = lite_b200.generate_crtt(3000, seed=5)
    ctx.tess_upload(t.export())
    ctx.energy_set(ACS)
    ctx.merges_flips()
    ctx.add_splits_philox(N_SPLITS, SEED, OFF)
    n_local, n_global = ctx.count()
    lpl, g, H = ctx.pl_eval(THETA)
    ctx.smf_create(N_CHAINS, 200, seed=3)
    ctx.smf_set_model(["is_segment_internal"], [-math.log(3.0)])
    rows = ctx.smf_run(2, 300)
    first = ctx.smf_count()[2]
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lpl=lpl, g=g, H=H, n_local=n_local, n_global=n_global,
             rows=rows, first=first)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs")
def test_two_ranks_equal_one_rank(tmp_path):
    import math
    import torch.multiprocessing as mp
    import lite_b200
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(str(tmp_path / ("rank%d.npz" % r))) for r in (0, 1))
    assert int(r0["n_local"]) + int(r1["n_local"]) == N_SPLITS == int(r0["n_global"])
    assert int(r0["n_local"]) == N_SPLITS // 2
    # every rank holds the all-reduced result
    assert float(r0["lpl"]) == float(r1["lpl"])
    np.testing.assert_array_equal(r0["H"], r1["H"])
    ctx = lite_b200.Context(0)
    t, _ = lite_b200.generate_crtt(3000, seed=5)
    ctx.tess_upload(t.export())
    ctx.energy_set(ACS)
    ctx.merges_flips()
    ctx.add_splits_philox(N_SPLITS, SEED, OFF)
    lpl, g, H = ctx.pl_eval(THETA)
    assert float(r0["lpl"]) == pytest.approx(lpl, rel=1e-12)
    np.testing.assert_allclose(r0["g"], g, rtol=1e-11, atol=1e-12 * np.abs(g).max())
    np.testing.assert_allclose(r0["H"], H, rtol=1e-11, atol=1e-12 * np.abs(H).max())
    # chains: rank r holds [r*n/2, (r+1)*n/2) and each chain is the one a single rank runs
    ctx.smf_create(N_CHAINS, 200, seed=3)
    ctx.smf_set_model(["is_segment_internal"], [-math.log(3.0)])
    rows = ctx.smf_run(2, 300)
    assert int(r0["first"]) == 0 and int(r1["first"]) == N_CHAINS // 2
    both = np.concatenate([r0["rows"], r1["rows"]], axis=2)
    np.testing.assert_array_equal(both, rows)
    ctx.close()
