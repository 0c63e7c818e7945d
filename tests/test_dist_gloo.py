"""CPU, world_size 2 over gloo: the multi-GPU harness of the path (image sharding + the counter
all-reduce / all-gather, kuzushiji_vision_b200/dist.py). The data path itself has no collective."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from kuzushiji_vision_b200 import dist as kdist, synth
    from oracle import torch_oracle as O

    r, lr, w = kdist.init(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = kdist.shard_range(n_images, rank, world)
    # every rank builds the same global batch (seeded) and assigns only its shard with the CPU oracle
    d = synth.assigner_inputs(n_images, 12, 128, 2, synth.P3_P5, seed=5)
    sl = slice(lo, hi)
    out = O.tal_assign(d["pd_scores"][sl], d["pd_bboxes"][sl], d["anc_points"], d["gt_labels"][sl], d["gt_bboxes"][sl],
                       d["mask_gt"][sl], 10, 2, 0.5, 6.0)
    counters = torch.tensor([float(out[2].sum()), float(out[3].sum()), float(hi - lo)], dtype=torch.float64)
    summ = kdist.all_gather_summaries(counters.clone())
    kdist.all_reduce_counters(counters)
    t = kdist.max_over_ranks(float(rank + 1))
    kdist.barrier()
    if rank == 0:
        full = O.tal_assign(d["pd_scores"], d["pd_bboxes"], d["anc_points"], d["gt_labels"], d["gt_bboxes"],
                            d["mask_gt"], 10, 2, 0.5, 6.0)
        q.put((counters.tolist(), summ.tolist(), t, float(full[2].sum()), float(full[3].sum())))
    dist.destroy_process_group()


def test_sharded_counters_match_unsharded():
    world, n_images = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    counters, summ, t, full_sum, full_fg = q.get()
    assert counters[2] == n_images          # every image assigned exactly once
    assert counters[1] == full_fg           # sharded fg count == unsharded
    assert abs(counters[0] - full_sum) < 1e-4 * max(1.0, full_sum)
    assert len(summ) == world and summ[0][2] + summ[1][2] == n_images
    assert t == world                       # max over ranks
