"""Config 4 over N GPUs: one long pair as block-cyclic systolic panels, boundary columns streamed over NVLink P2P.
Launch:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/strip_bench.py
Checks the N-GPU result against the 1-GPU result (rank 0) and, for small L, the CPU oracle; prints
one JSON line (rank 0) with GCUPS at N GPUs and at 1 GPU.  Env: C4_LEN, MODE (1 local / 0 global), REPS,
PANEL_STRIPS (force a panel width), OPTS (JSON dict of psa_ctx options), VARIANTS (JSON list of
{"opts": {...}, "panel_strips": N}: several option sets timed in one launch, one JSON line each)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse305_parallel_sequence_alignment_b200 as psa  # noqa: E402
from cse305_parallel_sequence_alignment_b200 import multigpu, sharding, synth  # noqa: E402
from cse305_parallel_sequence_alignment_b200.capi import ITEM_DTYPE  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    L = int(os.environ.get("C4_LEN", "200000"))
    mode = int(os.environ.get("MODE", str(psa.LOCAL)))
    reps = int(os.environ.get("REPS", "3"))
    forced = int(os.environ.get("PANEL_STRIPS", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    A, B = synth.mutated_pair(L, synth.SEED_C4)
    dA, dB = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    variants = json.loads(os.environ.get("VARIANTS", "null")) or [{"opts": json.loads(os.environ.get("OPTS", "{}")), "panel_strips": forced}]
    for var in variants:
        one_variant(rank, world, dev, stream, L, mode, reps, A, B, dA, dB, var.get("opts", {}), int(var.get("panel_strips", 0)))
    if world > 1:
        dist.destroy_process_group()


def one_variant(rank, world, dev, stream, L, mode, reps, A, B, dA, dB, opts, forced):
    ctx = psa.Context(dev.index)
    for k, v in opts.items():
        ctx.set_option(k, v)
    item = torch.zeros(10, dtype=torch.int32, device=dev)
    pipe = multigpu.CyclicPanels(ctx, L, rank, world)
    ps = forced or pipe.panel_strips(L)

    def run():
        ctx.align_long_cyclic_device(dA.data_ptr(), dB.data_ptr(), L, L, rank, world, ps, item.data_ptr(), L, pipe.xin, pipe.xout,
                                     mode, 1, 2, stream.cuda_stream)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    run(); sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(reps):
        e0.record(stream); run(); e1.record(stream); sync()
        times.append(sharding.max_over_ranks(e0.elapsed_time(e1), dev))
    # back-to-back calls without a barrier in between: the rings are never cleared, rows are numbered cumulatively
    e0.record(stream); run(); run(); run(); e1.record(stream); sync()
    ms_b2b = sharding.max_over_ranks(e0.elapsed_time(e1), dev) / 3
    ms = float(np.median(times))
    mine = item.cpu().numpy().view(ITEM_DTYPE)
    allitems = sharding.gather_items(mine, [1] * world, dev)
    if rank == 0:
        res = multigpu.merge_local_results(allitems) if mode == psa.LOCAL else allitems[multigpu.last_panel_rank(L, world, ps, pipe.strip_cols)]
        out = {"config": f"C4 {L} x {L} {'local' if mode else 'global'} score, {world} GPU(s), block-cyclic systolic panels of {ps} strips x {pipe.strip_cols} columns over NVLink P2P",
               "n_gpus": world, "panel_strips": ps, "opts": opts, "ms": ms, "ms_all": times, "ms_back_to_back": ms_b2b, "gcups": L * L / ms / 1e6, "score": int(res["score"]),
               "end": [int(res["end_i"]), int(res["end_j"])]}
        # single-GPU reference on rank 0 (default kernel choice of psa_align_long_device)
        it1 = torch.zeros(10, dtype=torch.int32, device=dev)
        f1 = lambda: ctx.align_long_device(dA.data_ptr(), dB.data_ptr(), L, L, it1.data_ptr(), 0, 0, mode, 1, 2, False, stream.cuda_stream)
        f1(); torch.cuda.synchronize()
        e0.record(stream); f1(); e1.record(stream); torch.cuda.synchronize()
        one = it1.cpu().numpy().view(ITEM_DTYPE)[0]
        out["ms_1gpu"] = e0.elapsed_time(e1)
        out["gcups_1gpu"] = L * L / out["ms_1gpu"] / 1e6
        fields = ("score", "end_i", "end_j") if mode == psa.LOCAL else ("t1", "t2", "t3", "end_state")
        out["matches_1gpu"] = all(int(res[f]) == int(one[f]) for f in fields)
        if L <= 40000:
            from oracle import pyoracle as po
            lin = po.score_linear(A.tobytes(), B.tobytes(), 1, 2, mode=mode)
            out["matches_oracle"] = (int(res["score"]) == lin.score and
                                     ((int(res["end_i"]), int(res["end_j"])) == (lin.end_i, lin.end_j) if mode == psa.LOCAL
                                      else (int(res["t1"]), int(res["t2"]), int(res["t3"])) == (lin.t1, lin.t2, lin.t3)))
        out["speedup"] = out["ms_1gpu"] / ms
        out["efficiency"] = out["speedup"] / world
        print(json.dumps(out), flush=True)
    sync()
    pipe.close()
    ctx.close()


if __name__ == "__main__":
    main()
