#!/usr/bin/env python
"""bench.py -- headline benchmark of the pairwise-alignment hot path on B200.

Metric (BASELINE.json): GCUPS = sum over pairs of m*n / seconds / 1e9 (one cell update = T1, T2
and T3 of one (i, j); cells recomputed for traceback are not counted).

Workload at any N: BASELINE config 2 per GPU -- 1M synthetic 150 bp x 150 bp read pairs, local
(Smith-Waterman) score + end cell + traceback ops; even pairs mutated copies, odd pairs random
(seed 20250002 + rank).  Shards are independent: no data-path collective ("weak" scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl ours|reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  `value` = device-resident inputs, CUDA-event timed, max over
ranks.  `e2e` = the same pass through the host-buffer C-ABI call (psa_align_batch) with pinned
host inputs, H2D and D2H copies inside the timed region.  `roofline` = the integer-issue
roofline SURVEY 8(d) defines for this path (peak measured live by psa_peak_int_ops), plus the HBM
view.  `cpu_baseline` = the reference's own CPU implementation (oracle/_ref) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "GCUPS (cell updates/sec)"
READ_LEN = 150
OPS_PER_CELL_LOCAL = 7      # 6 lane-ops per global cell, +1 running max for local
G, H = 1, 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def config_dict(args, world):
    return {"workload": "config2: batch of 1M synthetic 150bp x 150bp DNA read pairs, local SW score + traceback",
            "pairs_per_gpu": args.pairs, "read_len": READ_LEN, "mode": "local", "g": G, "h": H,
            "outputs": "score, end cell, start cell, 2-bit traceback ops",
            "parallelism": f"pair-shards x{world} (no collective)",
            "l2_policy": "inputs larger than L2 (300 MB of bases per step vs 126 MB L2)"}


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
def cpu_baseline(n_sample, seed, with_shipped=False):
    """The reference's CPU implementation of the path, pair-parallel over all host cores
    (the shape of test_n_cores_thread, testing.cpp:269-276), best-case thread budget p=1 per pair
    plus the as-shipped p=32 call on a smaller sample.  kind 'reference' = oracle/_ref built from
    the reference's own sources; 'port' = the oracle restatement when _ref is absent."""
    from oracle import pyoracle as po
    from cse305_parallel_sequence_alignment_b200 import synth
    cores = os.cpu_count() or 1
    A, B = synth.read_pair_batch(n_sample, READ_LEN, seed)
    off, ln = synth.fixed_length_layout(n_sample, READ_LEN)
    a, b = np.ascontiguousarray(A.reshape(-1)), np.ascontiguousarray(B.reshape(-1))
    cells = float(n_sample) * READ_LEN * READ_LEN
    if po.have_ref():
        sec = po.ref_time_batch(a, off, ln, b, off, ln, 1, G, H, cores)
        out = {"value": cells / sec / 1e9, "unit": "GCUPS", "cores": cores, "kind": "reference",
               "sample": f"{n_sample} pairs of 150x150 through main_alignment_function(p=1), one pair per host "
                         f"thread x{cores}, global mode (the reference has no local mode; same cell count), "
                         f"built -O2, stdout to /dev/null",
               "seconds": sec}
        if with_shipped:   # the call exactly as the harness issues it: p=32 (testing.cpp:134)
            n32 = min(n_sample, cores)
            sec32 = po.ref_time_batch(a, off[:n32], ln[:n32], b, off[:n32], ln[:n32], 32, G, H, cores)
            out["as_shipped_p32_gcups"] = n32 * READ_LEN * READ_LEN / sec32 / 1e9
            out["as_shipped_sample_pairs"] = n32
        return out
    # port: thread the linear-space oracle over chunks (ctypes releases the GIL)
    chunks = np.array_split(np.arange(n_sample), cores)
    t0 = time.perf_counter()
    th = [threading.Thread(target=po.score_batch, args=(a, off[c], ln[c], b, off[c], ln[c], G, H, po.LOCAL)) for c in chunks if len(c)]
    [t.start() for t in th]
    [t.join() for t in th]
    sec = time.perf_counter() - t0
    return {"value": cells / sec / 1e9, "unit": "GCUPS", "cores": cores, "kind": "port",
            "sample": f"{n_sample} pairs of 150x150, linear-space oracle port, {cores} threads", "seconds": sec}


def run_reference(args, rank, world):
    """--impl reference: time the reference's own CPU path on this arm's config."""
    if rank != 0:
        return
    n = args.cpu_sample or 4096
    vals = []
    res = None
    for s in range(args.warmup + args.steps):
        res = cpu_baseline(n, 20250002 + s)
        if s >= args.warmup:
            vals.append(res)
    sec = float(np.mean([v["seconds"] for v in vals]))
    value = n * READ_LEN * READ_LEN / sec / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (exact integers)", "data": "synthetic",
            "config": dict(config_dict(args, world), pairs_per_step_sample=n),
            "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": res["cores"], "kind": res["kind"],
                             "sample": res["sample"]},
            "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import cse305_parallel_sequence_alignment_b200 as psa
    from cse305_parallel_sequence_alignment_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    ctx = psa.Context(local_rank)

    # ---- this rank's shard, generated on the host into PINNED memory ----
    n = args.pairs
    A, B = synth.read_pair_batch(n, READ_LEN, synth.SEED_C2 + rank)
    off_np, len_np = synth.fixed_length_layout(n, READ_LEN)
    hA = torch.from_numpy(np.ascontiguousarray(A.reshape(-1))).pin_memory()
    hB = torch.from_numpy(np.ascontiguousarray(B.reshape(-1))).pin_memory()
    hOff = torch.from_numpy(off_np).pin_memory()
    hLen = torch.from_numpy(len_np).pin_memory()
    stride = (2 * READ_LEN + 15) // 16 + 1
    hItems = torch.zeros(n * 10, dtype=torch.int32).pin_memory()
    hOps = torch.zeros(n * stride, dtype=torch.int32).pin_memory()
    items_np = hItems.numpy().view(psa.capi.ITEM_DTYPE)
    ops_np = hOps.numpy().view(np.uint32).reshape(n, stride)

    # ---- device-resident copy for the kernel-only measurement ----
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    dA, dB = hA.to(dev, non_blocking=True), hB.to(dev, non_blocking=True)
    dOff, dLen = hOff.to(dev, non_blocking=True), hLen.to(dev, non_blocking=True)
    dItems = torch.zeros(n * 10, dtype=torch.int32, device=dev)
    dOps = torch.zeros(n * stride, dtype=torch.int32, device=dev)
    cells_per_step = float(n) * READ_LEN * READ_LEN

    def step_device():
        ctx.align_batch_device(dA.data_ptr(), dOff.data_ptr(), dLen.data_ptr(), dB.data_ptr(), dOff.data_ptr(),
                               dLen.data_ptr(), n, READ_LEN, READ_LEN, dItems.data_ptr(), dOps.data_ptr(), stride,
                               psa.LOCAL, G, H, True, stream.cuda_stream)

    def step_e2e():
        # every host array handed to the C-ABI lives in pinned memory (pageable offsets/lengths would turn the
        # library's asynchronous chunk copies into blocking staged ones)
        ctx.align_batch(hA.numpy(), hOff.numpy(), hLen.numpy(), hB.numpy(), hOff.numpy(), hLen.numpy(), psa.LOCAL, G, H,
                        True, items=items_np, ops=ops_np)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from cse305_parallel_sequence_alignment_b200 import sharding

    def max_over_ranks(ms):
        return sharding.max_over_ranks(ms, dev)

    # integer-issue peak, measured live (kind 1 = VIADDMNMX.S16x2: the ALU pipe every DPX cell op runs on)
    peak_lane_ops, _ = ctx.peak_int_ops(1)

    sampler = ClockSampler(local_rank)
    sampler.start()                      # runs through warm-up + timed steps: the GPU is under load throughout
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    kernel_ms = e0.elapsed_time(e1) / args.steps
    launches = ctx.launches - launches0
    kernel_ms = max_over_ranks(kernel_ms)

    # ---- the dominant kernel alone: same fill launches (direction codes included), walk kernels not launched ----
    ctx.set_option("pack_skip_walk", 1)      # measurement hook (csrc/psa_internal.h), not an environment switch
    for _ in range(2):
        step_device()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launches
    fill_steps = max(3, args.steps // 2)
    f0.record(stream)
    for _ in range(fill_steps):
        step_device()
    f1.record(stream)
    barrier()
    fill_ms = max_over_ranks(f0.elapsed_time(f1) / fill_steps)
    # per step: one fill launch and one flagged-pair launch per chunk
    fill_launches_per_step = (ctx.launches - l0) // fill_steps // 2
    ctx.set_option("pack_skip_walk", 0)
    step_device()                        # leave complete results behind for the comparison below
    barrier()

    # ---- end to end through the host-buffer C-ABI (pinned host buffers, copies timed) ----
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    clocks = sampler.stop()              # sampled while the GPU was busy (device-timed loop + e2e loop)

    # sanity: the e2e pass produced the same answers as the device-resident pass
    same = bool(np.array_equal(dItems.cpu().numpy().view(psa.capi.ITEM_DTYPE)["score"], items_np["score"]))

    value = world * cells_per_step / (kernel_ms * 1e-3) / 1e9
    e2e_value = world * cells_per_step / (e2e_ms * 1e-3) / 1e9
    h2d = int(hA.numel() + hB.numel() + 2 * hOff.numel() * 8 + 2 * hLen.numel() * 4)
    d2h = int(hItems.numel() * 4 + hOps.numel() * 4)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        per_gpu_cups = cells_per_step / (kernel_ms * 1e-3)
        PACK = 2                     # .S16x2: one lane-op updates two cells
        # roofline of the dominant kernel on ITS OWN duration (fill-only loop above); the whole-step view beside it
        fill_cups = cells_per_step / (fill_ms * 1e-3)
        achieved_lane = fill_cups * OPS_PER_CELL_LOCAL / PACK
        step_lane = per_gpu_cups * OPS_PER_CELL_LOCAL / PACK
        # dominant kernel = psa_pack_fill_kernel.  Algorithmic bytes per pair: both reads + offsets/lengths,
        # the 40 B result record, and the 5-bit direction codes it streams to the scratch ring
        # (ceil((150 + 7) / 4) blocks x 8 lanes x 128-byte lines for two pairs -> 20 480 B per pair).
        CODE_BYTES_PER_PAIR = ((READ_LEN + 7 + 3) // 4) * 8 * 128 // 2
        alg_bytes = n * (2 * READ_LEN + 2 * 8 + 2 * 4 + 40 + CODE_BYTES_PER_PAIR)
        # DRAM traffic of one fill launch (131 072 pairs) from the committed ncu --set full capture
        # (profiles/r01_pack_fill_tb_ncu.txt: 2.630 GB written + 0.041 GB read)
        NCU_PAIRS_PER_LAUNCH, NCU_DRAM_BYTES = 131072, 2.630448e9 + 0.040932e9
        pairs_per_launch = min(n, 131072)
        roofline = {"bound": "int-alu", "achieved": achieved_lane / 1e12, "peak": peak_lane_ops / 1e12,
                    "unit": "Tlane-op/s", "frac": achieved_lane / peak_lane_ops,
                    "traffic": NCU_DRAM_BYTES * pairs_per_launch / NCU_PAIRS_PER_LAUNCH,
                    "ops_per_cell": OPS_PER_CELL_LOCAL, "pack": PACK,
                    "kernel": "psa_pack_fill_kernel<8,19,LOCAL,DIRS> (.S16x2 lanes, two pairs per register)",
                    "duration_basis": "fill launches alone (option pack_skip_walk, CUDA events on the launch stream, "
                                      "includes the ~1 % flagged-pair kernel)",
                    "kernel_ms_per_step": fill_ms, "launches_per_step": int(fill_launches_per_step),
                    "kernel_ms_per_launch": fill_ms / max(1, fill_launches_per_step),
                    "whole_step_frac": step_lane / peak_lane_ops,
                    "peak_source": "psa_peak_int_ops(VIADDMNMX.S16x2) measured live in this run (ALU pipe, 64 lanes/clk/SM)",
                    "peak_tcups": peak_lane_ops * PACK / OPS_PER_CELL_LOCAL / 1e12,
                    "int32_equivalent_frac": fill_cups * OPS_PER_CELL_LOCAL / peak_lane_ops,
                    "hbm": {"achieved": alg_bytes / (fill_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": alg_bytes / (fill_ms * 1e-3) / 1e9 / hbm_peak,
                            "algorithmic_bytes_per_launch": alg_bytes * pairs_per_launch / n,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
        line = {"metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "s16x2", "data": "synthetic", "config": config_dict(args, world),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms, "matches_device_pass": same},
                "gpu_launches": int(launches), "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.cpu_sample or 20000, synth.SEED_C2, with_shipped=True)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
