#!/usr/bin/env python
"""bench.py -- FX Msamples/s per channel pair on the effex spectrum-mode hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo (CUDA, sm_100a)
    python bench.py --impl reference [--steps K] [--warmup W]     # reference CPU path (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU

Workload (BASELINE.json configs[0], the one the metric is quoted on): effex
spectrum mode, bandwidth 2.4e6, num_samp 262144, resolution 4096, 4-tap PFB,
60 s of data = 550 block pairs per GPU (synthetic correlated noise with a
37-sample delay, 577 MB of raw uint8 IQ per GPU -- larger than L2, so no flush
is needed between steps).  A step is one pass of the hot path over those 550
block pairs: unpack -> DC removal -> PFB -> FFT -> X-engine -> one cross-spectrum
row per block.  With N > 1 every rank runs its own 550-block slice (time
sharding, weak scaling), also accumulates the integrated spectrum, and one NCCL
reduce per step combines the small float64 accumulators on rank 0.

value = pair-samples of all ranks / max-over-ranks device time, inputs resident
in HBM.  e2e = the same metric through the host-buffer entry point
(FxEngine.process_host -> fx_process_host): pinned host bytes in, rows out,
H2D/D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S, N, T = 262144, 4096, 4
BW, FC = 2.4e6, 1.4204e9
RUN_SECONDS = 60
N_BLOCKS = int(np.ceil(RUN_SECONDS * BW / S))        # 550
DELAY = 37
ALG_BYTES_PER_BLOCK = 4 * S + 8 * N                  # uint8 I,Q x 2 channels in, complex64 row out
METRIC = "fx_msamples_per_s_per_channel_pair"
UNIT = "Msamples/s"
WORKLOAD = ("effex spectrum mode, configs[0]: bw=2.4e6 num_samp=262144 resolution=4096 ntaps=4 time=60s "
            f"({N_BLOCKS} block pairs/GPU)")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_per_launch():
    """dram bytes per fused-kernel launch from the committed ncu capture, or None."""
    p = os.path.join(ROOT, "profiles", "fused_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                clk = float(f[1]); smax = float(f[2]); pw = float(f[3])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk); power.append(pw)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# CPU baseline / reference arm: the float64 numpy/scipy oracle on host cores
# ---------------------------------------------------------------------------
_CPU_CACHE = {}


def _cpu_worker(args):
    """One worker: `nblk` passes of the reference's per-block chain (float64 oracle) over
    pre-generated block pairs; input generation is outside the timed work."""
    seed, nblk = args
    from oracle import fx_oracle as orc          # checker / reference arm only
    from effex_b200 import synth
    if "raw" not in _CPU_CACHE:
        _CPU_CACHE["raw"] = synth.correlated_pair(2 * S, delay=DELAY, seed=seed)
        _CPU_CACHE["w"] = orc.pfb_window(T, N)
    raw0, raw1 = _CPU_CACHE["raw"]
    w = _CPU_CACHE["w"]
    if nblk == 0:
        return 0.0
    t0 = time.perf_counter()
    for b in range(nblk):
        sl = slice(2 * S * (b & 1), 2 * S * ((b & 1) + 1))
        orc.process_block_u8(raw0[sl], raw1[sl], N, BW, FC, DELAY / BW, "SPECTRUM", T, w)
    return time.perf_counter() - t0


class CpuReference:
    """The oracle port of the reference's per-block chain on `workers` processes
    (independent blocks, like the reference's per-block loop)."""

    def __init__(self, workers=None):
        from concurrent.futures import ProcessPoolExecutor
        self.workers = workers or min(os.cpu_count() or 1, 64)
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        self.ex = ProcessPoolExecutor(max_workers=self.workers)
        list(self.ex.map(_cpu_worker, [(1000 + i, 1) for i in range(4 * self.workers)], chunksize=4))   # warm-up: imports, inputs, window

    def run(self, blocks_per_worker):
        t0 = time.perf_counter()
        list(self.ex.map(_cpu_worker, [(2000 + i, blocks_per_worker) for i in range(self.workers)]))
        wall = time.perf_counter() - t0
        samples = self.workers * blocks_per_worker * S
        sample = (f"{self.workers * blocks_per_worker} block pairs of configs[0] "
                  f"({self.workers} processes x {blocks_per_worker})")
        return samples / wall / 1e6, self.workers, sample

    def close(self):
        self.ex.shutdown()


def cpu_reference_run(blocks_per_worker=4, workers=None):
    ref = CpuReference(workers)
    try:
        return ref.run(blocks_per_worker)
    finally:
        ref.close()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference()
    for _ in range(args.warmup):
        ref.run(1)
    vals, sample, workers = [], "", ref.workers
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, workers, sample = ref.run(args.ref_blocks)
        vals.append(v)
    ms = (time.perf_counter() - t0) * 1e3 / max(args.steps, 1)
    ref.close()
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_per_step": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from effex_b200 import synth, sharding
    from effex_b200.engine import FxEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the FX hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1 and not args.no_numa_bind:
        # host buffers of the e2e leg are allocated below: keep them on the GPU's NUMA node
        from effex_b200 import hostmem
        hostmem.bind_to_gpu(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # synthetic recording slice of this rank: 8 fresh blocks tiled to 550
    raw0, raw1 = synth.tiled_recording(N_BLOCKS, S, base_blocks=8, delay=DELAY, seed=synth.SEED + rank)
    h0 = torch.from_numpy(raw0).pin_memory()
    h1 = torch.from_numpy(raw1).pin_memory()
    d0, d1 = h0.cuda(non_blocking=True), h1.cuda(non_blocking=True)
    eng = FxEngine(S, N, T, device=local, max_blocks=N_BLOCKS)
    if not eng.fused:
        raise SystemExit("fused sm_100a kernel not selected")
    eng.set_delay(BW, FC, DELAY / BW)
    out = (torch.empty((N_BLOCKS, N), dtype=torch.complex64, device="cuda"), None, None)
    # N > 1: two sets of accumulators, so the reduce of step k (on a side stream) overlaps step k+1
    accs = [eng.new_accumulators(), eng.new_accumulators()] if world > 1 else None
    side = torch.cuda.Stream() if world > 1 else None
    torch.cuda.synchronize()
    counter = [0]
    set_free = [None, None]

    def step():
        if world == 1:
            eng.process(d0, d1, N_BLOCKS, out=out, inputs_ready=True)    # the recording is resident in HBM
            return
        which = counter[0] & 1
        acc = accs[which]
        counter[0] += 1
        if set_free[which] is not None:
            eng.stream.wait_event(set_free[which])   # THIS set's previous reduce + clear have finished
        eng.process(d0, d1, N_BLOCKS, out=out, acc=acc, inputs_ready=True)
        side.wait_stream(eng.stream)
        with torch.cuda.stream(side):
            # the one collective of the path: reduce the small per-integration accumulators to rank 0
            sharding.reduce_accumulators(acc, dst=0)
            acc["flat"].zero_()
            set_free[which] = torch.cuda.Event()
            set_free[which].record(side)

    def barrier():
        if world > 1:
            torch.cuda.current_stream().wait_stream(side)
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    eng.reset_counters()
    eng.enable_timing(True)
    cur = torch.cuda.current_stream()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    ev0.record(cur)                     # engine stream work is ordered after/before `cur` by wait_stream
    t_issue0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    host_issue_ms = (time.perf_counter() - t_issue0) * 1e3 / args.steps
    if world > 1:
        cur.wait_stream(side)                        # the last reduce is inside the timed region
    ev1.record(cur)
    barrier()
    t_wall1 = time.time()
    dev_ms = ev0.elapsed_time(ev1)
    launches = eng.kernel_launches()
    kern_ms, kern_n = eng.dominant_kernel_time()
    eng.enable_timing(False)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    value = world * N_BLOCKS * S * args.steps / (max_ms * 1e-3) / 1e6

    # the fused kernel alone (no overlap with the next step's byte-sum pre-pass): a few synchronised launches
    eng.enable_timing(True)
    eng.reset_counters()
    for _ in range(5):
        eng.process(d0, d1, N_BLOCKS, out=out, inputs_ready=True)
        eng.sync()
    iso_ms, iso_n = eng.dominant_kernel_time()
    eng.enable_timing(False)
    kernel_ms_isolated = iso_ms / max(iso_n, 1)

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "value_not_a_bench_number": value, "kernel_ms": kern_ms / max(kern_n, 1)}))
        eng.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- e2e: host buffers through the public host entry point -------------------
    host_out = torch.empty((N_BLOCKS, N), dtype=torch.complex64).pin_memory().numpy()
    r0, r1 = h0.numpy(), h1.numpy()
    eng_h = FxEngine(S, N, T, device=local, max_blocks=64)
    eng_h.set_delay(BW, FC, DELAY / BW)
    e2e_steps = max(3, min(args.steps, 10))
    eng_h.process_host(r0, r1, N_BLOCKS, out=host_out)       # warm-up (allocations, plans)
    eng_h.process_host(r0, r1, N_BLOCKS, out=host_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng_h.process_host(r0, r1, N_BLOCKS, out=host_out)    # synchronous: returns when rows are on the host
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * N_BLOCKS * S * e2e_steps / float(te.item()) / 1e6
    # sanity: the e2e rows agree with the device-resident rows (different segment plans -> not bit-equal)
    dev_rows = out[0].cpu().numpy()
    same = bool(np.abs(host_out - dev_rows).max() <= 2e-6 * np.abs(dev_rows).max())

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        per_launch_ms = kern_ms / max(kern_n, 1)
        achieved = ALG_BYTES_PER_BLOCK * N_BLOCKS / (per_launch_ms * 1e-3) / 1e9 if kern_n else None
        # FP32-pipe view of the same kernel (DESIGN.md "roofline"): per frame and thread the SASS holds 776
        # packed (FFMA2/FADD2/FMUL2, 2 issue cycles of the FMA pipe per warp) and 64 scalar FFMA; one SM
        # sub-partition runs 2 of the CTA's 8 warps.
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        frames = N_BLOCKS * (S // N)
        fma_cycles = (776 * 2 + 64) * 2 * (frames / 148.0)
        fp32_min_ms = fma_cycles / (sm_mhz * 1e3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": f"time-sharded x{world}", "l2": "inputs (577 MB/GPU) larger than L2, no flush",
                       "collective": "1 NCCL reduce of float64 accumulators per step" if world > 1 else "none",
                       "e2e_rows_match_device_rows": same},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(2 * raw0.nbytes),
                    "d2h_bytes_per_step": int(host_out.nbytes), "steps": e2e_steps,
                    "api": "FxEngine.process_host -> fx_process_host (pinned host buffers)"},
            "gpu_launches": int(launches) * world, "host_issue_ms_per_step": host_issue_ms,
            "roofline": {"bound": "hbm", "kernel": "fx::fused4096::fused_kernel_stag", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic_per_launch(), "peak_source": peak_src,
                         "kernel_ms_per_launch": per_launch_ms, "kernel_ms_isolated": kernel_ms_isolated, "kernel_share_of_step": kern_ms / dev_ms if dev_ms else None,
                         "algorithmic_bytes_per_launch": ALG_BYTES_PER_BLOCK * N_BLOCKS,
                         "fp32_pipe": {"frac": fp32_min_ms / per_launch_ms if kern_n else None,
                                       "frac_isolated": fp32_min_ms / kernel_ms_isolated if kernel_ms_isolated else None,
                                       "fma_pipe_ms_at_full_issue": fp32_min_ms,
                                       "basis": "SASS FMA-pipe issue cycles per frame (776 packed x2 + 64 scalar per thread) at the sampled SM clock, 148 SMs"},
                         "note": "FP32-pipe bound, not HBM bound: see DESIGN.md (roofline). kernel_ms_per_launch is event-bracketed inside the timed region, where the byte-sum pre-pass of the next step runs on the same SMs underneath this kernel; kernel_ms_isolated is the same kernel launched alone"},
            "clocks": clocks,
        }
        if world == 1:
            v, cores, sample = cpu_reference_run(args.cpu_blocks)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    eng.close()
    eng_h.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-blocks", type=int, default=4, help="oracle blocks per worker for cpu_baseline")
    ap.add_argument("--ref-blocks", type=int, default=4, help="oracle blocks per worker per step (--impl reference)")
    ap.add_argument("--no-numa-bind", dest="no_numa_bind", action="store_true",
                    help="N>1: do not pin each rank to its GPU's NUMA node")
    ap.add_argument("--profile", action="store_true", help="device-resident leg only (for runs under ncu; not a bench value)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
