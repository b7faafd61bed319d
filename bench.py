#!/usr/bin/env python
"""bench.py -- throughput of the Ravvent inference hot path on B200 (contract: see DESIGN.md §6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--chunks C] [--beam 1|5] [--precision fp32|bf16] [--single-process]

A "step" is one pass of the hot path (encoders -> attention decoder -> beam search) over one batch of
synthetic joint chunks (raw [C,200,1] + event [C,30,5]).

N = 1   BASELINE.json configs[2]: joint model, beam 1, 100k chunks on one B200 (`value`, `e2e`, `roofline`,
        `kernels`), plus in the same line: `beam5` (configs[3], with its own `kernels` / `roofline`),
        `raw_greedy_1k` (configs[0]), `event_only` (configs[1]: reads -> GPU event detection -> snippet
        builder -> event model, beam 1), `depth_3_2` (the reference's best-accuracy configuration),
        `event_path`, `read_path`, `cpu_baseline`.
N > 1   (torchrun, one rank per GPU) BASELINE.json configs[4]: ONE fixed batch of --sweep-chunks (1M) joint
        chunks, read-sharded into contiguous ranges (strong scaling).  `value`: device-resident shards,
        max over ranks.  `e2e`: every rank copies its shard in from pageable host memory, decodes it and
        writes its rows at their input positions of one host array shared by the ranks (/dev/shm) -- the
        scatter and the ordered host-side gather are inside the timed region.  `weak_100k_per_gpu` keeps the
        round-1 figure (every rank its own 100k chunks).  NCCL carries only the contract's barrier and the
        max-over-ranks of the timing; there is no collective on the data path.
--single-process   the same 1M sweep through ShardedBasecaller (one process, one host thread + handle +
        staging ring per GPU) instead of torchrun ranks.

metric  = read-equivalent bases/s = chunks/s x 6.4 (stride-6 windows advance ~6.4 bases per chunk on the
          synthetic generator) -- identical definition for the CUDA path and the CPU arm.
value   = device-resident inputs, CUDA-event timed, max over ranks.
e2e     = same metric through Basecaller.beam_search_prediction with plain (pageable) numpy buffers,
          H2D + D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BASES_PER_CHUNK = 6.4
MAX_OUTPUT_LEN = 34          # S = 33 decode iterations
T_RAW, T_EV = 200, 30
# algorithmic work per joint chunk: FLOPs at S=33 and compulsory fp32 HBM bytes
FLOP_PER_CHUNK = {1: 274.98e6, 5: 347.06e6}
REC_FLOP_PER_STEP_ROW = 131072            # recurrent FLOPs per timestep per direction per layer per snippet
REC_BYTES_L0 = 512 + 4                    # per step per row per direction, raw layer 0 (write h + read x)
REC_BYTES_L1 = 2560                       # layer > 0: read 512 pre-gates + write 128 h (fp32)
# dram__bytes_read+write per launch from the committed ncu --set full capture (profiles/), 9472-chunk wave
# dram bytes of ONE launch on a 9 472-chunk wave (ncu --set full captures summarised under profiles/)
NCU_TRAFFIC = {"recurrent_lstm": 9.71e9, "decoder": 63.7e9, "projection_gemm": 9.65e9, "attention": 1.951e9}
NCU_TRAFFIC_BEAM = {"attention": 2.331e9}      # beam >= 2: attention_tc_kernel reads all Tm rows of both fp16 planes (profiles/r2c_ncu_full_dec5_digest.txt)


def synth_chunks(rng, n):
    """Joint chunks: N(0,1) values, random valid length, zero tail, no exact zeros inside."""
    raw = rng.standard_normal((n, T_RAW, 1), dtype=np.float32)
    raw[raw == 0] = 1e-3
    rl = rng.integers(159, 196, size=n)
    raw[np.arange(T_RAW)[None, :] >= rl[:, None]] = 0.0
    ev = rng.standard_normal((n, T_EV, 5), dtype=np.float32)
    ev[ev == 0] = 1e-3
    el = rng.integers(16, 28, size=n)
    ev[np.arange(T_EV)[None, :] >= el[:, None]] = 0.0
    return raw, ev


GEN_BLOCK = 10000


def synth_range(lo, hi, seed=1234):
    """Chunks [lo, hi) of the one synthetic chunk set (independent of how it is sharded): block i of GEN_BLOCK chunks
    comes from default_rng(seed + i)."""
    raws, evs = [], []
    for blk in range(lo // GEN_BLOCK, (max(hi, lo + 1) - 1) // GEN_BLOCK + 1):
        r, e = synth_chunks(np.random.default_rng(seed + blk), GEN_BLOCK)
        a, b = max(lo, blk * GEN_BLOCK) - blk * GEN_BLOCK, min(hi, (blk + 1) * GEN_BLOCK) - blk * GEN_BLOCK
        raws.append(r[a:b]); evs.append(e[a:b])
    return np.concatenate(raws), np.concatenate(evs)


def shard_range(n, rank, world):
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled in-process through NVML
    (pynvml).  A polling `nvidia-smi -lms` child process perturbs the measurement by several percent
    on this driver, so it is only the fallback when NVML cannot be imported."""

    def __init__(self, index, period=0.05):
        self.index, self.period, self.rows, self.stop_flag, self.thread, self.mode = index, period, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.mode = "nvml"
        except Exception:
            self.mode = "none"
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.time(), sm, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.mode != "nvml":
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.thread.join(timeout=1.0)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = [r[1] for r in self.rows if t0 <= r[0] <= t1 + 0.1]
        reasons = set()
        for ts, _, rs in self.rows:
            if t0 <= ts <= t1 + 0.1:
                for k, bit in names.items():
                    if rs & bit:
                        reasons.add(k)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max, "samples": len(sm),
                "reasons": sorted(reasons), "source": "nvml"}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the TF path on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_reference_rate(n_chunks, beam, repeats=1):
    from threadpoolctl import threadpool_limits
    from oracle import model_ref as mr
    w = mr.init_weights(22)
    raw, ev = synth_range(0, n_chunks)
    cores = os.cpu_count() or 1
    with threadpool_limits(limits=cores):         # torchrun exports OMP_NUM_THREADS=1: give the BLAS its cores back
        t0 = time.perf_counter()
        for _ in range(repeats):
            for b0 in range(0, n_chunks, 1024):       # reference predict batch (ravvent_performance_evaluator.py:24)
                enc, mask = mr.encode_input(w, (raw[b0:b0 + 1024], ev[b0:b0 + 1024]), "joint")
                mr.beam_search(w, enc, mask, beam, MAX_OUTPUT_LEN, full_length=True)
        dt = (time.perf_counter() - t0) / repeats
    return n_chunks / dt * BASES_PER_CHUNK, dt, cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    n = args.ref_chunks
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, thr = cpu_reference_rate(n, args.beam)
        if i >= args.warmup:
            vals.append((v, dt))
    v = float(np.mean([a for a, _ in vals])); dt = float(np.mean([b for _, b in vals]))
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "read-equivalent bases/sec (joint model, beam%d)" % args.beam, "value": v,
        "unit": "bases/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "joint raw+event model, beam%d, S=33, %d-chunk sample of the synthetic chunk set, "
                               "predict batch 1024" % (args.beam, n)},
        "cpu_baseline": {"value": v, "unit": "bases/s", "cores": cores, "kind": "port",
                         "sample": "%d joint chunks, beam%d, numpy/BLAS restatement of the TF path "
                                   "(oracle/model_ref.py; TensorFlow is not installable here)" % (n, args.beam)},
        "e2e": {"value": v, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# CUDA arm
# ----------------------------------------------------------------------------------------------
class Timer:
    """The contract's timing: W untimed warm-up steps, then exactly K steps between barrier + synchronize, CUDA events,
    max over ranks; L2 flushed between timed iterations."""

    def __init__(self, dev, local_rank, dist):
        import torch
        self.torch, self.dev, self.local_rank, self.dist = torch, dev, local_rank, dist
        self.l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def run(self, fn, steps, warmup, sample_clocks=False, profile=False):
        """-> dict(ms per step (max over ranks), launches, clocks, last result, per-kernel profile, per-step ms)."""
        torch = self.torch
        from ravvent_basecaller_b200 import _lib
        held = []                           # keep two generations of outputs alive so the caching allocator owns both
        for _ in range(max(warmup, 2)):     # buffer sets before the timed region (the timed loop holds one while making the next)
            held.append(fn())
            held = held[-2:]
        del held
        sampler = ClockSampler(self.local_rank) if (sample_clocks and not os.environ.get('BENCH_NO_CLOCKS')) else None
        self.barrier()
        if sampler:
            sampler.start(); time.sleep(0.3)
        keep = fn()                         # one more untimed step: the GPU idled during the set-up above (clock ramp)
        self.barrier()
        del keep
        n0 = _lib.launch_count()
        t0w = time.time()
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        if profile:
            _lib.profile(True)
        ev0.record()
        out, marks = None, []
        for _ in range(steps):
            if not os.environ.get('BENCH_NO_FLUSH'):
                self.l2_flush.zero_()       # flush L2 between timed iterations
            out = fn()
            m = torch.cuda.Event(enable_timing=True); m.record(); marks.append(m)
        ev1.record()
        self.barrier()
        t1w = time.time()
        ms = ev0.elapsed_time(ev1)
        launches = _lib.launch_count() - n0
        clocks = sampler.stop(t0w, t1w) if sampler else None
        prof = None
        if profile:
            prof = _lib.profile_read()
            _lib.profile(False)
        prev, per_step = ev0, []
        for m in marks:
            per_step.append(round(prev.elapsed_time(m), 2)); prev = m
        return {"ms": self.max_over_ranks(ms) / steps, "launches": launches, "clocks": clocks, "out": out, "prof": prof,
                "per_step": per_step}


def roofline_entry(kr, ms_step, steps, peak, peak_src, precision, valid_rows, beam=1):
    dom = max(kr, key=lambda k: kr[k]["ms"])
    traffic = (NCU_TRAFFIC_BEAM.get(dom) if beam >= 2 else None) or NCU_TRAFFIC.get(dom)
    return {"bound": "hbm", "achieved": kr[dom]["gbs"], "peak": peak, "unit": "GB/s",
            "frac": kr[dom]["frac_hbm"], "traffic": traffic if precision == "fp32" else None, "peak_source": peak_src,
            "kernel": dom, "share_of_step": kr[dom]["ms"] / (ms_step * steps),
            "launches": kr[dom]["launches"], "avg_launch_ms": kr[dom]["ms"] / max(1, kr[dom]["launches"]),
            "achieved_tflops": kr[dom]["tflops"], "valid_memory_rows_per_chunk": valid_rows,
            "note": "achieved = algorithmic bytes (DESIGN.md 4.5-4.6: memory rows admitted by the mask x 1 KB per snippet "
                    "and decode step) / CUDA-event time of the kernel's launches in the timed region; peak = measured "
                    "copy bandwidth (a read-only stream can slightly exceed it); traffic = dram bytes of ONE launch on a "
                    "9472-chunk wave from the committed ncu capture (profiles/), launches here cover up to 9472 chunks each"}


def run_ours(args, rank, local_rank, world):
    import torch
    import ravvent_basecaller_b200 as rb

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)      # barrier + max of the timing only
    tm = Timer(dev, local_rank, dist)
    strong = world > 1
    total = args.sweep_chunks if strong else args.chunks
    lo, hi = shard_range(total, rank, world)
    C = hi - lo
    S = MAX_OUTPUT_LEN - 1

    bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, "joint", 0., encoder_depth=2, decoder_depth=1, device=local_rank,
                       precision=args.precision)
    bc.compile(optimizer=None)
    bc.load_weights(seed=22)
    raw_h, ev_h = synth_range(lo, hi)                       # plain (pageable) numpy: what a reference caller holds
    raw_d = torch.from_numpy(raw_h).to(dev); ev_d = torch.from_numpy(ev_h).to(dev)

    # host-side gather target of the sharded e2e run: one [total, S] array pair shared by the ranks
    shm = None
    if strong:
        base = "/dev/shm/rvb_bench_%s" % os.environ.get("MASTER_PORT", "0")
        if rank == 0:
            mm = np.lib.format.open_memmap(base + "_ids.npy", mode="w+", dtype=np.int32, shape=(total, S))
            mm[:, 0] = -1                                   # sentinel: every row must be overwritten by the rank that owns it
            mm.flush(); del mm
            np.lib.format.open_memmap(base + "_sc.npy", mode="w+", dtype=np.float32, shape=(total, S)).flush()
        tm.barrier()
        shm = (np.load(base + "_ids.npy", mmap_mode="r+"), np.load(base + "_sc.npy", mmap_mode="r+"))

    def step_device(beam):
        return bc.beam_search_prediction((raw_d, ev_d), beam, MAX_OUTPUT_LEN)[0]

    def step_host(beam):
        ids, sc = bc.beam_search_prediction((raw_h, ev_h), beam, MAX_OUTPUT_LEN)
        if shm is not None:                                 # ordered gather: this rank's rows at their input positions
            T = ids.shape[1]
            shm[0][lo:hi, :T] = ids
            shm[1][lo:hi, :T] = sc
        return ids

    r1 = tm.run(lambda: step_device(args.beam), args.steps, args.warmup, sample_clocks=True, profile=True)
    other = 5 if args.beam == 1 else 1
    ro = tm.run(lambda: step_device(other), max(1, args.steps // 2), 1, profile=True)
    re = tm.run(lambda: step_host(args.beam), max(1, args.steps // 2), 1)
    ms1, ms_o, ms_e2e = r1["ms"], ro["ms"], re["ms"]
    gathered_ok = None
    if strong:
        ids_chk = np.asarray(re["out"])                     # this rank's own result: its rows of the shared array must equal it
        tm.barrier()
        if rank == 0:                                       # every shard landed where its inputs were
            probe = np.linspace(0, total - 1, 64).astype(np.int64)
            gathered_ok = bool((shm[0][probe, 0] >= 0).all() and (shm[0][:, 0] >= 0).all()
                               and np.array_equal(np.asarray(shm[0][lo:lo + 64, :ids_chk.shape[1]]), ids_chk[:64]))

    weak = None
    if strong and not args.no_weak:
        # round-1 figure: every rank its own 100k chunks (weak scaling of replicas)
        n_w = 100000
        rw, ew = synth_range(rank * n_w, (rank + 1) * n_w, seed=99000)
        rw_d, ew_d = torch.from_numpy(rw).to(dev), torch.from_numpy(ew).to(dev)
        rwk = tm.run(lambda: bc.beam_search_prediction((rw_d, ew_d), args.beam, MAX_OUTPUT_LEN)[0], max(1, args.steps // 2), 1)
        weak = {"value": n_w * world / (rwk["ms"] * 1e-3) * BASES_PER_CHUNK, "unit": "bases/s", "ms_per_step": rwk["ms"],
                "chunks_per_gpu": n_w, "scaling": "weak"}
        del rw_d, ew_d

    rate = lambda ms: total / (ms * 1e-3) if strong else C * world / (ms * 1e-3)
    ids1 = r1["out"]
    ids_np = ids1.cpu().numpy() if hasattr(ids1, "cpu") else np.asarray(ids1)
    is_end = ids_np == 1
    first_end = np.where(is_end.any(axis=1), is_end.argmax(axis=1), ids_np.shape[1])
    called = int((((ids_np >= 3) & (ids_np <= 6)) & (np.arange(ids_np.shape[1])[None, :] < first_end[:, None])).sum())

    # ---- per-kernel rooflines, timed live with the library's CUDA events on the launching stream
    peak, peak_src = measured_peaks()
    # memory rows the input mask admits (masked rows contribute nothing to the softmax and are not read)
    valid_rows = float(((raw_d != 0).all(dim=-1).sum() + (ev_d != 0).all(dim=-1).sum()).item()) / max(C, 1)
    kr = kernel_rooflines(r1["prof"], args.steps, C, args.beam, peak, args.precision, valid_rows)
    kro = kernel_rooflines(ro["prof"], max(1, args.steps // 2), C, other, peak, args.precision, valid_rows)

    extras = {}
    if rank == 0 and world == 1 and not args.no_configs:
        del raw_d, ev_d
        extras = other_configs(args, local_rank, tm)
    if rank == 0:
        h2d = C * (T_RAW + T_EV * 5) * 4
        d2h = C * S * 8 + 4
        cpu = None
        if not args.no_cpu_baseline and world == 1:          # rank 0 at N=1 only (contract)
            v, dt, thr = cpu_reference_rate(args.ref_chunks, args.beam)
            cpu = {"value": v, "unit": "bases/s", "cores": thr, "kind": "port",
                   "sample": "%d joint chunks, beam%d, numpy/BLAS restatement of the TF path (oracle/model_ref.py), "
                             "%.1f s" % (args.ref_chunks, args.beam, dt)}
        workload = ("joint raw+event model (enc 2x BiLSTM-128, dec LSTM-128 + Luong), beam%d, S=33 decode steps, " % args.beam) + (
            "ONE fixed batch of %d synthetic chunks read-sharded over %d GPUs (BASELINE configs[4])" % (total, world) if strong
            else "%d synthetic chunks on one GPU (BASELINE configs[2])" % C)
        line = {
            "metric": "read-equivalent bases/sec (joint model, beam%d)" % args.beam,
            "value": rate(ms1) * BASES_PER_CHUNK, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms1, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f16 operands / f32 accumulate", "data": "synthetic",
            "config": {"workload": workload, "chunks_total": total if strong else C * world, "chunks_per_gpu": C,
                       "max_output_len": MAX_OUTPUT_LEN, "weights": "Keras-default init, seed 22",
                       "l2": "256 MiB buffer written between timed iterations",
                       "parallelism": ("contiguous read shards x%d, one process per GPU, no data-path collective; NCCL carries only the "
                                       "barrier and the max-over-ranks of the timing; e2e gathers the shards in input order into one "
                                       "host array (/dev/shm)" % world) if strong else "single GPU"},
            "chunks_per_s": rate(ms1), "called_bases_per_s": called * (1 if strong else world) / (ms1 * 1e-3),
            "tflops_algorithmic": rate(ms1) * FLOP_PER_CHUNK[args.beam] / 1e12,
            "beam%d" % other: {"value": rate(ms_o) * BASES_PER_CHUNK, "unit": "bases/s", "ms_per_step": ms_o,
                               "chunks_per_s": rate(ms_o), "kernels": kro,
                               "roofline": roofline_entry(kro, ms_o, max(1, args.steps // 2), peak, peak_src, args.precision, valid_rows, other)},
            "e2e": {"value": rate(ms_e2e) * BASES_PER_CHUNK, "unit": "bases/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "host_buffers": "pageable numpy (staged through the handle's pinned ring)",
                    "gathered_in_input_order": gathered_ok},
            "gpu_launches": r1["launches"], "clocks": r1["clocks"],
            "roofline": roofline_entry(kr, ms1, args.steps, peak, peak_src, args.precision, valid_rows, args.beam),
            "kernels": kr, "step_ms": r1["per_step"],
        }
        if weak:
            line["weak_100k_per_gpu"] = weak
        line.update(extras)
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if strong:
        tm.barrier()
        del shm
        if rank == 0:
            for suf in ("_ids.npy", "_sc.npy"):
                try:
                    os.unlink("/dev/shm/rvb_bench_%s%s" % (os.environ.get("MASTER_PORT", "0"), suf))
                except OSError:
                    pass
    if dist is not None:
        dist.destroy_process_group()


def run_single_process(args):
    """BASELINE configs[4] through the product's own multi-GPU call: ShardedBasecaller.beam_search_prediction on ONE host
    batch (pageable numpy) -- one handle, host thread, stream set and pinned staging ring per GPU, results written in input
    order into one host array.  Host wall clock around the call (H2D, compute, D2H, gather all inside)."""
    import torch
    import ravvent_basecaller_b200 as rb
    n_dev = min(args.gpus, torch.cuda.device_count())
    total = args.sweep_chunks
    sb = rb.ShardedBasecaller(128, 128, 128, rb.nuc_tk, "joint", 0., devices=list(range(n_dev)), precision=args.precision)
    sb.load_weights(seed=22)
    raw_h, ev_h = synth_range(0, total)
    res = {}
    for beam in (args.beam, 5 if args.beam == 1 else 1):
        for _ in range(max(1, min(args.warmup, 2))):
            sb.beam_search_prediction((raw_h, ev_h), beam, MAX_OUTPUT_LEN)
        ts = []
        for _ in range(max(1, args.steps)):
            for d in range(n_dev):
                torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            ids, sc = sb.beam_search_prediction((raw_h, ev_h), beam, MAX_OUTPUT_LEN)
            ts.append(time.perf_counter() - t0)
        res[beam] = float(np.mean(ts))
    dt = res[args.beam]
    S = MAX_OUTPUT_LEN - 1
    line = {"metric": "read-equivalent bases/sec (joint model, beam%d)" % args.beam, "value": total / dt * BASES_PER_CHUNK,
            "unit": "bases/s", "n_gpus": n_dev, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "joint raw+event model, beam%d, S=33, ONE fixed batch of %d synthetic chunks (BASELINE configs[4])"
                                   % (args.beam, total),
                       "parallelism": "single process: ShardedBasecaller, one host thread + handle + pinned staging ring per GPU, "
                                      "contiguous shards, results written in input order into one host array; no NCCL"},
            "e2e": {"value": total / dt * BASES_PER_CHUNK, "unit": "bases/s", "ms_per_step": dt * 1e3,
                    "h2d_bytes_per_step": total * (T_RAW + T_EV * 5) * 4, "d2h_bytes_per_step": total * S * 8,
                    "host_buffers": "pageable numpy", "timing": "host wall clock around the call (it returns host arrays)"},
            "beam%d" % (5 if args.beam == 1 else 1): {"value": total / res[5 if args.beam == 1 else 1] * BASES_PER_CHUNK,
                                                      "ms_per_step": res[5 if args.beam == 1 else 1] * 1e3},
            "gpu_launches": int(rb.launch_count())}
    print(json.dumps(line), flush=True)


def other_configs(args, local_rank, tm):
    """The BASELINE.json configs that are not the headline, each timed like the headline (N = 1 only)."""
    import torch
    import ravvent_basecaller_b200 as rb
    dev = torch.device("cuda", local_rank)
    out = {}
    k = max(2, args.steps // 2)
    # configs[0]: raw-only model, greedy (beam 1), 1k chunks -- Basecaller.greedy_search_prediction
    bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, "raw", 0., device=local_rank, precision=args.precision).load_weights(seed=22)
    raw_h = synth_range(0, 1000)[0]
    raw_d = torch.from_numpy(raw_h).to(dev)
    a = tm.run(lambda: bc.greedy_search_prediction(raw_d, MAX_OUTPUT_LEN)[0], 4 * k, 2)
    b = tm.run(lambda: bc.greedy_search_prediction(raw_h, MAX_OUTPUT_LEN)[0], 4 * k, 2)
    out["raw_greedy_1k"] = {"config": "BASELINE configs[0]: raw-only model, greedy search, 1000 chunks, S=33", "ms_per_step": a["ms"],
                            "value": 1000 / (a["ms"] * 1e-3) * BASES_PER_CHUNK, "unit": "bases/s",
                            "e2e": {"value": 1000 / (b["ms"] * 1e-3) * BASES_PER_CHUNK, "ms_per_step": b["ms"]},
                            "launches_per_step": a["launches"] // (4 * k),
                            "note": "one partial wave (1000 of 9472 rows): latency of 2 x 200 recurrent timesteps + 33 decode steps, not throughput"}
    big = synth_range(0, 100000)[0]
    big_d = torch.from_numpy(big).to(dev)
    a = tm.run(lambda: bc.greedy_search_prediction(big_d, MAX_OUTPUT_LEN)[0], k, 1)
    out["raw_greedy_100k"] = {"ms_per_step": a["ms"], "value": 100000 / (a["ms"] * 1e-3) * BASES_PER_CHUNK, "unit": "bases/s"}
    del bc, raw_d, big_d, big
    # configs[1]: event-only model fed by GPU event detection
    out["event_only"] = event_only_bench(args, local_rank, tm)
    # the reference's best-accuracy configuration: encoder depth 3, decoder depth 2 (accuracy_results_all.lambda.beam5.json)
    n32 = 30000
    bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, "joint", 0., encoder_depth=3, decoder_depth=2, device=local_rank,
                       precision=args.precision).load_weights(seed=22)
    rh, eh = synth_range(0, n32)
    rd, ed = torch.from_numpy(rh).to(dev), torch.from_numpy(eh).to(dev)
    d32 = {"config": "joint model, encoder_depth 3, decoder_depth 2, %d chunks, S=33" % n32}
    for beam in (1, 5):
        a = tm.run(lambda: bc.beam_search_prediction((rd, ed), beam, MAX_OUTPUT_LEN)[0], k, 1)
        d32["beam%d" % beam] = {"ms_per_step": a["ms"], "value": n32 / (a["ms"] * 1e-3) * BASES_PER_CHUNK, "unit": "bases/s"}
    out["depth_3_2"] = d32
    del bc
    # per-snippet early exit of the wave decoder: an end-token bias (as in tests/test_gpu_model.py::test_beam_early_termination)
    # makes the beams of most snippets finish after a few steps, as trained weights do on real reads (mean 22 of 33 steps)
    from ravvent_basecaller_b200 import weights as _w
    ee = {"config": "joint model (2,1), beam 5, %d chunks, S=33; fc end-token bias +b" % n32}
    for bias in (0.0, 1.0, 2.5):
        w = _w.random_weights(22)
        w["decoder/fc/bias"][1] += bias
        bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, "joint", 0., device=local_rank, precision=args.precision).load_weights(w)
        a = tm.run(lambda: bc.beam_search_prediction((rd, ed), 5, MAX_OUTPUT_LEN)[0], k, 1, profile=True)
        ee["bias_%g" % bias] = {"ms_per_step": a["ms"], "steps_executed": int(a["out"].shape[1]),
                                "attention_ms": a["prof"]["attention"]["ms"] / k}
        del bc
    out["early_exit"] = ee
    del rd, ed
    if not args.no_event_path:
        out["event_path"] = event_path_bench(local_rank)
        out["read_path"] = read_path_bench(local_rank, args.beam, args.precision)
    return out


def synth_reads(n_reads, read_len, seed=77, distinct=8):
    rng = np.random.default_rng(seed)
    one = []
    for _ in range(distinct):          # a few distinct reads tiled to n_reads (generation is the slow part on the host)
        n_lvl = read_len // 3 + 8
        dwell = 2 + rng.geometric(1.0 / 7.0, size=n_lvl)
        level = rng.uniform(250.0, 550.0, size=n_lvl)
        sig = np.repeat(level, dwell)[:read_len] + rng.normal(0.0, 8.0, size=read_len)
        one.append(np.rint(sig).astype(np.int32))
    return one, np.concatenate([one[i % distinct] for i in range(n_reads)])


def event_only_bench(args, local_rank, tm, n_reads=96, read_len=60000):
    """BASELINE configs[1] as ONE timed pipeline: raw reads on the device -> K1 event scan (all reads in one launch) ->
    batched snippet builder (all reads in one call) -> event encoder + attention decoder + beam 1.
    96 reads x 60 000 samples ~ 100k snippets."""
    import ctypes as C
    import torch
    import ravvent_basecaller_b200 as rb
    from ravvent_basecaller_b200 import _lib
    from ravvent_basecaller_b200 import data_loader as dl
    dev = torch.device("cuda", local_rank)
    _, sig = synth_reads(n_reads, read_len)
    offs = np.arange(n_reads + 1, dtype=np.int64) * read_len
    d_sig = torch.from_numpy(sig).to(dev)
    det = rb.EventDetector(6, 9, device=local_rank)
    bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, "event", 0., device=local_rank, precision=args.precision).load_weights(seed=22)
    state = {}

    def pipeline():
        # K1 for all reads in one launch, then the batched snippet builder: one device-to-host round trip for the batch
        _, ev_all, soff = dl.load_data_from_signals(d_sig, offs, stride=6, detector=det, with_raw=False)
        state["n"] = int(ev_all.shape[0])
        return bc.beam_search_prediction(ev_all, 1, MAX_OUTPUT_LEN)[0]

    k = max(2, args.steps // 2)
    # this pipeline is host sensitive (~150 us of GPU work per decode step at Tm = 30, against five launches): the wall time
    # comes from an unprofiled run, the kernels' share from a second, profiled one
    res = tm.run(pipeline, k, 1)
    res["prof"] = tm.run(pipeline, k, 1, profile=True)["prof"]
    n = state["n"]
    return {"config": "BASELINE configs[1]: event-only model fed by GPU event detection, beam 1, S=33", "reads": n_reads,
            "samples": int(n_reads * read_len), "snippets": n, "ms_per_step": res["ms"],
            "value": n / (res["ms"] * 1e-3) * BASES_PER_CHUNK, "unit": "bases/s", "samples_per_s": n_reads * read_len / (res["ms"] * 1e-3),
            "kernel_ms_per_step": {kk: round(v["ms"] / k, 3) for kk, v in res["prof"].items() if v["ms"] > 0},
            "note": "wall = event scan (one launch for all reads) + batched snippet builder (one call, one host round trip for the "
                    "snippet count) + event model; the model itself is the kernel_ms figures"}


def kernel_rooflines(prof, steps, chunks, beam, peak_hbm, precision="fp32", valid_rows=float(T_RAW + T_EV)):
    """Per-kernel achieved rates from the library's CUDA-event timings over the timed region.
    Algorithmic work per chunk: see the constants at the top."""
    out = {}
    units = chunks * steps
    # K3: recurrent LSTM (4 launches per wave: raw L0/L1, event L0/L1)
    rec_ms = prof["recurrent_lstm"]["ms"]
    rec_flop = units * (T_RAW + T_EV) * 2 * 2 * REC_FLOP_PER_STEP_ROW
    rec_bytes = units * (T_RAW + T_EV) * 2 * (REC_BYTES_L0 + REC_BYTES_L1)
    out["recurrent_lstm"] = {"ms": rec_ms, "launches": prof["recurrent_lstm"]["launches"],
                             "gbs": rec_bytes / (rec_ms * 1e-3) / 1e9, "tflops": rec_flop / (rec_ms * 1e-3) / 1e12}
    # K2: projection GEMMs of encoder layer 1 (raw + event): 524288 FLOP and (256 in + 1024 out) * 4 B per timestep
    g_ms = prof["projection_gemm"]["ms"]
    g_flop = units * (T_RAW + T_EV) * 524288.0
    # reduced-precision mode: one fp16 plane in (fp32 mode reads hi + lo planes = the same bytes as fp32)
    g_bytes = units * (T_RAW + T_EV) * (256 * (2.0 if precision == "bf16" else 4.0) + 1024 * 4.0)
    out["projection_gemm"] = {"ms": g_ms, "launches": prof["projection_gemm"]["launches"],
                              "gbs": g_bytes / (g_ms * 1e-3) / 1e9, "tflops": g_flop / (g_ms * 1e-3) / 1e12}
    # K4+K5: decoder: values streamed once per decode step (folded query), 33 steps
    d_ms = prof["decoder"]["ms"]
    d_bytes = units * 33 * (T_RAW + T_EV) * 256 * (2.0 if precision == "bf16" else 4.0)   # fp16 memory copy in reduced mode
    d_flop = units * 33 * beam * 546048.0
    a_ms = prof.get("attention", {"ms": 0.0})["ms"]
    if a_ms > 0:
        # wave-level decoder: the attention kernel (one launch per decode step and wave) is timed on its own; algorithmic
        # bytes per launch = valid memory rows x 1 KB per snippet (all beams of a snippet share one pass) + the query / context rows
        a_bytes = units * 33 * (valid_rows * 256 * 4.0 + beam * (256 + 256) * 4.0)
        a_flop = units * 33 * beam * 768.0 * valid_rows
        out["attention"] = {"ms": a_ms, "launches": prof["attention"]["launches"],
                            "gbs": a_bytes / (a_ms * 1e-3) / 1e9, "tflops": a_flop / (a_ms * 1e-3) / 1e12}
        d_bytes = units * 33 * beam * (256 + 512 + 512 + 384 + 256 + 384 + 128 + 128 + 128) * 4.0      # activations of the dense phases
        d_flop -= a_flop
    if d_ms > 0:
        out["decoder"] = {"ms": d_ms, "launches": prof["decoder"]["launches"],
                          "gbs": d_bytes / (d_ms * 1e-3) / 1e9, "tflops": d_flop / (d_ms * 1e-3) / 1e12}
    for k in out:
        out[k]["frac_hbm"] = out[k]["gbs"] / peak_hbm
    return out


def read_path_bench(local_rank, beam, precision, n_reads=5, read_len=60000):
    """The reference's own loop and metric (RavventPerformanceEvaluator.run, ravvent_performance_evaluator.py:24-87,
    bases/s = bases_num / (predict + post-process + merge) of one read at a time) on synthetic .signal/.label files:
    file -> event scan -> snippets -> encoders -> beam search -> read stitching.  One ~1 050-snippet read per call
    is latency-bound on a B200; the headline `value` batches ~95 reads per step instead."""
    import tempfile
    from pathlib import Path
    from ravvent_basecaller_b200.evaluator import RavventPerformanceEvaluator
    rng = np.random.default_rng(123)
    ev = RavventPerformanceEvaluator(beam_width=beam, device=local_rank, precision=precision)
    ev.setup_basecaller(None, "joint")
    res = []
    with tempfile.TemporaryDirectory() as td:
        for i in range(n_reads):
            n_lvl = read_len // 3 + 8
            dwell = 2 + rng.geometric(1.0 / 7.0, size=n_lvl)
            level = rng.uniform(250.0, 550.0, size=n_lvl)
            sig = np.rint(np.repeat(level, dwell)[:read_len] + rng.normal(0.0, 8.0, size=read_len)).astype(np.int64)
            edges = np.concatenate([[0], np.cumsum(dwell)])
            edges = edges[edges < read_len]
            edges = np.append(edges, read_len)
            sp = Path(td) / ("read%d.signal" % i)
            np.savetxt(sp, sig.reshape(1, -1), fmt="%d")
            syms = rng.choice(list("ACGT"), size=len(edges) - 1)
            with open(sp.with_suffix(".label"), "w") as f:
                f.write("".join("%d %d %s\n" % (a, b, c) for a, b, c in zip(edges[:-1], edges[1:], syms)))
            res.append(ev.run(str(sp), chunk_size=4096))
        paths = [str(Path(td) / ("read%d.signal" % i)) for i in range(n_reads)]
        batch_paths = [paths[i % n_reads] for i in range(24)]           # 24 reads (~25 000 snippets) per device call
        ev.run_batch(batch_paths)
        rb_ = ev.run_batch(batch_paths)
    res = res[1:]                                   # the first read warms allocators and lazy kernel loading
    tp = sum(r["total_processing"] for r in res)
    return {"reads": len(res), "samples_per_read": read_len, "bases_per_read": float(np.mean([r["bases_num"] for r in res])),
            "beam": beam, "bases_per_s": sum(r["bases_num"] for r in res) / tp,
            "samples_per_s": sum(r["samples_num"] for r in res) / tp,
            "ms_per_read": {k: 1e3 * float(np.mean([r[k] for r in res])) for k in ("t_data_loading", "t_predicting", "t_merge")},
            "merged_bases_per_read": float(np.mean([len(r["merged_seq"]) for r in res])),
            "batched": {"reads_per_call": len(batch_paths), "bases_per_s": rb_["bases_num"] / rb_["total_processing"],
                        "samples_per_s": rb_["samples_num"] / rb_["total_processing"],
                        "ms": {k: 1e3 * rb_[k] for k in ("t_data_loading", "t_predicting", "t_merge")}},
            "note": "reference metric definition: label bases / (t_predicting + t_postprocessing + t_merge), one read per call; "
                    "t_data_loading (text parsing, event scan, snippet building) is reported but excluded, as in the reference; "
                    "random-init weights, so merged sequences are not meaningful"}


def event_path_bench(local_rank, n_reads=256, read_len=60000, iters=5):
    """K1 (event scan) + snippet builder on synthetic reads: samples/s and achieved HBM GB/s
    (algorithmic 6.5 B/sample), plus the C oracle on one read as the CPU figure."""
    import ctypes
    import torch
    import ravvent_basecaller_b200 as rb
    from ravvent_basecaller_b200 import _lib
    one, sig = synth_reads(n_reads, read_len)
    offs = np.arange(n_reads + 1, dtype=np.int64) * read_len
    dev = torch.device("cuda", local_rank)
    det = rb.EventDetector(6, 9, device=local_rank)
    d_sig = torch.from_numpy(sig).to(dev)
    out = det.detect_batch(d_sig, offs)                       # warm-up
    n_events = int(out["count"].sum().item())
    _lib.profile(True)
    for _ in range(iters):
        det.detect_batch(d_sig, offs)
    prof = _lib.profile_read(); _lib.profile(False)
    ms = prof["event_scan"]["ms"] / max(1, prof["event_scan"]["launches"])
    samples = n_reads * read_len
    bytes_alg = samples * 4 + n_events * 24
    peak, _ = measured_peaks()
    res = {"reads": n_reads, "samples": samples, "events": n_events, "ms_per_launch": ms,
           "samples_per_s": samples / (ms * 1e-3), "gbs": bytes_alg / (ms * 1e-3) / 1e9,
           "frac_hbm": bytes_alg / (ms * 1e-3) / 1e9 / peak,
           "note": "latency / barrier bound (ncu: FP64 pipe ~6 %, the rest is block-barrier time between the serial phases of a chunk), "
                   "not HBM bound; < 0.3 % of the pipeline time"}
    # snippet builder on one read (per-read API), wall-clock incl. its stream sync
    raw0 = one[0]
    rb.data_loader.load_data_from_signal(raw0, stride=6, detector=det)
    torch.cuda.synchronize(dev)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        rs, es = rb.data_loader.load_data_from_signal(raw0, stride=6, detector=det)
        torch.cuda.synchronize(dev)
        ts.append(time.perf_counter() - t0)
    res["read_to_snippets_ms"] = float(np.median(ts)) * 1e3          # wall clock, one 60k-sample read -> padded snippets
    res["snippets_per_read"] = int(rs.shape[0])
    # CPU: C restatement of the reference's per-sample loop, one core, one read
    try:
        subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle")])
        lib = ctypes.CDLL(str(ROOT / "oracle" / "libravvent_oracle.so"))
        lib.rvo_detect_events.restype = ctypes.c_int64
        cap = read_len // 2 + 4
        st = np.empty(cap, np.int32); ln = np.empty(cap, np.int32); mu = np.empty(cap, np.float64); sd = np.empty(cap, np.float64)
        t0 = time.perf_counter()
        for _ in range(20):
            lib.rvo_detect_events(ctypes.c_void_p(raw0.ctypes.data), ctypes.c_int64(read_len), 6, 9, ctypes.c_double(1.4),
                                  ctypes.c_double(9.0), ctypes.c_double(0.2), ctypes.c_void_p(st.ctypes.data),
                                  ctypes.c_void_p(ln.ctypes.data), ctypes.c_void_p(mu.ctypes.data), ctypes.c_void_p(sd.ctypes.data),
                                  ctypes.c_int64(cap))
        res["cpu_c_port_samples_per_s_1core"] = 20 * read_len / (time.perf_counter() - t0)
    except Exception as e:                                     # the CPU figure is optional
        res["cpu_c_port_samples_per_s_1core"] = None
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunks", type=int, default=100000)
    ap.add_argument("--ref-chunks", type=int, default=2048)
    ap.add_argument("--beam", type=int, default=1)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--sweep-chunks", type=int, default=1000000, help="N > 1: size of the ONE fixed batch that is sharded")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-event-path", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip raw_greedy_1k / event_only / depth_3_2 (N = 1)")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-scaling second figure")
    ap.add_argument("--single-process", action="store_true", help="N > 1 through ShardedBasecaller instead of torchrun ranks")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1 and args.single_process:
        run_single_process(args)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-exec under torchrun when called as `python bench.py --gpus N`
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", __file__] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
