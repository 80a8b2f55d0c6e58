#!/usr/bin/env python
"""bench.py -- spectrogram frames/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One process per GPU (under torchrun for N > 1: RANK / LOCAL_RANK / WORLD_SIZE from the
environment).  A "step" is one pass of the fused spectrogram path over one rank's shard of
the recording: the N=4096 Hann 50 %-overlap periodogram of a 1-hour 48 kHz synthetic
QRSS/DFCW stream (84 375 frames; 691 MB of float32 samples in, 691 MB of PSD rows out --
both far larger than the 126 MB L2, so no flush is needed between steps).  Ranks are time
shards of an N-hour recording (weak scaling): no data-path collective, torch.distributed is
only used for the barrier and the max-over-ranks of the timings.

value  = frames/s with the samples already resident in HBM (device time, CUDA events on the
         library's stream, max over ranks).
e2e    = the same metric through glfer_gram_run() with pinned HOST buffers: H2D of the samples
         and D2H of the rows inside the timed region, pipelined in chunks on two streams.
roofline = algorithmic bytes (hop*4 + (N/2+1)*4 per frame) / the fused kernel's own duration,
         against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
cpu_baseline / --impl reference = the reference's own C code (oracle/_ref, float radix-2
         build = what glfer ships) on the host cores, on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 48000
_STDOUT_FD = 1
WORKLOADS = {
    # name: (description, plan kwargs, seconds of signal per rank)
    "metric": ("periodogram N=4096 Hann 50% ovl, 1 h @ 48 kHz per GPU (BASELINE metric config)",
               dict(n=4096, window_type=0, overlap=0.5, sub_mean=True), 3600),
    "c1": ("periodogram N=1024 Hanning 50% ovl (BASELINE configs[0] shape), 1 h @ 48 kHz per GPU",
           dict(n=1024, window_type=0, overlap=0.5, sub_mean=True), 3600),
    "n2048": ("periodogram N=2048 Hanning 50% ovl, 1 h @ 48 kHz per GPU",
              dict(n=2048, window_type=0, overlap=0.5, sub_mean=True), 3600),
    "c2": ("periodogram N=4096 Kaiser 75% ovl + avg.c plain averaging, 1 h @ 48 kHz",
           dict(n=4096, window_type=7, overlap=0.75, sub_mean=True, avg_mode=2, avg_depth=4, avg_minbin=34,
                avg_maxbin=102), 3600),
    "c3": ("multitaper N=4096 K'=8 (mtm_k=7) NW=4 50% ovl, 1 h @ 48 kHz",
           dict(n=4096, mode=1, overlap=0.5, sub_mean=True, mtm_w=4.0, mtm_kmax=7), 3600),
    "c4": ("periodogram N=16384 Hann 50% ovl, 3 h @ 48 kHz per GPU (24 h over 8 GPUs)",
           dict(n=16384, window_type=0, overlap=0.5, sub_mean=True), 3 * 3600),
    "c5": ("multitaper N=32768 K'=16 (mtm_k=15) NW=8 50% ovl, 1 h @ 48 kHz per GPU",
           dict(n=32768, mode=1, overlap=0.5, sub_mean=True, mtm_w=8.0, mtm_kmax=15), 3600),
    "odd": ("periodogram N=4096 Hann 90% ovl (hop 409: irregular geometry, general kernel), 20 min @ 48 kHz",
            dict(n=4096, window_type=0, overlap=0.9, sub_mean=True), 1200),
    "lmp": ("LMP detector (lmp.c) N=4096 rectangular 50% ovl, ring of 4 frames, 1 h @ 48 kHz",
            dict(n=4096, mode=3, overlap=0.5, sub_mean=True, lmp_av=4), 3600),
}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None
        self.th = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "no samples"}
        # "under load": samples in the upper half of the observed power range
        thr = (max(pw) + min(pw)) / 2 if pw else 0
        load = [s for s, p in zip(sm, pw) if p >= thr] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ----------------------------------------------------------------------------- reference arm
def _ref_worker(args):
    kind, kw, nsamp, reps, seed = args
    sys.path.insert(0, ROOT)
    from glfer_b200 import synth
    x = synth.tiled_stream(nsamp, fs=FS, block_s=20.0, seed=seed)
    frames = 0
    t = 0.0
    if kind == "reference":
        from oracle import ref_lib as R
        for _ in range(reps):
            if kw.get("mode", 0) == 1:
                dt, nf = R.time_mtm(x, kw["n"], kw["overlap"], kw["mtm_w"], kw["mtm_kmax"], kw.get("sub_mean", True))
            elif kw.get("mode", 0) == 3:
                t0 = time.perf_counter()
                nf = R.lmp(x, kw["n"], kw["overlap"], kw["lmp_av"], kw.get("sub_mean", True), kind="f32").shape[0]
                dt = time.perf_counter() - t0
            else:
                dt, nf = R.time_periodogram(x, kw["n"], kw["window_type"], kw["overlap"], kw.get("sub_mean", True))
            t += dt
            frames += nf
    else:
        from oracle import glfer_oracle as O
        for _ in range(reps):
            t0 = time.perf_counter()
            if kw.get("mode", 0) == 1:
                rows = O.multitaper(x, kw["n"], kw["overlap"], kw["mtm_w"], kw["mtm_kmax"], kw.get("sub_mean", True))
            elif kw.get("mode", 0) == 3:
                rows = O.lmp(x, kw["n"], kw["overlap"], kw["lmp_av"], kw.get("sub_mean", True))
            else:
                rows = O.periodogram(x, kw["n"], kw["window_type"], kw["overlap"], kw.get("sub_mean", True))
            t += time.perf_counter() - t0
            frames += rows.shape[0]
    return frames, t


def cpu_reference_rate(kw, target_s=6.0, cores=None):
    """frames/s of the reference's own C implementation with one process per host core
    (mtm.c keeps file-static state, so processes, not threads), each timing the reference
    loop `fft_do; fft_psd` / `mtm_do` over its own copy of a bounded sample."""
    from oracle import ref_lib as R
    kind = "reference" if R.available("f32") else "port"
    cores = cores or len(os.sched_getaffinity(0))
    hop = int(kw["n"] * (1.0 - kw["overlap"]))
    # per-core sample: ~2 s of single-core work per repetition at the survey's probe rates
    probe = {1024: 7e4, 4096: 1.7e4, 16384: 2e3, 32768: 9e2}.get(kw["n"], 1e4)
    if kw.get("mode", 0) == 1:
        probe /= 1.6 * (kw["mtm_kmax"] + 2)
    if kind == "port":
        probe *= 0.5
    # bounded per-core buffer (<= 4096 frames), repeated until ~target_s of work per core
    frames_per_rep = max(16, min(int(probe * 2.0), 4096))
    nsamp = frames_per_rep * hop
    reps = max(1, int(round(target_s * probe / frames_per_rep)))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(kind, kw, nsamp, reps, 100 + i) for i in range(cores)])
        wall = time.perf_counter() - t0
    frames = sum(r[0] for r in res)
    tmax = max(r[1] for r in res)
    return {"value": frames / tmax, "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": f"{reps} x {frames_per_rep} frames ({nsamp / FS:.0f} s of signal) per core, all cores concurrently; "
                      f"reference float radix-2 build (gcc -O2), fft_do+fft_psd loop; wall {wall:.1f} s",
            "frames": frames, "seconds": tmax}


# ----------------------------------------------------------------------------- main
def pin_to_gpu_numa_node(local_rank: int):
    """Run this process (and place the pinned host buffers it allocates next) on the CPUs of the
    NUMA node the GPU hangs off: with several ranks per box the end-to-end copies otherwise cross
    the socket interconnect.  Returns the previous affinity (restored for the CPU baseline leg)."""
    old = os.sched_getaffinity(0)
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1} & old
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception as e:      # no NVML, no topology information: leave the affinity alone
        sys.stderr.write(f"bench.py: NUMA pinning skipped ({e})\n")
    return old


def emit(line: dict) -> None:
    """The one JSON line of the contract, on the real stdout."""
    os.write(_STDOUT_FD, (json.dumps(line) + "\n").encode())


def main():
    global _STDOUT_FD
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner
    # under torchrun, for one) is sent to stderr instead
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="metric", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--seconds", type=int, default=0, help="override the seconds of signal per rank")
    ap.add_argument("--kernel-pref", type=int, default=0, help="experiment: 0 auto, 1 general, 2 ring, 3 warp-per-frame, 4 two frames per thread")
    ap.add_argument("--no-submean", action="store_true", help="experiment: opt.autoscale = 0 (no block-mean removal)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    desc, kw, seconds = WORKLOADS[args.workload]
    kw = dict(kw)
    if args.no_submean:
        kw["sub_mean"] = False
    if args.seconds:
        seconds = args.seconds
    n = kw["n"]
    hop = int(n * (1.0 - kw["overlap"]))
    bins = n // 2 + 1
    ntap = kw["mtm_kmax"] + 1 if kw.get("mode", 0) == 1 else 1
    config = {"workload": desc, "n": n, "hop": hop, "window": "Hann" if kw.get("window_type", 5) == 0 else kw.get("window_type"),
              "tapers": ntap, "sub_mean": bool(kw.get("sub_mean", True)), "sample_rate": FS,
              "seconds_per_gpu": seconds, "frames_per_gpu": seconds * FS // hop,
              "parallelism": f"time-sharded x{world}, (N-hop)-sample halo, no collective",
              "l2": "inputs and outputs each > 126 MB L2; no flush between steps"}

    if args.impl == "reference":
        # the reference's CPU implementation on this host; rank 0 alone works
        if rank != 0:
            return
        steps, warm = max(1, args.steps), max(0, args.warmup)
        vals = []
        base = None
        for i in range(warm + steps):
            base = cpu_reference_rate(kw, target_s=4.0)
            if i >= warm:
                vals.append(base)
        frames = sum(v["frames"] for v in vals)
        secs = sum(v["seconds"] for v in vals)
        v = frames / secs
        line = {"impl": "reference", "metric": "spectrogram_frames_per_sec", "value": v, "unit": "frames/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * secs / steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic QRSS/DFCW multi-tone + noise, 20 s block tiled", "config": config,
                "samples_per_sec": v * hop,
                "cpu_baseline": {"value": v, "unit": "frames/s", "cores": base["cores"], "kind": base["kind"],
                                 "sample": base["sample"]},
                "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from glfer_b200 import api, shard, synth

    if not torch.cuda.is_available() or api.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (libglfer_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    affinity0 = pin_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # this rank's shard of the (world x seconds) recording
    nframes_total = world * seconds * FS // hop
    first, nf = shard.frame_range(nframes_total, world, rank)
    ring = kw.get("lmp_av", 0) if kw.get("mode", 0) == 3 else (kw.get("avg_depth", 0) if kw.get("avg_mode") else 0)
    lo, hi = shard.sample_span(n, hop, first, nf, sub_mean=kw.get("sub_mean", True), avg_depth=ring)
    nsamp = hi - lo
    x_host = api.pinned_empty((nsamp,), np.float32)
    x_host[:] = synth.tiled_stream(nsamp, fs=FS, block_s=20.0, seed=0x5EED + rank)
    if args.kernel_pref:
        api.set_kernel_preference(args.kernel_pref)
    plan = api.GramPlan(device=local_rank, **kw)
    plan.stage(x_host, origin=lo)
    plan.sync()

    # ---- device-resident timing -------------------------------------------------
    # the sampler runs from before the warm-up to the end of the end-to-end region (nvidia-smi
    # needs ~0.1 s to start); "under load" = samples in the upper half of the power range
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(3, args.warmup)):
        plan.exec(first, nf, timed=True)
    barrier()
    launches0 = api.kernel_launches()
    step_ms, gram_ms = [], []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        step_ms.append(plan.exec(first, nf, timed=True))
        gram_ms.append(plan.last_gram_ms())
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = api.kernel_launches() - launches0
    family = api.last_kernel_family()
    dev_s = max_over_ranks(sum(step_ms) * 1e-3)
    gram_s = sum(gram_ms) * 1e-3 / len(gram_ms)
    value = world * nf * args.steps / dev_s          # all ranks do nf (+-1) frames per step

    # ---- end to end through the host-buffer API -----------------------------------
    e2e = None
    if not args.no_e2e:
        rows_host = api.pinned_empty((nf, bins), np.float32)
        out = {"psd": rows_host}
        if plan.avg:
            out["avg"] = api.pinned_empty((nf, bins), np.float32)
        e2e_steps = max(3, min(args.steps, 8))
        plan.run(x_host, origin=lo, first_frame=first, nframes=nf, out=out)       # warm-up (allocations)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.run(x_host, origin=lo, first_frame=first, nframes=nf, out=out)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        d2h = rows_host.nbytes * (2 if plan.avg else 1)
        e2e = {"value": world * nf * e2e_steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int(x_host.nbytes),
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
               "api": "glfer_gram_run (pinned host buffers, 2-slot chunked pipeline)"}
        checksum = float(rows_host[:: max(1, nf // 97)].sum())
        # the same call with the recording as the 16-bit PCM a WAV file holds (glfer_gram_run_pcm16: the
        # int16 -> float conversion of wav_fmt.c:113 runs on the device, half the bytes go up);
        # reported beside the float32 figure, which stays the headline
        pcm_host = api.pinned_empty((nsamp,), np.int16)
        pcm_host[:] = np.rint(x_host * 32768.0).astype(np.int16)
        plan.run(pcm_host, origin=lo, first_frame=first, nframes=nf, out=out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.run(pcm_host, origin=lo, first_frame=first, nframes=nf, out=out)
        barrier()
        pcm_s = max_over_ranks(time.perf_counter() - t0)
        e2e["pcm16_input"] = {"value": world * nf * e2e_steps / pcm_s, "unit": "frames/s",
                              "h2d_bytes_per_step": int(pcm_host.nbytes), "d2h_bytes_per_step": int(d2h),
                              "ms_per_step": 1e3 * pcm_s / e2e_steps, "api": "glfer_gram_run_pcm16",
                              "rows_identical_to_float_input": bool(float(rows_host[:: max(1, nf // 97)].sum()) == checksum)}
        # the reference's real end-to-end product is the waterfall: 16-bit PCM of a WAV file in, 8-bit palette
        # indices out (main_window_draw, g_main.c:1186-1229, fixed display range = autoscale off).  The
        # spectrogram kernel writes the levels itself: 2 B/sample up, 1 B/bin down, no float row in HBM.
        if not plan.avg and kw.get("mode", 0) != 3:
            lev_host = api.pinned_empty((nf, bins), np.uint8)
            disp = dict(log_scale=True, autoscale=False, max_level_db=-20.0, min_level_db=-80.0, thr_level=0.0,
                        out={"levels": lev_host})
            plan.run_display(pcm_host, origin=lo, first_frame=first, nframes=nf, **disp)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                plan.run_display(pcm_host, origin=lo, first_frame=first, nframes=nf, **disp)
            barrier()
            u8_s = max_over_ranks(time.perf_counter() - t0)
            e2e["pcm16_in_u8_out"] = {"value": world * nf * e2e_steps / u8_s, "unit": "frames/s",
                                      "h2d_bytes_per_step": int(pcm_host.nbytes), "d2h_bytes_per_step": int(lev_host.nbytes),
                                      "ms_per_step": 1e3 * u8_s / e2e_steps, "api": "glfer_gram_run_display_pcm16 (fused 8-bit levels)",
                                      "levels_histogram_nonconstant": bool(lev_host[:: max(1, nf // 97)].min() < lev_host[:: max(1, nf // 97)].max())}
    else:
        checksum = None

    clocks = sampler.stop()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    traffic = None          # DRAM read+write bytes of the kernel per launch, from the committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        if tj.get("workload") == args.workload and not args.seconds:
            traffic = tj["traffic_bytes_per_launch"]
    except Exception:
        pass
    alg_bytes = nf * (hop * 4 + bins * 4)
    achieved = alg_bytes / gram_s / 1e9
    line = {"metric": "spectrogram_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": 1e3 * dev_s / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic QRSS/DFCW multi-tone + noise (int16-quantised, /32768), 20 s block tiled to length",
            "config": config, "samples_per_sec": value * hop,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": family,
                         "kernel_ms": 1e3 * gram_s,
                         # kernels after the spectrogram kernel in a step (frame averaging, LMP statistic): not in `achieved`
                         "post_kernels_ms": max(0.0, 1e3 * dev_s / args.steps - 1e3 * gram_s),
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "fp32_flops_per_launch": nf * ntap * 5 * n * int(np.log2(n)),
                         "fp32_tflops_5nlogn": nf * ntap * 5 * n * np.log2(n) / gram_s / 1e12},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": t_wall,
            "checksum": checksum}
    if not args.no_cpu and world == 1:
        try:
            os.sched_setaffinity(0, affinity0)           # the baseline uses every core of the box
            cb = cpu_reference_rate(kw, target_s=6.0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the GPU line must not be lost to a baseline problem
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
