#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric: batched tracker
frame-updates/sec, GMM-KF + systematic resampling) on config 2: 4096 independent synthetic
tracks x 500 slots per GPU, 15-component / 12-D PCA arm model, one shared measurement column
per track-frame.

One "step" = one frame: ParticleFilter::update for every track (K->N indicator resample, fused
per-slot KF predict + innovation likelihood + KF update, weight normalisation, N->N systematic
resample) followed by getEstimator + PCA reconstruction.

  python bench.py [--gpus N --steps K --warmup W]            B200 arm (one process per GPU under torchrun)
  python bench.py --impl reference [...]                      the reference's CPU path (oracle/) on host cores

Prints ONE JSON line (rank 0).  `value` = whole-job frame-updates/s with inputs resident in HBM;
`e2e` = the same through the C ABI with pinned HOST buffers (H2D of the step's measurements and
draws, D2H of the per-track pose) inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched tracker frame-updates/sec (GMM-KF+PDAF)"
UNIT = "frame-updates/s"
SEED = 0x5EED0002
BYTES_PER_SLOT_UPDATE = 1500  # read parent 720 + write child 720 + meas 48 + weight 8 + index 4


def workload_name(T, N):
    return (f"config 2: {T} independent synthetic tracks/GPU x {N} slots, data13D_PCA_100000_15_12 (K=15, d=12), "
            "full GMM-KF bank + single-candidate (shared column) update")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        if os.environ.get("MKF_BENCH_SMI_MS") == "0":
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("MKF_BENCH_SMI_MS", "20"), "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark(self, which):
        setattr(self, which, time.perf_counter())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t0", None), getattr(self, "t1", None)
        inside = [r for (ts, r) in self.rows if t0 is not None and t1 is not None and t0 <= ts <= t1 + 0.03]
        window = "timed region" if inside else "warm-up + timed region (region shorter than the sampling period)"
        for r in (inside or [r for (_, r) in self.rows]):
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window,
                "reasons": sorted(reasons)}


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic():
    """dram bytes per k_slot_update launch from the committed ncu --set full summary, if any"""
    p = os.path.join(ROOT, "profiles", "slot_update_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return None


def host_cores():
    """threads the CPU legs use: every core this process may run on (torchrun exports OMP_NUM_THREADS=1,
    which must not throttle the reference arm, so the count is passed explicitly)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def cpu_arm():
    """the reference's CPU implementation of the path: oracle/_ref (the reference's own KF_model.cpp, my_gmm.cpp,
    pf2DRao.cpp compiled against oracle/cvshim) when it was built, else the oracle port.  Returns
    (kind, run(T, N, frames, seed) -> (seconds, threads))."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mkfbodytracker_pdaf_b200 as mk
    m = mk.Model.load(mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL)
    a = m.arrays()
    arrays = {k: a[k] for k in ("means", "covs", "weights", "gamma", "pca_proj", "pca_mean")}
    if os.environ.get("MKF_BENCH_CPU_KIND", "") != "port":
        try:
            import mkf_ref
            if os.path.exists(mkf_ref.SO):
                mkf_ref.lib()

                def run(T, N, frames, seed):
                    return mkf_ref.bench_tracks(arrays, T, N, frames, per_slot=False, seed=seed, jitter=1,
                                                threads=host_cores())
                return "reference", run
        except Exception:
            pass
    import mkf_oracle as orc
    om = orc.Model(*[arrays[k] for k in ("means", "covs", "weights", "gamma", "pca_proj", "pca_mean")])

    def run(T, N, frames, seed):
        secs, used, _ = orc.bench_tracks(om, T, N, frames, per_slot=False, seed=seed, jitter=1, threads=host_cores())
        return secs, used
    return "port", run


CPU_DESC = {"reference": "oracle/_ref: the reference's own src/{KF_model,my_gmm,pf2DRao}.cpp on the OpenCV-subset shim",
            "port": "oracle/mkf_oracle.cpp (C++ restatement)"}


def cpu_baseline(N, budget_s=12.0):
    """the reference CPU path on all host cores over a bounded sample of the same workload"""
    kind, run = cpu_arm()
    cores = host_cores()
    T_s = 8 * cores
    secs, used = run(T_s, N, 2, SEED)  # calibrate
    rate = T_s * 2 / max(secs, 1e-9)
    frames = int(max(2, min(400, budget_s * rate / T_s)))
    secs, used = run(T_s, N, frames, SEED)
    fu = T_s * frames / secs
    return {"value": fu, "unit": UNIT, "cores": used, "kind": kind, "slot_updates_per_s": fu * N,
            "sample": f"{T_s} tracks x {N} slots x {frames} frames of the same synthetic workload ({CPU_DESC[kind]}; "
                      f"one track per task, single-threaded within a track, OpenMP over tracks; {secs:.2f} s)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, run = cpu_arm()
    cores = host_cores()
    N = args.slots
    T_s = 8 * cores
    for _ in range(args.warmup):
        run(T_s, N, 1, SEED)
    used = cores
    dt = 0.0
    for k in range(args.steps):
        secs, used = run(T_s, N, 1, SEED + k)  # timed inside: the frame loop only (filter construction excluded)
        dt += secs
    val = T_s * args.steps / dt
    sample = (f"each step = 1 frame over {T_s} tracks x {N} slots (a bounded sample of the {args.tracks}-track workload), "
              f"{CPU_DESC[kind]}, OpenMP over tracks")
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(args.tracks, N), "sample": sample},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "slot_updates_per_s": val * N, "gpu_launches": 0}
    emit(out)


_JSON_FD = None


def reserve_stdout():
    """stdout carries exactly one JSON line: keep a private handle on it and point fd 1 at stderr for the rest of the
    run, so that nothing a library prints (NCCL's version banner, for one) can land next to that line"""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tracks", type=int, default=4096, help="tracks per GPU")
    ap.add_argument("--slots", type=int, default=500)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist

    import mkfbodytracker_pdaf_b200 as mk
    from mkfbodytracker_pdaf_b200.sharding import gather_summaries, pack_summary, shard_tracks

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if mk.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (libmkf_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the single JSON line: the image exports NCCL_DEBUG=VERSION, whose banner goes to stdout
        # (WARN prints the banner as well; it is dropped unless MKF_NCCL_DEBUG asks for a level)
        os.environ.pop("NCCL_DEBUG", None)
        if os.environ.get("MKF_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = os.environ["MKF_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=dev)
    T, N, K, W = args.tracks, args.slots, args.steps, args.warmup
    F = K + W
    model = mk.Model.load(mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL)
    # everything (our kernels, torch's packing ops, NCCL, the timing events) runs on ONE explicit
    # stream: torch's default stream has handle 0, which the C ABI reads as "make a private stream"
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    batch = mk.TrackBatch(model, T, N, device=local, stream=stream.cuda_stream)
    track0, _ = shard_tracks(world * T, world, rank)  # weak scaling: T tracks on every rank

    # synthetic inputs of every frame, generated on the device by the shared counter-based generator
    meas = torch.empty((F, T, 6), dtype=torch.float64, device=dev)
    ui = torch.empty((F, T), dtype=torch.float64, device=dev)
    up = torch.empty((F, T), dtype=torch.float64, device=dev)
    for f in range(F):
        batch.synth_fill(SEED, track0, f, 1, mk.MEAS_SHARED, meas[f], ui[f], up[f])
    u0 = torch.empty(T, dtype=torch.float64, device=dev)
    batch.synth_fill(SEED, track0, 0xFFFFFF, 1, mk.MEAS_SHARED, meas[0].clone(), u0, None)
    pose = torch.empty((T, model.D), dtype=torch.float64, device=dev)
    wsum_d = torch.empty(T, dtype=torch.float64, device=dev)
    status_d = torch.empty(T, dtype=torch.int32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(f):
        batch.update(meas[f], ui[f], up[f])
        batch.estimate_into(None, pose)

    # ---------------- device-resident run ----------------
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    batch.reset(u0)
    for f in range(W):
        step(f)
    # warm the summary/gather path too (torch loads its kernels lazily on first use)
    batch.summary_into(wsum_d, status_d)
    gather_summaries(pack_summary(pose, wsum_d, status_d), world)
    barrier()
    # per-kernel CUDA events inside the timed region, on every PROF_EVERY-th step: a sampled step pays ~12 us for
    # its four event records (and loses the kernels' programmatic overlap), so sampling all of them would tax the
    # number they are there to explain
    PROF_EVERY = 8
    batch.profile((K + PROF_EVERY - 1) // PROF_EVERY, PROF_EVERY)
    launches0 = mk.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clk.mark("t0")
    e0.record()
    for f in range(W, F):
        step(f)
    # final per-track summaries {pose[D], wsum, status}; the only collective is this gather
    batch.summary_into(wsum_d, status_d)
    gathered = gather_summaries(pack_summary(pose, wsum_d, status_d), world)
    e1.record()
    barrier()
    clk.mark("t1")
    launches = mk.launch_count() - launches0
    clocks = clk.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    prof = batch.profile_read_stages()
    batch.profile(0)
    rec, nslots = batch.shared_records()  # distinct Gaussians the last frame stored (record sharing)
    status_bad = int((batch.status() & (mk._lib.ST_POST_DEGENERATE | mk._lib.ST_CHOL_FAIL)).astype(bool).sum())

    # ---------------- end to end through the C ABI with pinned host buffers ----------------
    # every frame's inputs sit in pinned HOST memory; each step copies them to the device, runs the frame and
    # copies the per-track pose back.  Two variants: "sync" waits for the pose after every step (one frame in
    # flight, what a single-frame caller sees); the headline `e2e` is the pipelined use of the same API
    # (MKF_MEM_HOST_ASYNC: copies ordered on the stream, one synchronisation at the end of the timed region).
    h_meas = meas.cpu().pin_memory()
    h_ui = ui.cpu().pin_memory()
    h_up = up.cpu().pin_memory()
    h_pose = torch.empty((2, T, model.D), dtype=torch.float64).pin_memory()

    def e2e_step(f, mem):
        batch.update(h_meas[f], h_ui[f], h_up[f], mem=mem)
        mk._lib.check(mk._lib.lib.mkf_batch_estimate(batch._h, None, h_pose[f & 1].data_ptr(), mem))

    def e2e_run(mem):
        batch.reset(u0)
        for f in range(W):
            e2e_step(f, mem)
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for f in range(W, F):
            e2e_step(f, mem)
        batch.join()  # the closing event waits for the copies on the library's internal copy streams too
        eb.record()
        barrier()
        t_ms = torch.tensor([ea.elapsed_time(eb)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        return float(t_ms.item())

    ms_e2e_sync = e2e_run(mk.MEM_HOST)
    pose_sync = h_pose[(F - 1) & 1].clone()
    ms_e2e = e2e_run(mk.MEM_HOST_ASYNC)
    assert torch.equal(pose_sync, h_pose[(F - 1) & 1]), "pipelined and synchronous e2e runs must agree"
    pose_check = float(h_pose[(F - 1) & 1][:, :2].mean())

    # ---------------- the same frames with every slot computed (record sharing off) ----------------
    # what per-slot measurements (the association path, config 4) always run: k_slot_update moves SURVEY 8(d)'s
    # 1 500 B for every slot, so this leg is the one to hold against the HBM roofline.  Rank 0, no collectives.
    every_slot = None
    if rank == 0 and os.environ.get("MKF_DEDUP", "1") != "0":
        os.environ["MKF_DEDUP"] = "0"  # read when a batch is created
        try:
            b2 = mk.TrackBatch(model, T, N, device=local, stream=stream.cuda_stream)
        finally:
            del os.environ["MKF_DEDUP"]
        K2 = min(K, 48)
        b2.reset(u0)
        for f in range(W):
            b2.update(meas[f], ui[f], up[f])
            b2.estimate_into(None, pose)
        torch.cuda.synchronize()
        b2.profile((K2 + 3) // 4, 4)
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for f in range(W, W + K2):
            b2.update(meas[f], ui[f], up[f])
            b2.estimate_into(None, pose)
        eb.record()
        torch.cuda.synchronize()
        p2 = b2.profile_read_stages()
        b2.profile(0)
        ms2 = ea.elapsed_time(eb) / K2
        k2_ms = p2["ms_slot_kernel"] / max(p2["n"], 1)
        every_slot = {"steps": K2, "ms_per_step": ms2, "value": T * 1e3 / ms2, "unit": UNIT,
                      "kernel": "k_slot_update<12, 0>", "kernel_ms": k2_ms, "kernel_samples": p2["n"],
                      "algorithmic_bytes_per_launch": T * N * BYTES_PER_SLOT_UPDATE,
                      "achieved": T * N * BYTES_PER_SLOT_UPDATE / (k2_ms * 1e-3) / 1e9 if k2_ms > 0 else None}
        b2.close()

    if rank == 0:
        value = world * T * K / (ms * 1e-3)
        e2e_val = world * T * K / (ms_e2e * 1e-3)
        peak, peak_src = load_peak()
        n_prof = max(prof["n"], 1)
        slot_ms = prof["ms_slot_kernel"] / n_prof
        sharing = rec < nslots
        # units one launch processes: the distinct Gaussians (records) when identical children are shared -- the last
        # frame's count, stationary after the first ~20 frames -- else every slot
        units = rec if sharing else nslots
        achieved = units * BYTES_PER_SLOT_UPDATE / (slot_ms * 1e-3) / 1e9 if slot_ms > 0 else None
        tr = load_traffic()
        split = sharing and prof["ms_share_keys"] > 0
        kname = ("k_slot_update_heads_direct" if split else "k_slot_update_shared") if sharing else "k_slot_update"
        ktr = ((tr or {}).get("kernels") or {}).get(kname) or {}
        if every_slot and every_slot["achieved"]:
            every_slot["peak"] = peak
            every_slot["frac"] = every_slot["achieved"] / peak
            every_slot["traffic"] = (((tr or {}).get("kernels") or {}).get("k_slot_update") or {}).get("dram_bytes_per_launch")
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(T, N), "tracks_per_gpu": T, "slots": N, "components": model.K,
                       "state_dim": model.d, "measurement": "shared column per track-frame", "chol_mode": "CV24_LITERAL",
                       "alias_mode": "INDEPENDENT", "seed": hex(SEED),
                       "l2": f"no flush: per-step working set {2 * T * N * 720 / 1e9:.2f} GB >> 126 MB L2",
                       "parallelism": f"tracks sharded, {world} x {T}; one final NCCL all_gather of per-track summaries"},
            "slot_updates_per_s": value * N,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": ktr.get("dram_bytes_per_launch"),
                         "kernel": ktr.get("kernel", kname), "kernel_ms": slot_ms,
                         "kernel_samples": prof["n"], "kernel_sampling": f"CUDA events on every {PROF_EVERY}th step of the timed region",
                         "units_per_launch": int(units), "slots_per_launch": int(nslots),
                         "algorithmic_bytes_per_launch": int(units) * BYTES_PER_SLOT_UPDATE, "peak_source": peak_src,
                         "distinct_records_fraction": rec / max(nslots, 1),
                         "note": "SURVEY 8(d): 1500 B per slot-update.  With one measurement per track, children that "
                                 "drew the same parent record and component are identical Gaussians and are computed "
                                 "and stored once: a launch processes units_per_launch distinct slot-updates for "
                                 "slots_per_launch slots.  every_slot_computed is the same workload with the sharing "
                                 "off (k_slot_update, 1500 B for every slot)",
                         "effective_all_slots_gbs": nslots * BYTES_PER_SLOT_UPDATE / (slot_ms * 1e-3) / 1e9 if slot_ms > 0 else None,
                         "stage_ms": {"indicator_bounds": prof["ms_bounds"] / n_prof,
                                      "share_keys": prof["ms_share_keys"] / n_prof,
                                      "slot_kernel": slot_ms,
                                      "repair": prof["ms_repair"] / n_prof,
                                      "normalise_resample": prof["ms_resample"] / n_prof},
                         "every_slot_computed": every_slot},
            "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": world * T * 8 * 8, "d2h_bytes_per_step": world * T * model.D * 8,
                    "mode": "pinned host buffers, MKF_MEM_HOST_ASYNC: copies on the library's copy streams overlap the neighbouring frames' kernels, one sync at the end",
                    "sync_every_step": {"value": world * T * K / (ms_e2e_sync * 1e-3), "ms_per_step": ms_e2e_sync / K}},
            "gpu_launches": int(launches), "clocks": clocks,
            "status_flagged_tracks": status_bad, "pose_check": pose_check, "gathered_rows": int(gathered.shape[0]),
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(N)
        else:
            out["cpu_baseline"] = None
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
