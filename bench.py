#!/usr/bin/env python
"""bench.py -- benchmark of the hot path on the configurations BASELINE.json names.

Headline (`metric`, `value`, `e2e`, `roofline`): config 2 -- 4096 independent synthetic tracks x 500 slots per GPU,
15-component / 12-D PCA arm model, one shared measurement column per track-frame ("single-candidate update").
One "step" = one frame: ParticleFilter::update for every track (K->N indicator resample, fused per-slot KF predict +
innovation likelihood + KF update, weight normalisation, N->N systematic resample) followed by getEstimator + PCA
reconstruction; the timed region ends with the gather of per-track summaries (ncclAllGather through the C ABI).

Top-level legs of the same JSON line, each with ms_per_step, its own roofline {kernel, kernel_ms, frac} and clocks:
  config2_every_slot     config 2 with record sharing off (every slot computed: the roofline-facing figure)
  config2_literal_alias  config 2 in the reference binary's shallow-copy alias mode (quirk B3)
  config2_bank           config 2's secondary shape: 15 slots per track (one per component), the frame in one launch
  config3                16 384 persons x 500 slots x 17 candidates per hand: association + both arm updates
  config4                256 tracks x 65 536 slots, per-slot measurement columns (+ config4_pf2d: the legacy plain
                         filter at the same size -> particle likelihoods/s)
  config5                1 048 576 tracks x 15 slots (data23D model), STRONG-scaled: 1 M / N tracks per rank, 201 MB
                         of summaries gathered by NCCL; every rank checks gathered rows against a local recomputation

  python bench.py [--gpus N --steps K --warmup W]            B200 arm (one process per GPU under torchrun)
  python bench.py --impl reference [...]                      the reference's CPU path (oracle/) on host cores

Prints ONE JSON line (rank 0).  `value` = whole-job frame-updates/s with inputs resident in HBM; `e2e` = the same
work through the C ABI with pinned HOST buffers (H2D of the step's measurements and draws, D2H of the per-track pose)
inside the timed region.
"""
import argparse
import json
import os
import re
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
MODEL_DIR = os.path.join(ROOT, "mkfbodytracker_pdaf_b200", "models")
LEFT_YML = os.path.join(MODEL_DIR, "data13D_PCA_100000_15_12.yml")
RIGHT_YML = os.path.join(MODEL_DIR, "data23D_PCA_100000_15_12.yml")


def workload_name(T, N):
    return (f"config 2: {T} independent synthetic tracks/GPU x {N} slots, data13D_PCA_100000_15_12 (K=15, d=12), "
            "full GMM-KF bank + single-candidate (shared column) update")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe): one sampler runs for
    the whole process; window(t0, t1) summarises the samples that fell inside one timed region"""
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
                                          "-lms", os.environ.get("MKF_BENCH_SMI_MS", "20"), "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            for _ in range(150):  # the first sample takes nvidia-smi about a second
                if self.rows:
                    break
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if not self.proc:
            return
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.proc = None

    def defer(self, t0, t1):
        """a window to be summarised once sampling has stopped (resolve() replaces it in the result object)"""
        return ("__clock_window__", t0, t1)

    def resolve(self, obj):
        if isinstance(obj, tuple) and len(obj) == 3 and obj[0] == "__clock_window__":
            return self.window(obj[1], obj[2])
        if isinstance(obj, dict):
            return {k: self.resolve(v) for k, v in obj.items()}
        if isinstance(obj, list):
            return [self.resolve(v) for v in obj]
        return obj

    def window(self, t0, t1):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, pw = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for (ts, r) in self.rows if t0 <= ts <= t1 + 0.03]
        window = "timed region"
        if not inside:  # region shorter than the sampling period: the nearest samples on either side
            near = sorted(self.rows, key=lambda x: min(abs(x[0] - t0), abs(x[0] - t1)))[:4]
            inside = [r for (_, r) in near]
            window = "nearest samples (region shorter than the sampling period)"
        for r in inside:
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
    """dram bytes per launch of the slot kernels from the newest committed ncu --set full summary
    (profiles/slot_update_traffic.json; `captured` says which round / commit it belongs to)"""
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


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (the only code in this file that touches oracle/).  No product code is imported or mapped here: the model
# files are parsed by the few lines below (OpenCV-YAML-1.0 `!!opencv-matrix` blocks, as cv::FileStorage reads them at
# src/pfPose.cpp:34-55; f32 blocks are widened through float32 like src/pfPose.cpp:44-51 does).
# ---------------------------------------------------------------------------------------------------------------------
def read_opencv_yaml(path):
    import numpy as np
    txt = open(path).read()
    out = {}
    for m in re.finditer(r"^(\w+): !!opencv-matrix\s+rows: (\d+)\s+cols: (\d+)\s+dt: (\w)\s+data: \[(.*?)\]", txt, re.S | re.M):
        key, rows, cols, dt, data = m.group(1), int(m.group(2)), int(m.group(3)), m.group(4), m.group(5)
        vals = np.array([float(x) for x in data.replace("\n", " ").split(",") if x.strip()], dtype=np.float64)
        if dt == "f":
            vals = vals.astype(np.float32).astype(np.float64)
        out[key] = vals.reshape(rows, cols)
    return out


def model_arrays(path, gamma_path):
    import numpy as np
    y = read_opencv_yaml(path)
    g = read_opencv_yaml(gamma_path)["gamma"] if gamma_path else y["gamma"]  # quirk B4: src/pfPose.cpp:52-53
    K, d = y["means"].shape
    return dict(means=y["means"], covs=np.ascontiguousarray(y["covs"].reshape(K, d, d)), weights=y["weights"].reshape(-1),
                gamma=g.reshape(-1), pca_proj=y["pca_proj"], pca_mean=y["pca_mean"].reshape(-1))


CPU_DESC = {"reference": "oracle/_ref: the reference's own src/{KF_model,my_gmm,pf2DRao}.cpp on the OpenCV-subset shim "
                         "(literal cv::Mat aliasing, quirk B3)",
            "port": "oracle/mkf_oracle.cpp (C++ restatement, alias INDEPENDENT)"}


def cpu_arms():
    """{kind: run(T, N, frames, seed) -> (seconds, threads)} for the reference's CPU implementation of the path:
    'reference' = oracle/_ref (the reference's own sources compiled against oracle/cvshim) when it was built,
    'port' = the oracle restatement (always there)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    arrays = model_arrays(LEFT_YML, RIGHT_YML)
    arms = {}
    if os.environ.get("MKF_BENCH_CPU_KIND", "") != "port":
        try:
            import mkf_ref
            if os.path.exists(mkf_ref.SO):
                mkf_ref.lib()

                def run_ref(T, N, frames, seed):
                    return mkf_ref.bench_tracks(arrays, T, N, frames, per_slot=False, seed=seed, jitter=1,
                                                threads=host_cores())
                arms["reference"] = run_ref
        except Exception:
            pass
    import mkf_oracle as orc
    om = orc.Model(*[arrays[k] for k in ("means", "covs", "weights", "gamma", "pca_proj", "pca_mean")])

    def run_port(T, N, frames, seed):
        secs, used, _ = orc.bench_tracks(om, T, N, frames, per_slot=False, seed=seed, jitter=1, threads=host_cores())
        return secs, used
    arms["port"] = run_port
    return arms


def cpu_sample(run, N, budget_s):
    cores = host_cores()
    T_s = 8 * cores
    secs, used = run(T_s, N, 2, SEED)  # calibrate
    rate = T_s * 2 / max(secs, 1e-9)
    frames = int(max(2, min(400, budget_s * rate / T_s)))
    secs, used = run(T_s, N, frames, SEED)
    return T_s * frames / secs, used, T_s, frames, secs


def cpu_baseline(N, budget_s=10.0):
    """the reference CPU path on all host cores over a bounded sample of the same workload; the reference's own
    sources (kind "reference") when built, the port's rate beside it"""
    arms = cpu_arms()
    kind = "reference" if "reference" in arms else "port"
    fu, used, T_s, frames, secs = cpu_sample(arms[kind], N, budget_s)
    out = {"value": fu, "unit": UNIT, "cores": used, "kind": kind, "slot_updates_per_s": fu * N,
           "sample": f"{T_s} tracks x {N} slots x {frames} frames of the same synthetic workload ({CPU_DESC[kind]}; "
                     f"one track per task, single-threaded within a track, OpenMP over tracks; {secs:.2f} s)"}
    if kind == "reference":
        fu_p, used_p, T_p, fr_p, secs_p = cpu_sample(arms["port"], N, 4.0)
        out["port"] = {"value": fu_p, "unit": UNIT, "cores": used_p,
                       "sample": f"{T_p} tracks x {N} slots x {fr_p} frames, {CPU_DESC['port']}; {secs_p:.2f} s"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arms = cpu_arms()
    kind = "reference" if "reference" in arms else "port"
    run = arms[kind]
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
    port = None
    if kind == "reference":  # the restatement beside the reference's own sources (it is ~3x faster: no cv::Mat temporaries)
        fu_p, used_p, T_p, fr_p, secs_p = cpu_sample(arms["port"], N, 4.0)
        port = {"value": fu_p, "unit": UNIT, "cores": used_p,
                "sample": f"{T_p} tracks x {N} slots x {fr_p} frames, {CPU_DESC['port']}; {secs_p:.2f} s"}
    loaded = sorted({ln.split()[-1] for ln in open("/proc/self/maps") if ROOT in ln and ".so" in ln})
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(args.tracks, N), "sample": sample,
                      "alias_mode": "CV_SHALLOW_LITERAL (what the reference's sources compute; the B200 arm's headline "
                                    "runs INDEPENDENT and reports the literal mode as config2_literal_alias)"
                      if kind == "reference" else "INDEPENDENT",
                      "opencv": "oracle/cvshim restatement of the OpenCV-2.4 subset (unoptimised; a real OpenCV build "
                                "is not available): see `port` for the allocation-free restatement's rate"},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
           "port": port,
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "slot_updates_per_s": val * N, "gpu_launches": 0, "native_so_loaded": loaded}
    emit(out)


_JSON_FD = None


def reserve_stdout():
    """stdout carries exactly one JSON line: keep a private handle on it and point fd 1 at stderr for the rest of the
    run, so that nothing a library prints (NCCL's banner and INFO lines, for one) can land next to that line"""
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


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def timed_region(cx, fn, steps, warmup, tail=None, collective=True):
    """W untimed + K timed calls of fn(i), CUDA events on the launching stream, barrier + synchronize on both sides,
    MAX over ranks.  `tail` runs once inside the timed region after the last step.  collective=False: a leg that only
    rank 0 runs (no barrier, no reduction).  Returns (ms_total, t0, t1)."""
    torch, dist = cx.torch, cx.dist
    sync = cx.barrier if collective else torch.cuda.synchronize
    for i in range(warmup):
        fn(i)
    if tail:
        tail()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(warmup, warmup + steps):
        fn(i)
    if tail:
        tail()
    e1.record()
    sync()
    t1 = time.perf_counter()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=cx.dev)
    if cx.world > 1 and collective:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), t0, t1


def stage_means(prof):
    n = max(prof["n"], 1)
    out = {"indicator_bounds": prof["ms_bounds"] / n, "share_keys": prof["ms_share_keys"] / n,
           "slot_kernel": prof["ms_slot_kernel"] / n, "repair": prof["ms_repair"] / n,
           "normalise_resample": prof["ms_resample"] / n, "samples": prof["n"]}
    if prof.get("n_span"):  # the slot kernel's own span on the device (earliest CTA start .. latest CTA end)
        out["slot_kernel_device_span"] = prof["ms_slot_span"] / prof["n_span"]
    return out


def roofline_of(cx, kernel, units, kernel_ms, samples, traffic_key=None, extra=None, span_ms=None):
    ach = units * BYTES_PER_SLOT_UPDATE / (kernel_ms * 1e-3) / 1e9 if kernel_ms and kernel_ms > 0 else None
    tr = ((cx.traffic or {}).get("kernels") or {}).get(traffic_key or kernel) or {}
    out = {"bound": "hbm", "kernel": kernel, "kernel_ms": kernel_ms, "kernel_samples": samples,
           "units_per_launch": int(units), "algorithmic_bytes_per_launch": int(units) * BYTES_PER_SLOT_UPDATE,
           "achieved": ach, "peak": cx.peak, "unit": "GB/s", "frac": (ach / cx.peak) if ach else None,
           "traffic": tr.get("dram_bytes_per_launch"), "traffic_source": (cx.traffic or {}).get("captured")}
    if span_ms:
        # kernel_ms is the interval between two CUDA events, which on a sampled step includes the launch gap the event
        # records open in the programmatic-launch chain; the kernel's own span (first CTA start to last CTA end,
        # %globaltimer stamps taken by the kernel itself) is reported beside it
        out["kernel_ms_device_span"] = span_ms
        out["frac_device_span"] = units * BYTES_PER_SLOT_UPDATE / (span_ms * 1e-3) / 1e9 / cx.peak
    if extra:
        out.update(extra)
    return out


def leg_config2(cx, args):
    """headline: device-resident run, the same work end to end from pinned host buffers, and the two variants"""
    torch, mk = cx.torch, cx.mk
    from mkfbodytracker_pdaf_b200.sharding import gather_summaries_native, shard_tracks
    T, N, K, W = args.tracks, args.slots, args.steps, args.warmup
    F = K + W
    dev, world, rank = cx.dev, cx.world, cx.rank
    model = cx.left
    batch = mk.TrackBatch(model, T, N, device=cx.local, stream=cx.stream.cuda_stream)
    track0, _ = shard_tracks(world * T, world, rank)  # weak scaling: T tracks on every rank

    meas = torch.empty((F, T, 6), dtype=torch.float64, device=dev)
    ui = torch.empty((F, T), dtype=torch.float64, device=dev)
    up = torch.empty((F, T), dtype=torch.float64, device=dev)
    for f in range(F):
        batch.synth_fill(SEED, track0, f, 1, mk.MEAS_SHARED, meas[f], ui[f], up[f])
    u0 = torch.empty(T, dtype=torch.float64, device=dev)
    batch.synth_fill(SEED, track0, 0xFFFFFF, 1, mk.MEAS_SHARED, meas[0].clone(), u0, None)
    pose = torch.empty((T, model.D), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * T, model.D + 2), dtype=torch.float64, device=dev)

    def tail():  # final per-track summaries {pose[D], wsum, status}; the only collective is this gather
        gather_summaries_native(batch, cx.comm, gathered)

    def step(f):
        batch.update(meas[f], ui[f], up[f])
        batch.estimate_into(None, pose)

    # ---------------- device-resident run ----------------
    batch.reset(u0)
    # per-kernel CUDA events inside the timed region, on every PROF_EVERY-th step: a sampled step pays ~12 us for its
    # event records (and loses the kernels' programmatic overlap), so sampling all of them would tax the number they
    # are there to explain
    PROF_EVERY = 8
    for f in range(W):
        step(f)
    batch.profile((K + PROF_EVERY - 1) // PROF_EVERY, PROF_EVERY)
    launches0 = mk.launch_count()
    ms, t0, t1 = timed_region(cx, lambda i: step(W + i), K, 0, tail)
    launches = mk.launch_count() - launches0
    prof = batch.profile_read_stages()
    batch.profile(0)
    clocks = cx.clk.defer(t0, t1) if rank == 0 else None
    # repeat the timed region (same frames, fresh reset) and keep the median: K = 20 steps last ~3 ms
    reps = [ms]
    for _ in range(args.repeats - 1):
        batch.reset(u0)
        for f in range(W):
            step(f)
        r_ms, _, _ = timed_region(cx, lambda i: step(W + i), K, 0, tail)
        reps.append(r_ms)
    ms_med = sorted(reps)[len(reps) // 2]
    rec, nslots = batch.shared_records()  # distinct Gaussians the last frame stored (record sharing)
    heads_kernel = batch.heads_kernel()   # which kernel ran them (k_slot_update_heads_tma<...> unless MKF_HEADS_TMA=0)
    status_bad = int((batch.status() & (mk._lib.ST_POST_DEGENERATE | mk._lib.ST_CHOL_FAIL)).astype(bool).sum())
    rows_ok = bool(torch.equal(gathered[rank * T:(rank + 1) * T, :model.D], pose))
    pose_check = float(gathered[:, :2].mean())

    # ---------------- end to end through the C ABI with pinned host buffers ----------------
    # every frame's inputs sit in pinned HOST memory; each step copies them to the device, runs the frame and copies
    # the per-track pose back; the region ends with the same summary gather as above.  "sync_every_step" waits for the
    # pose after every step (one frame in flight: what a closed tracker loop sees, since the next association needs
    # this frame's estimate); `e2e` is the pipelined use of the same API (MKF_MEM_HOST_ASYNC: copies on the library's
    # copy streams, one synchronisation at the end of the timed region).
    h_meas = meas.cpu().pin_memory()
    h_ui = ui.cpu().pin_memory()
    h_up = up.cpu().pin_memory()
    h_pose = torch.empty((2, T, model.D), dtype=torch.float64).pin_memory()

    def e2e_run(mem):
        def e2e_step(f):
            batch.update(h_meas[f], h_ui[f], h_up[f], mem=mem)
            mk._lib.check(mk._lib.lib.mkf_batch_estimate(batch._h, None, h_pose[f & 1].data_ptr(), mem))

        def e2e_tail():
            batch.join()  # the closing event waits for the copies on the library's internal copy streams too
            tail()
        best = []
        for _ in range(args.repeats):
            batch.reset(u0)
            for f in range(W):
                e2e_step(f)
            t_ms, _, _ = timed_region(cx, lambda i: e2e_step(W + i), K, 0, e2e_tail)
            best.append(t_ms)
        return sorted(best)[len(best) // 2]

    ms_e2e_sync = e2e_run(mk.MEM_HOST)
    pose_sync = h_pose[(F - 1) & 1].clone()
    ms_e2e = e2e_run(mk.MEM_HOST_ASYNC)
    assert torch.equal(pose_sync, h_pose[(F - 1) & 1]), "pipelined and synchronous e2e runs must agree"
    batch.close()

    # ---------------- variants of the same frames on rank 0 (no collectives) ----------------
    def variant(model_v, env, K2, label_kernel, slots=None):
        N2 = slots or N
        for k, v in env.items():
            os.environ[k] = v  # read when a batch is created
        try:
            b2 = mk.TrackBatch(model_v, T, N2, device=cx.local, stream=cx.stream.cuda_stream)
        finally:
            for k in env:
                del os.environ[k]
        b2.reset(u0)

        def st2(f):
            b2.update(meas[f % F], ui[f % F], up[f % F])
            b2.estimate_into(None, pose)
        for f in range(W):
            st2(f)
        torch.cuda.synchronize()
        b2.profile((K2 + 3) // 4, 4)
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ta = time.perf_counter()
        ea.record()
        for f in range(W, W + K2):
            st2(f)
        eb.record()
        torch.cuda.synchronize()
        tb = time.perf_counter()
        p2 = b2.profile_read_stages()
        b2.profile(0)
        bad = int((b2.status() & (mk._lib.ST_POST_DEGENERATE | mk._lib.ST_CHOL_FAIL)).astype(bool).sum())
        b2.close()
        ms2 = ea.elapsed_time(eb) / K2
        sm = stage_means(p2)
        return {"steps": K2, "ms_per_step": ms2, "value": T * 1e3 / ms2, "unit": UNIT,
                "slot_updates_per_s": T * N2 * 1e3 / ms2, "stage_ms": sm, "status_flagged_tracks": bad,
                "roofline": roofline_of(cx, label_kernel, T * N2, sm["slot_kernel"], sm["samples"],
                                        span_ms=sm.get("slot_kernel_device_span")),
                "clocks": cx.clk.defer(ta, tb)}

    every_slot = literal = bank = None
    if rank == 0 and not args.headline_only:
        if os.environ.get("MKF_DEDUP", "1") != "0":
            every_slot = variant(model, {"MKF_DEDUP": "0"}, min(K, 48), "k_slot_update<12, 0>")
        prm = mk.default_params()
        prm.alias_mode = mk._lib.ALIAS_CV_SHALLOW_LITERAL
        a = model.arrays()
        lit = mk.Model.from_arrays(a["means"], a["covs"], a["weights"], a["gamma"], a["pca_proj"], a["pca_mean"], prm)
        literal = variant(lit, {}, min(K, 32), "k_slot_update<12, 1>")
        literal["note"] = ("MKF_ALIAS_CV_SHALLOW_LITERAL: slots that drew the same parent are filtered sequentially in "
                           "place (quirk B3, src/pf2DRao.cpp:153-156) -- what the reference binary and the CPU arm's "
                           "oracle/_ref compute; every slot is computed")

        # config 2's secondary shape: the plain GMM-KF bank, one slot per component
        bank = variant(model, {}, min(K, 100), "k_frame_small<12>" if T <= 16384 else "k_slot_update<12, 0>", slots=15)
        bank["workload"] = f"config 2 (bank): {T} tracks x 15 slots, shared column"
        bank["note"] = ("short tracks: the whole frame (indicator draw, slot update, resample, estimate) is ONE launch, half "
                        "a warp per track; launch / latency bound at this size -- the roofline entry is the frame "
                        "kernel's interval against 1500 B per slot-update")
    res = Ctx()
    res.__dict__.update(T=T, N=N, K=K, W=W, ms=ms_med, bank=bank, ms_first=ms, ms_reps=reps, ms_e2e=ms_e2e, ms_e2e_sync=ms_e2e_sync,
                        prof=prof, rec=rec, nslots=nslots, launches=launches, clocks=clocks, status_bad=status_bad,
                        rows_ok=rows_ok, pose_check=pose_check, gathered_rows=int(gathered.shape[0]),
                        every_slot=every_slot, literal=literal, prof_every=PROF_EVERY, model=model,
                        heads_kernel=heads_kernel)
    return res


def leg_config3(cx, steps=12, warmup=3):
    """config 3: association (gate, L/Z weights, C -> N resample) followed by both arm updates and the output
    estimates, 16 384 persons x 500 slots per arm x 17 candidates per hand (1 detection + 16 clutter)"""
    torch, mk = cx.torch, cx.mk
    T, N, Cn = 16384, 500, 17
    dev = cx.dev
    b0 = mk.TrackBatch(cx.left, T, N, cx.local, cx.stream.cuda_stream)
    b1 = mk.TrackBatch(cx.right, T, N, cx.local, cx.stream.cuda_stream)
    g = torch.Generator(device=dev)
    g.manual_seed(0x5EED0003)
    rnd = lambda *s: torch.rand(s, dtype=torch.float64, device=dev, generator=g)
    u0 = rnd(T)
    b0.reset(u0)
    b1.reset(u0)
    F = 4  # candidate sets cycled over the frames (resident in HBM)
    cand = torch.empty((F, T, 2, 2, Cn), dtype=torch.float64, device=dev)
    cand[:, :, :, 0] = rnd(F, T, 2, Cn) * 704 - 32
    cand[:, :, :, 1] = rnd(F, T, 2, Cn) * 528 - 24
    ph = rnd(2, T)
    for f in range(F):  # candidate 0 of each hand: the detection, near the moving hand
        hx = 388.0 + 60.0 * torch.sin(6.283185307179586 * (f / 75.0 + ph[0]))
        hy = 250.0 + 70.0 * torch.sin(6.283185307179586 * (f / 50.0 + ph[1]))
        nz = torch.randn((4, T), dtype=torch.float64, device=dev, generator=g) * 3
        cand[f, :, 0, 0, 0], cand[f, :, 0, 1, 0] = hx + nz[0], hy + nz[1]
        cand[f, :, 1, 0, 0], cand[f, :, 1, 1, 0] = hx - 140 + nz[2], hy + nz[3]
    Lv = torch.randint(1, 129, (F, T, 2, Cn), dtype=torch.uint8, device=dev, generator=g)
    Lv[torch.rand((F, T, 2, Cn), device=dev, generator=g) < 0.5] = 0
    Lv[:, :, :, 0] = torch.randint(200, 256, (F, T, 2), dtype=torch.uint8, device=dev, generator=g)
    roi = torch.tensor([300.0, 51.0, 47.0, 47.0], dtype=torch.float64, device=dev).repeat(T, 1).contiguous()
    us = [rnd(3, T, 2) for _ in range(F)]
    pose0 = torch.empty((T, cx.left.D), dtype=torch.float64, device=dev)
    pose1 = torch.empty((T, cx.right.D), dtype=torch.float64, device=dev)

    def assoc_only(i):
        f = i % F
        mk.associate(b0, b1, cand[f], Lv[f], roi, us[f][0], None, None, do_update=False)

    def frame(i):  # what a tracker loop runs per frame (src/pfPose.cpp:332-351)
        f = i % F
        mk.associate(b0, b1, cand[f], Lv[f], roi, us[f][0], us[f][1], us[f][2], do_update=True)
        b0.estimate_into(None, pose0)
        b1.estimate_into(None, pose1)

    for i in range(warmup):
        frame(i)
    ms_a, _, _ = timed_region(cx, assoc_only, steps, 1, collective=False)
    b0.profile((steps + 1) // 2, 2)
    ms_f, t0, t1 = timed_region(cx, lambda i: frame(warmup + i), steps, 0, collective=False)
    p0 = b0.profile_read_stages()
    b0.profile(0)
    rec, nslots = b0.shared_records()
    hk3 = b0.heads_kernel() or "k_slot_update_heads_direct<12>"
    st = b0.status() | b1.status()
    bad = int(((st & (mk._lib.ST_POST_DEGENERATE | mk._lib.ST_CHOL_FAIL | mk._lib.ST_CAND_DEGENERATE)) != 0).sum())
    b0.close()
    b1.close()
    sm = stage_means(p0)
    sharing = rec < nslots
    ms_step = ms_f / steps
    return {"workload": f"config 3: {T} persons x 2 arms x {N} slots, {Cn} candidates per hand (1 detection + 16 clutter), "
                        "association + both arm updates + output estimates per step; inputs resident",
            "steps": steps, "ms_per_step": ms_step, "person_frames_per_s": T * 1e3 / ms_step,
            "value": 2 * T * 1e3 / ms_step, "unit": UNIT,
            "nominal_slot_updates_per_s": 2 * T * N * 1e3 / ms_step,
            "assoc_only_ms": ms_a / steps, "candidate_weights_per_s": T * 2 * Cn / (ms_a / steps) * 1e3,
            "gate_decisions_per_s": T * 2 * Cn / (ms_a / steps) * 1e3,
            "distinct_records_fraction": rec / max(nslots, 1), "stage_ms_left_arm": sm, "status_flagged_persons": bad,
            "roofline": roofline_of(cx, hk3 + " (left arm)" if sharing else "k_slot_update<12, 0>",
                                    rec if sharing else nslots, sm["slot_kernel"], sm["samples"],
                                    traffic_key=None if sharing else "k_slot_update",
                                    span_ms=sm.get("slot_kernel_device_span")),
            "clocks": cx.clk.defer(t0, t1)}


def leg_config4(cx, steps=8, warmup=3):
    """config 4: 256 tracks x 65 536 slots, per-slot measurement columns (every slot computed), and the legacy plain
    particle filter (src/pf2D.cpp) at the same T x N for "particle likelihoods/s" """
    torch, mk, np = cx.torch, cx.mk, cx.np
    T, N, seed = 256, 65536, 0x5EED0004
    dev = cx.dev
    b = mk.TrackBatch(cx.left, T, N, cx.local, cx.stream.cuda_stream)
    F = 3  # measurement sets cycled over the frames (805 MB each, resident)
    meas = torch.empty((F, T, 6, N), dtype=torch.float64, device=dev)
    ui = torch.empty((F, T), dtype=torch.float64, device=dev)
    up = torch.empty((F, T), dtype=torch.float64, device=dev)
    for f in range(F):
        b.synth_fill(seed, 0, f, 0, mk.MEAS_PER_SLOT, meas[f], ui[f], up[f])
    u0 = torch.empty(T, dtype=torch.float64, device=dev)
    b.synth_fill(seed, 0, 0xFFFFFF, 0, mk.MEAS_SHARED, torch.empty((T, 6), dtype=torch.float64, device=dev), u0, None)
    pose = torch.empty((T, cx.left.D), dtype=torch.float64, device=dev)
    b.reset(u0)

    def step(i):
        f = i % F
        b.update(meas[f], ui[f], up[f])
        b.estimate_into(None, pose)
    for i in range(warmup):
        step(i)
    b.profile(steps, 1)
    ms, t0, t1 = timed_region(cx, lambda i: step(warmup + i), steps, 0, collective=False)
    p = b.profile_read_stages()
    b.profile(0)
    st = b.status()
    fb = int(((st & mk._lib.ST_POST_FALLBACK) != 0).sum())
    bad = int(((st & (mk._lib.ST_POST_DEGENERATE | mk._lib.ST_CHOL_FAIL)) != 0).sum())
    b.close()
    del meas
    sm = stage_means(p)
    ms_step = ms / steps
    out4 = {"workload": f"config 4: {T} tracks x {N} slots, per-slot measurement columns, per-slot KF + systematic "
                        "resampling; inputs resident (3 measurement sets cycled)",
            "steps": steps, "ms_per_step": ms_step, "value": T * 1e3 / ms_step, "unit": UNIT,
            "slot_updates_per_s": T * N * 1e3 / ms_step, "stage_ms": sm, "status_flagged_tracks": bad,
            "literal_loop_tracks_last_frame": fb,
            "roofline": roofline_of(cx, "k_slot_update<12, 0>", T * N, sm["slot_kernel"], sm["samples"],
                                    traffic_key="k_slot_update", span_ms=sm.get("slot_kernel_device_span")),
            "clocks": cx.clk.defer(t0, t1)}

    # legacy plain particle filter, d = 8, K = 15 synthetic SPD GMM (no model file for it ships)
    d, K = 8, 15
    rng = np.random.default_rng(2)
    means = rng.uniform(100, 400, (K, d))
    covs = np.stack([40 * (a @ a.T + d * np.eye(d)) for a in rng.standard_normal((K, d, d))])
    wts = rng.dirichlet(np.ones(K))
    pf = mk.Pf2dBatch(T, N, means, covs, wts, cx.local, cx.stream.cuda_stream)
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    parts = torch.tensor(means, device=dev)[torch.randint(0, K, (T, N), device=dev, generator=g)]
    parts = parts + torch.randn((T, N, d), dtype=torch.float64, device=dev, generator=g) * 6
    pf.set_particles(parts.contiguous())
    zm = torch.stack([parts[:, :, 6].mean(1), parts[:, :, 7].mean(1), parts[:, :, 0].mean(1), parts[:, :, 1].mean(1)],
                     1).reshape(T, 2, 2).contiguous()
    del parts
    u = torch.rand(T, dtype=torch.float64, device=dev, generator=g)
    noise = torch.randn((T, N, d), dtype=torch.float64, device=dev, generator=g)

    def pstep(i):
        pf.update(zm, u, noise)
    for i in range(warmup):
        pstep(i)
    pf.profile(steps)
    msp, t0, t1 = timed_region(cx, pstep, steps, 0, collective=False)
    pp = pf.profile_read()
    pf.profile(0)
    pf.close()
    n = max(pp["n"], 1)
    w_ms = pp["ms_weights"] / n
    msp_step = msp / steps
    # per particle: d*8 B read + 8 B weight written by the weight kernel; K*(d*d + d) FMA-equivalents of f64 arithmetic
    flops = T * N * K * (2 * d * d + 3 * d)
    outp = {"workload": f"legacy plain particle filter (src/pf2D.cpp:148-268): {T} filters x {N} particles, d={d}, "
                        f"K={K}: GMM prior with float expf x two isotropic likelihoods, normalise, resample, predict",
            "steps": steps, "ms_per_step": msp_step, "particle_likelihoods_per_s": T * N * 1e3 / msp_step,
            "unit": "particle likelihoods/s",
            "stage_ms": {"weights": w_ms, "normalise_resample": pp["ms_resample"] / n,
                         "gather_predict": pp["ms_predict"] / n, "samples": pp["n"]},
            "weight_kernel_likelihoods_per_s": T * N * 1e3 / w_ms if w_ms > 0 else None,
            "roofline": {"bound": "hbm", "kernel": "k_pf2d_weight<8>", "kernel_ms": w_ms, "kernel_samples": pp["n"],
                         "algorithmic_bytes_per_launch": T * N * 72,
                         "achieved": T * N * 72 / (w_ms * 1e-3) / 1e9 if w_ms > 0 else None, "peak": cx.peak,
                         "unit": "GB/s", "frac": T * N * 72 / (w_ms * 1e-3) / 1e9 / cx.peak if w_ms > 0 else None,
                         "traffic": None,
                         "co_bound": {"pipe": "fp64", "gflops": flops / (w_ms * 1e-3) / 1e9 if w_ms > 0 else None,
                                      "note": "the reference's unfused mul/add order is kept (bit-exact weights), so the "
                                              "kernel is FP64-issue bound, not HBM bound: ncu sm__inst_executed_pipe_fp64 "
                                              "in profiles/"}},
            "clocks": cx.clk.defer(t0, t1)}
    return out4, outp


def leg_config5(cx, steps=12, warmup=3):
    """config 5: 1 048 576 tracks x 15 slots, data23D model, tracks block-partitioned over the ranks (STRONG scaling),
    summaries gathered by ncclAllGather (201 MB per rank); every rank checks a sample of gathered rows of EVERY shard
    against its own recomputation of those tracks"""
    torch, mk, dist = cx.torch, cx.mk, cx.dist
    from mkfbodytracker_pdaf_b200.sharding import gather_summaries_native, shard_tracks
    total, N, seed = 1 << 20, 15, 0x5EED0005
    dev, world, rank = cx.dev, cx.world, cx.rank
    first, T = shard_tracks(total, world, rank)
    rows = max(shard_tracks(total, world, r)[1] for r in range(world))
    model = cx.right
    D = model.D
    b = mk.TrackBatch(model, T, N, cx.local, cx.stream.cuda_stream)
    F = warmup + steps
    meas = torch.empty((F, T, 6), dtype=torch.float64, device=dev)
    ui = torch.empty((F, T), dtype=torch.float64, device=dev)
    up = torch.empty((F, T), dtype=torch.float64, device=dev)
    for f in range(F):
        b.synth_fill(seed, first, f, 1, mk.MEAS_SHARED, meas[f], ui[f], up[f])
    u0 = torch.empty(T, dtype=torch.float64, device=dev)
    b.synth_fill(seed, first, 0xFFFFFF, 1, mk.MEAS_SHARED, meas[0].clone(), u0, None)
    pose = torch.empty((T, D), dtype=torch.float64, device=dev)
    gathered = torch.zeros((world * rows, D + 2), dtype=torch.float64, device=dev)
    b.reset(u0)

    def step(f):
        b.update(meas[f], ui[f], up[f])
        b.estimate_into(None, pose)

    def tail():
        gather_summaries_native(b, cx.comm, gathered, rows)
    for f in range(warmup):
        step(f)
    b.profile((steps + 1) // 2, 2)
    ms, t0, t1 = timed_region(cx, lambda i: step(warmup + i), steps, 0, tail)
    p = b.profile_read_stages()
    b.profile(0)
    # the gather alone (its share of the timed region)
    cx.barrier()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    tail()
    eb.record()
    cx.barrier()
    ms_g = torch.tensor([ea.elapsed_time(eb)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_g, op=dist.ReduceOp.MAX)
    bad = int(((b.status() & (mk._lib.ST_POST_DEGENERATE | mk._lib.ST_CHOL_FAIL)) != 0).sum())
    b.close()
    del meas
    # cross-rank check: 8 tracks at the start, middle and end of every rank's shard, recomputed here from reset in
    # small batches over the same frames, against the rows that rank delivered (bit-exact: tracks are independent and
    # the arithmetic of a track does not depend on which batch holds it)
    mism = 0
    checked = 0
    for r in range(world):
        f_r, n_r = shard_tracks(total, world, r)
        for off in (0, max(0, n_r // 2 - 4), max(0, n_r - 8)):
            g0 = f_r + off
            nt = min(8, n_r - off)
            sb = mk.TrackBatch(model, nt, N, cx.local, cx.stream.cuda_stream)
            m_s = torch.empty((nt, 6), dtype=torch.float64, device=dev)
            a_s = torch.empty(nt, dtype=torch.float64, device=dev)
            c_s = torch.empty(nt, dtype=torch.float64, device=dev)
            sb.synth_fill(seed, g0, 0xFFFFFF, 1, mk.MEAS_SHARED, m_s, a_s, None)
            sb.reset(a_s.clone())
            for f in range(F):
                sb.synth_fill(seed, g0, f, 1, mk.MEAS_SHARED, m_s, a_s, c_s)
                sb.update(m_s, a_s, c_s)
            want = torch.empty((nt, D + 2), dtype=torch.float64, device=dev)
            mk._lib.check(mk._lib.lib.mkf_batch_summaries(sb._h, nt, want.data_ptr(), mk.MEM_DEVICE))
            sb.sync()
            got = gathered[r * rows + off: r * rows + off + nt]
            mism += int((got != want).any(dim=1).sum())
            checked += nt
            sb.close()
    flag = torch.tensor([mism], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    sm = stage_means(p)
    ms_step = ms / steps
    return {"workload": f"config 5: {total} tracks x {N} slots (data23D_PCA_100000_15_12), shared column, tracks "
                        f"block-partitioned over {world} rank(s) ({T} on rank 0), {steps} frames + ONE ncclAllGather of "
                        f"{world * rows} x {D + 2} f64 summary rows ({world * rows * (D + 2) * 8 / 1e6:.0f} MB per rank) "
                        "inside the timed region",
            "scaling": "strong", "n_gpus": world, "steps": steps, "ms_per_step": ms_step,
            "value": total * 1e3 / ms_step, "unit": UNIT, "slot_updates_per_s": total * N * 1e3 / ms_step,
            "gather_ms": float(ms_g.item()), "gather_bytes_per_rank": world * rows * (D + 2) * 8,
            "gathered_rows": world * rows, "cross_rank_rows_checked_per_rank": checked,
            "cross_rank_mismatches_max_over_ranks": int(flag.item()), "cross_rank_check": "ok" if flag.item() == 0 else "FAILED",
            "stage_ms": sm, "status_flagged_tracks": bad,
            "roofline": roofline_of(cx, "k_slot_update<12, 0>", T * N, sm["slot_kernel"], sm["samples"],
                                    traffic_key="k_slot_update", span_ms=sm.get("slot_kernel_device_span")),
            "clocks": cx.clk.defer(t0, t1)}


def main():
    reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tracks", type=int, default=4096, help="tracks per GPU (config 2)")
    ap.add_argument("--slots", type=int, default=500)
    ap.add_argument("--repeats", type=int, default=5, help="repetitions of the timed region (median reported)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="config 2 only (skip the other configs' legs)")
    ap.add_argument("--legs", default="3,4,5", help="which of the other configs' legs to run (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist

    import mkfbodytracker_pdaf_b200 as mk
    from mkfbodytracker_pdaf_b200.sharding import Comm

    cx = Ctx()
    cx.torch, cx.dist, cx.mk, cx.np = torch, dist, mk, np
    cx.rank = rank = int(os.environ.get("RANK", "0"))
    cx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    if mk.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (libmkf_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    cx.dev = dev = torch.device("cuda", local)
    # NCCL's own account of the communicators goes to stderr (fd 1 points there for the whole run, so its lines cannot
    # touch the JSON line): INIT-level INFO unless the caller chose a level
    if world > 1 and os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
        os.environ["NCCL_DEBUG"] = os.environ.get("MKF_NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
    if world > 1:
        import datetime
        # (a short collective timeout: a rank that dies must not leave the others waiting for ten minutes)
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
    # everything (our kernels, NCCL, the timing events) runs on ONE explicit stream: torch's default stream has
    # handle 0, which the C ABI reads as "make a private stream"
    cx.stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(cx.stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    cx.barrier = barrier
    cx.peak, cx.peak_src = load_peak()
    cx.traffic = load_traffic()
    cx.left = mk.Model.load(mk.LEFT_ARM_MODEL, mk.RIGHT_ARM_MODEL)
    cx.right = mk.Model.load(mk.RIGHT_ARM_MODEL, mk.RIGHT_ARM_MODEL)
    # the C ABI's own communicator (ncclCommInitRank through mkf_comm_create); torch.distributed only carries the id
    cx.comm = Comm.from_torch_distributed(local) if world > 1 else Comm(1, 0, local, Comm.unique_id())
    nccl_version = cx.comm.nccl_version()
    cx.clk = ClockSampler(local)
    if rank == 0:
        cx.clk.start()

    legs = set() if args.headline_only else set(args.legs.split(","))
    c2 = leg_config2(cx, args)
    # config 5 runs up to the 100 frames BASELINE names (as many as --steps asks for), then gathers once
    c5 = leg_config5(cx, steps=max(4, min(args.steps, 100))) if "5" in legs else None
    c3 = c4 = pf = None
    if rank == 0:  # single-GPU legs (independent persons: they scale like config 2)
        if "3" in legs:
            c3 = leg_config3(cx)
        if "4" in legs:
            c4, pf = leg_config4(cx)
    if rank == 0:
        cx.clk.stop()
    barrier()

    if rank == 0:
        T, N, K, W = c2.T, c2.N, c2.K, c2.W
        model = c2.model
        value = world * T * K / (c2.ms * 1e-3)
        e2e_val = world * T * K / (c2.ms_e2e * 1e-3)
        sm = stage_means(c2.prof)
        slot_ms = sm["slot_kernel"]
        sharing = c2.rec < c2.nslots
        # units one launch processes: the distinct Gaussians (records) when identical children are shared -- the last
        # frame's count, stationary after the first ~20 frames -- else every slot
        units = c2.rec if sharing else c2.nslots
        split = sharing and sm["share_keys"] > 0
        kname = ((c2.heads_kernel or "k_slot_update_heads_direct") if split else "k_slot_update_shared") if sharing \
            else "k_slot_update"
        roof = roofline_of(cx, kname, units, slot_ms, sm["samples"], traffic_key=kname.split("<")[0],
                           span_ms=sm.get("slot_kernel_device_span"), extra={
            "kernel_sampling": f"CUDA events on every {c2.prof_every}th step of the timed region",
            "slots_per_launch": int(c2.nslots), "peak_source": cx.peak_src,
            "distinct_records_fraction": c2.rec / max(c2.nslots, 1),
            "kernel_share_of_step": slot_ms / (c2.ms / K) if c2.ms > 0 else None,
            "note": "SURVEY 8(d): 1500 B per slot-update.  With one measurement per track, children that drew the same "
                    "parent record and component are identical Gaussians and are computed and stored once: a launch "
                    "processes units_per_launch DISTINCT slot-updates for slots_per_launch slots; `achieved` counts only "
                    "those.  config2_every_slot is the same workload with the sharing off (1500 B for every slot)"})
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": c2.ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(T, N), "tracks_per_gpu": T, "slots": N, "components": model.K,
                       "state_dim": model.d, "measurement": "shared column per track-frame", "chol_mode": "CV24_LITERAL",
                       "alias_mode": "INDEPENDENT", "seed": hex(SEED),
                       "l2": f"no flush: per-step working set {2 * T * N * 720 / 1e9:.2f} GB >> 126 MB L2",
                       "timed_region": f"{K} frames (update + estimate) + the summary gather, median of {args.repeats} "
                                       "repetitions of the region",
                       "parallelism": f"tracks sharded, {world} x {T}; one final ncclAllGather of per-track summaries "
                                      "issued by the C ABI (mkf_batch_gather_summaries)",
                       "reference_arm_differs": "the CPU arm (--impl reference) times oracle/_ref on a bounded 8-tracks-"
                                                "per-core sample in the reference's literal alias mode; this arm's "
                                                "headline is alias INDEPENDENT with identical children stored once "
                                                "(config2_literal_alias / config2_every_slot are the like-for-like legs)"},
            "ms_per_step_repeats": [m / K for m in c2.ms_reps],
            # slot-level accounting: `computed` = distinct slot-updates the kernels executed, `nominal` = N per
            # frame-update (what N independent slots would be; every per-slot output is delivered bit-identically)
            "computed_slot_updates_per_s": world * units * K / (c2.ms * 1e-3),
            "nominal_slot_updates_per_s": value * N,
            "roofline": roof,
            "stage_ms": sm,
            "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": c2.ms_e2e / K,
                    "h2d_bytes_per_step": world * T * 8 * 8, "d2h_bytes_per_step": world * T * model.D * 8,
                    "mode": "pinned host buffers, MKF_MEM_HOST_ASYNC: copies on the library's copy streams overlap the "
                            "neighbouring frames' kernels, one sync at the end; same work as `value` (frames + gather)"},
            "e2e_sync_every_step": {"value": world * T * K / (c2.ms_e2e_sync * 1e-3), "unit": UNIT,
                                    "ms_per_step": c2.ms_e2e_sync / K,
                                    "mode": "MKF_MEM_HOST: the caller waits for each frame's pose before the next frame "
                                            "(a closed tracker loop)"},
            "gpu_launches": int(c2.launches), "clocks": c2.clocks,
            "status_flagged_tracks": c2.status_bad, "pose_check": c2.pose_check, "gathered_rows": c2.gathered_rows,
            "gathered_rows_match_local": c2.rows_ok,
            "nccl": {"version": nccl_version, "nranks": world, "debug": os.environ.get("NCCL_DEBUG"),
                     "comm": "mkf_comm_create (ncclCommInitRank) + torch.distributed process group"},
            "config2_every_slot": c2.every_slot, "config2_literal_alias": c2.literal, "config2_bank": c2.bank,
            "config3": c3, "config4": c4, "config4_pf2d": pf, "config5": c5,
            "particle_likelihoods_per_s": pf["particle_likelihoods_per_s"] if pf else None,
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(N)
        else:
            out["cpu_baseline"] = None
        emit(cx.clk.resolve(out))
    cx.comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
