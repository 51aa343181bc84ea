#!/usr/bin/env python
"""bench.py -- lattice site-updates/s of the Rietkerk Dornic+RK4 step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU integrator

A "step" is one full Dornic time step (vegetation Dornic sampling + RK4 over P,W,O + clamp +
reductions) of every site of the workload lattice(s).  N=1 workload: Rietkerk SPB=1, L=4096,
one trial per GPU (the configuration the metric and the 60 %-of-HBM target are quoted on);
N>1 shards independent trials over ranks with no data-path collective ("weak").

value  : whole-job site-updates/s, state resident in HBM, device-timed (CUDA events on the
         kernel's stream), max over ranks.
e2e    : same metric through tm_run_host(): every step copies the step's input frame from pinned
         host memory to the device and reads the step's reductions back.
roofline: algorithmic bytes (48 B per site-update: 3 fields x (8 read + 8 written)) / the average
         duration of one step, against the measured HBM copy bandwidth.  A step is two launches (the vegetation
         Dornic sweep, then the RK4 marching kernel); `roofline.kernels` times each of them with CUDA events
         recorded between the launches in a second, separate pass (tm_set_kernel_timing), with its own
         algorithmic bytes (16 B and 48 B per site: the pair moves 64 B where a fused kernel would move 48).
other_configs / slab_L16384: the other BASELINE configurations beside the headline, measured by the same run after the headline
         (configs 2, 3, 4 as one GPU's share each, weak; config 5 as ONE L=16384 lattice over the N ranks, strong), each under a
         watchdog so that the headline line is printed whatever happens to them.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BYTES_PER_SITE_UPDATE = 48.0  # Rietkerk SPB=1, 3 fields x (8 B read + 8 B written)
METRIC = "lattice site-updates/sec (FP64, device-timed), Rietkerk Dornic+RK4 step"
UNIT = "site-updates/s"
PEER_FAILED = 1.0e300  # what a rank that failed inside an extra section contributes to its peers' MAX all-reduce


class PeerFailed(RuntimeError):
    pass


def run_guarded(out, rank, dist, key, budget_s, fn, keep=None):
    """One extra section of the default run (the headline in `out` is already measured).  `fn` runs under a watchdog and a `try`:
    its result (rank 0: the whole dict, or the keys in `keep`) goes to out[key]; an exception is reported there as {"error": ...}; a
    section that does not finish within `budget_s` seconds is reported as skipped, the ONE JSON line is printed by rank 0 and every
    rank leaves (exit code 0).  Returns False when the section failed on this rank.  With several ranks a failing rank contributes
    PEER_FAILED to the one-element MAX all-reduce its peers are in (the only collective an extra section may use), so that they fail
    at the same point instead of waiting for it."""
    def give_up():
        if rank == 0:
            out[key] = {"error": "this section did not finish within %d s; skipped" % budget_s}
            print(json.dumps(out), flush=True)
        os._exit(0)
    dog = threading.Timer(budget_s, give_up)
    dog.daemon = True
    dog.start()
    ok = True
    try:
        r = fn()
        if rank == 0:
            out[key] = r if keep is None else {k: r[k] for k in keep}
    except Exception as e:   # the headline must still be printed
        ok = False
        if rank == 0:
            out[key] = {"error": "%s: %s" % (type(e).__name__, e)}
        if dist is not None and not isinstance(e, PeerFailed):
            try:
                import torch
                t = torch.tensor([PEER_FAILED], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            except Exception:
                pass
    dog.cancel()
    return ok


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_lines(sites, ms_dornic, ms_rk4, n, peak):
    """Per-kernel roofline lines of the pair: algorithmic bytes of each kernel / its own average duration."""
    out = []
    for name, ms, bps in (("tml::riet_dornic_kernel<STRICT=0>", ms_dornic, 16.0),
                          ("tml::riet_step_kernel<DET=0,STRICT=0,SRC=pre,DORNIC_ONLY=0>", ms_rk4, 48.0)):
        if n > 0 and ms > 0:
            gbs = sites * bps / (ms * 1e-3 / n) / 1e9
            out.append({"kernel": name, "ms_per_launch": ms / n, "algorithmic_bytes_per_site": bps,
                        "achieved": gbs, "frac": gbs / peak, "launches_timed": n})
    return out


def ncu_traffic_per_launch(L, trials, with_source=False):
    """dram bytes per launch from the committed ncu --set full capture, if it matches this workload (with_source: and which build
    the capture was taken of -- the number is not measured by this run)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            d = json.load(fh)
        if d.get("L") == L and d.get("trials") == trials:
            return (float(d["dram_bytes_per_launch"]), d.get("source")) if with_source else float(d["dram_bytes_per_launch"])
    except Exception:
        pass
    return (None, None) if with_source else None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for (_, l) in self.lines]
        for l in inside:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU reference arm: the reference's own integrator (oracle/_ref) or, failing that, the oracle port
# ---------------------------------------------------------------------------------------------
def _ref_worker(args):
    """One process = one a-value, like one OpenMP thread of the reference's sweep (R/:4519)."""
    L, steps, a, seed = args
    import ctypes as C
    import numpy as np
    import oracle_lib as O
    os.chdir(tempfile.mkdtemp(prefix="tmref_"))
    frame = np.ascontiguousarray(O.synthetic_rietkerk_frame(L, seed=seed, a=a, fringe=False).reshape(3, -1))
    so = ref_riet_so()
    dt = 0.1
    if os.path.exists(so):
        ref = C.CDLL(so)

        class RP(C.Structure):
            _fields_ = [(n, C.c_double) for n in "a c gmax alpha rW W0 d K1 K2 D0 D1 D2 sigma0 dt dx".split()]
        D0 = 0.00025 / 24
        rp = RP(a, 1e4, 0.05e-3 / 24, 0.2 / 24, 0.2 / 24, 0.2, 0.25 / 24, 5.0, 5000.0, D0, D0, 0.025 / 24,
                (D0 / 0.01) ** 0.5, dt, 0.1)
        dp = C.POINTER(C.c_double)

        def run(n):
            tm_ = np.array([0.0, dt * (n - 1)])
            ft = np.array([1e300, 1e300])
            rows = np.zeros((8, 7)); nfr = C.c_int()
            t0 = time.perf_counter()
            ref.ref_riet_run(C.byref(rp), L, 1, frame.ctypes.data_as(dp), tm_.ctypes.data_as(dp), 2,
                             ft.ctypes.data_as(dp), 2, None, C.c_uint64(0), None, None, 0, C.byref(nfr),
                             rows.ctypes.data_as(dp), 8, None, None)
            return time.perf_counter() - t0
        kind = "reference"
    else:
        import trophic_maelstorm_b200 as tm
        op = O.to_oparams(tm.rietkerk_params(L, a=a))

        def run(n):
            f = frame.copy()
            t0 = time.perf_counter()
            for s in range(n):
                O.step(op, O.make_rng(O.SRC_PHILOX, seed=seed, trial=0, step=s), f)
            return time.perf_counter() - t0
        kind = "port"
    # two run lengths: the difference removes set-up (neighbour tables, frame copies)
    t_short = run(1)
    t_long = run(1 + steps)
    return kind, max(t_long - t_short, 1e-9)


def ref_riet_so():
    """The unmodified reference compiled here by oracle/Makefile (it cannot be rebuilt on the GPU host: /root/reference
    does not travel).  Two builds travel: -march=x86-64-v4 (AVX-512; what -march=native gives on the Xeon hosts of this
    pool) and -march=x86-64-v3; the best one this host can execute is used."""
    d = os.path.join(ROOT, "oracle", "_ref")
    try:
        with open("/proc/cpuinfo") as fh:
            flags = fh.read()
    except Exception:
        flags = ""
    v4 = os.path.join(d, "libtm_ref_riet_v4.so")
    if os.path.exists(v4) and all(f in flags for f in ("avx512f", "avx512bw", "avx512vl", "avx512dq", "avx512cd")):
        return v4
    return os.path.join(d, "libtm_ref_riet.so")


def ref_procs_for(L, cores):
    """How many L x L reference runners fit in host memory at once (the runner holds ~21 fields plus nested-vector
    neighbour tables: ~390 B per site, measured 1.6 GB at L = 2048)."""
    need = 400.0 * L * L + 3e8
    try:
        import psutil
        avail = float(psutil.virtual_memory().available)
        try:
            with open("/sys/fs/cgroup/memory.max") as fh:
                lim = fh.read().strip()
            if lim != "max":
                avail = min(avail, float(lim) * 0.9)
        except Exception:
            pass
    except Exception:
        avail = 16e9
    return int(max(1, min(cores, (0.6 * avail) // need)))


def cpu_reference(L, steps, procs):
    """site-updates/s of the CPU integrator on `procs` cores (one a-value per core)."""
    import multiprocessing as mp
    procs = ref_procs_for(L, procs)
    ctx = mp.get_context("spawn")
    jobs = [(L, steps, 0.05 + 0.001 * i, 100 + i) for i in range(procs)]
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    kind = res[0][0]
    slowest = max(r[1] for r in res)
    rate = procs * L * L * steps / slowest
    return {"value": rate, "unit": UNIT, "cores": procs, "kind": kind,
            "sample": "Rietkerk SPB=1 L=%d lattice per core, %d Dornic+RK4 steps each, %d a-values in parallel "
                      "(the reference parallelises over a-values only, R/:4509; as many as fit in host memory); set-up excluded; "
                      "wall %.1f s; %s (g++ -O3, no -fopenmp: g++ 13 rejects R/:2621 under it; built in the build container, "
                      "/root/reference does not travel to the GPU host)" % (L, steps, procs, wall, os.path.basename(ref_riet_so()) if kind == "reference" else "oracle port")}


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def workload_config(L, trials):
    """The `config` object: the same in both arms (the reference arm times THIS workload, see cpu_baseline.sample)."""
    return {"workload": "Rietkerk SPB=1 (veg/soil-water/surface-water) Dornic+RK4 step, L=%d, %d trial(s) per GPU, "
                        "a=0.05 dt=0.1, MFT speckle initial frame" % (L, trials),
            "parallelism": "independent trials per rank, no data-path collective",
            "l2": ("state %.0f MiB per GPU (x2 ping-pong) exceeds the 126 MB L2; no flush needed" if trials * 3 * L * L * 8 > 126e6 else
                   "state %.0f MiB per GPU fits the 126 MB L2 (not the headline configuration); no flush") % (trials * 3 * L * L * 8 / 2**20)}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    steps = max(1, a.steps)
    # Same configuration as the product arm: L x L lattices of the named L (one per host core that fits in memory, the
    # reference's own a-value parallelism), `steps` time steps each after the runner's first (warm-up) step; the set-up
    # (neighbour tables, frame copies) is excluded by differencing two run lengths.  At L = 4096 a step costs ~2 s per core.
    ref_L = a.ref_L or a.L
    base = cpu_reference(ref_L, steps, cores)
    v = base["value"]
    out = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": steps, "warmup": a.warmup,
           "ms_per_step": 1e3 * a.L * a.L * a.trials * a.gpus / v, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
           "config": workload_config(a.L, a.trials),
           "cpu_baseline": base, "gpu_launches": 0,
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(a):
    import numpy as np
    import torch
    import oracle_lib as O          # only for the synthetic frame rule and the cpu_baseline leg
    import trophic_maelstorm_b200 as tm

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    L, trials, K, W = a.L, a.trials, a.steps, max(a.warmup, 3)
    p = tm.rietkerk_params(L)
    lat = tm.Lattice(p, trials, device=local)
    lat.seed(20261017, trial0=rank * trials)
    host = torch.empty((trials, 3, L, L), dtype=torch.float64, pin_memory=True)
    hnp = host.numpy()
    for t in range(trials):
        hnp[t] = O.synthetic_rietkerk_frame(L, seed=1234 + rank * trials + t, fringe=False)
    lat.upload(hnp)

    def barrier():
        torch.cuda.synchronize()
        lat.sync()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def maxr(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lat.step(W)
    lat.sync()

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    barrier()
    t0 = time.time()
    lat.step(K)                      # K steps = 2K launches (Dornic sweep + RK4 kernel) on the ctx stream
    lat.sync()
    t1 = time.time()
    ms_dev, n_launch = lat.last_step_ms()   # CUDA events recorded around the K launches
    barrier()
    # Sustained figure: the K-step region is only a few ms, a real run is 1e5-1e6 steps.  The same launches run back to
    # back for >= --sustain-s seconds in chunks of 256 steps (device time per chunk from the same CUDA events), the
    # clock sampler covers the whole window, and the line carries both numbers.
    sus_ms, sus_steps, t_probe = 0.0, 0, time.time()
    while time.time() - t_probe < max(a.sustain_s, 0.1):
        lat.step(256)
        lat.sync()
        ms_c, _ = lat.last_step_ms()
        sus_ms += ms_c; sus_steps += 256
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    clocks["window"] = "the K timed steps + %.1f s of the same launches back to back (the sustained window)" % (t1 - t_probe)
    ms_dev = maxr(ms_dev)
    sus_ms = maxr(sus_ms)
    # per-kernel durations: a second pass over K steps with an event between the two launches of each step
    lat.set_kernel_timing(True)
    lat.step(min(K, 64))
    lat.sync()
    ms_dornic, ms_rk4, n_timed = lat.kernel_ms()
    lat.set_kernel_timing(False)
    stats = lat.measure()
    ok = all(s.nonfinite == 0 for s in stats)

    sites_per_step = float(world) * trials * L * L
    value = sites_per_step * K / (ms_dev * 1e-3)
    peak, peak_src = measured_peaks()
    achieved = (trials * L * L * BYTES_PER_SITE_UPDATE) / (ms_dev * 1e-3 / K) / 1e9  # per GPU, GB/s

    # end to end through the host-buffer entry point: H2D of the step's frame + step + D2H of the reductions
    Ke = max(1, min(K, a.e2e_steps))
    lat.run_host(hnp, 1, download=False)
    barrier()
    te0 = time.perf_counter()
    for _ in range(Ke):
        lat.run_host(hnp, 1, download=False)
    lat.sync()
    te = maxr(time.perf_counter() - te0)
    e2e_value = sites_per_step * Ke / te
    # the same entry point at the reference runner's own granularity: the frame goes in once per call
    # (rietkerk_Dornic_2D_MultiSp takes the initial frame and t_max, R/:3397), then `runner_steps` time steps, reductions back
    Kr = max(1, a.runner_steps)
    lat.run_host(hnp, 2, download=False)
    barrier()
    tr0 = time.perf_counter()
    for _ in range(2):
        lat.run_host(hnp, Kr, download=False)
    lat.sync()
    tr = maxr(time.perf_counter() - tr0)
    import ctypes as C
    h2d = trials * 3 * L * L * 8
    d2h = trials * C.sizeof(tm.Stats)
    # what bounds e2e: the host->device link.  Measured here, in the same process and at the same N (all ranks copy at
    # once, as they do in the e2e loop): the bare upload of the same pinned frame through the same call.
    lat.upload(hnp)
    barrier()
    tu0 = time.perf_counter()
    for _ in range(3):
        lat.upload(hnp)
    tu = maxr((time.perf_counter() - tu0) / 3)
    h2d_gbs = h2d / tu / 1e9
    e2e_ceiling = sites_per_step / (tu + (ms_dev * 1e-3 / K))   # serial upload + step, no overlap

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(L, trials),
        "timing": "CUDA events on the kernel stream around K steps, max over ranks", "finite": ok,
        "gpu_launches": n_launch,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic_per_launch(L, trials), "peak_source": peak_src,
                     "traffic_source": "not measured by this run: " + str(ncu_traffic_per_launch(L, trials, True)[1])
                                       + " (the kernels' memory access pattern has not changed since)",
                     "algorithmic_bytes_per_step": trials * L * L * BYTES_PER_SITE_UPDATE,
                     "kernel": "one step = tml::riet_dornic_kernel<STRICT=0> + tml::riet_step_kernel<DET=0,STRICT=0,SRC=pre,DORNIC_ONLY=0>",
                     "kernels": kernel_lines(trials * L * L, ms_dornic, ms_rk4, n_timed, peak)},
        "sustained": {"value": sites_per_step * sus_steps / (sus_ms * 1e-3), "unit": UNIT, "steps": sus_steps,
                      "device_seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / sus_steps,
                      "frac": (trials * L * L * BYTES_PER_SITE_UPDATE) / (sus_ms * 1e-3 / sus_steps) / 1e9 / peak,
                      "vs_burst": (sus_ms / sus_steps) / (ms_dev / K),
                      "note": "same launches back to back, CUDA events per 256-step chunk, max over ranks; clocks.sm_mhz is the median "
                              "over this window"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": Ke, "call": "tm_run_host(ctx, pinned_host_frame, nsteps=1, download=0, stats)",
                "h2d_gbs_measured": h2d_gbs, "link_bound_value": e2e_ceiling, "frac_of_link_bound": e2e_value / e2e_ceiling,
                "note": "bounded by the host->device copy of the 3 L x L planes (measured rate beside it, all ranks copying at "
                        "once); link_bound_value = sites / (upload time + device step time)"},
        "e2e_runner": {"value": sites_per_step * Kr * 2 / tr, "unit": UNIT, "calls": 2, "steps_per_call": Kr,
                       "h2d_bytes_per_call": h2d, "d2h_bytes_per_call": d2h,
                       "call": "tm_run_host(ctx, pinned_host_frame, nsteps=%d, download=0, stats): the frame is copied in once per "
                               "call, as the reference's runner takes one initial frame and t_max; not the headline e2e" % Kr},
    }
    if rank == 0 and world == 1 and not a.no_cpu:
        # same configuration (L x L lattices of the workload's L, as many as fit in host memory), bounded to 2 steps;
        # and the round-1 sample (cache-friendly L = 512, one lattice per core) beside it
        out["cpu_baseline"] = cpu_reference(a.ref_L or L, a.ref_steps, host_cores())
        out["cpu_baseline_small_lattice"] = cpu_reference(512, 12, host_cores())
    lat.close()
    torch.cuda.empty_cache()

    def guarded(key, budget_s, fn, keep=None):
        return run_guarded(out, rank, dist, key, budget_s, fn, keep)

    all_ok = True
    if not a.no_other_configs:
        # BASELINE configs 2, 3 (one GPU's share) and 4 (one GPU's share) beside the headline
        all_ok = guarded("other_configs", a.other_budget_s, lambda: other_configs_measure(rank, world, local, dist))
        torch.cuda.empty_cache()
    if not a.no_other_configs and world == 1:
        # The two "next" rows of SURVEY 8 that have a bench workload of their own, beside the headline, each in a process of its own:
        #   fke_movement_L4096  f2: the FKE movement step, with the reference's own CUDA kernel timed on the same buffers (a kernel
        #                       that is not the product's runs in that process, not in this one)
        #   spb2_step_L4096     f3: the vegetation + grazer (SPB = 2) time step resident on the device
        def sub(workload):
            def run():
                cmd = [sys.executable, os.path.abspath(__file__), "--workload", workload, "--L", "4096", "--steps", "10", "--warmup", "3"]
                pr = subprocess.run(cmd, capture_output=True, text=True, timeout=max(10, a.other_budget_s - 10))
                lines = [l for l in pr.stdout.splitlines() if l.startswith("{")]
                if pr.returncode != 0 or not lines:
                    raise RuntimeError("bench.py --workload %s: rc %d: %s" % (workload, pr.returncode, (pr.stderr or "").strip()[-300:]))
                return json.loads(lines[-1])
            return run
        guarded("fke_movement_L4096", a.other_budget_s, sub("fke"))
        guarded("spb2_step_L4096", a.other_budget_s, sub("spb2"))
    if not a.no_slab and dist is not None and not all_ok:
        if rank == 0:
            out["slab_L16384"] = {"error": "skipped: the other_configs section failed on a rank (the ranks' collectives may be out of step)"}
    elif not a.no_slab:
        # BASELINE config 5 beside the headline, so that every N of a scaling run carries it: ONE L=16384 lattice over the N GPUs
        # of this run (strong scaling; N=1 is the undivided lattice, the base of the speed-up)
        guarded("slab_L16384", a.slab_budget_s, lambda: slab_measure(a, 16384, 10, 3, rank, world, local, dist),
                keep=("value", "unit", "n_gpus", "steps", "ms_per_step", "scaling", "config", "check", "comm", "e2e", "finite"))
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        # the line is out; a rank that failed above must not keep the others in the closing barrier for long
        bye = threading.Timer(60, lambda: os._exit(0))
        bye.daemon = True
        bye.start()
        dist.barrier()
        dist.destroy_process_group()
        bye.cancel()


def other_configs_measure(rank, world, local, dist, shapes=((256, 16), (1024, 32), (2048, 8), (4096, 1)), K=20, W=5):
    """BASELINE configs 2, 3 and 4 beside the headline, so that the driver's own runs carry them (each N of a scaling run included):
    every rank steps its share, no data-path collective, CUDA events on the kernel stream, max over ranks, whole-job value.
      rietkerk_L256x16      config 2: Rietkerk SPB=1, L=256, 16 trials per GPU (state fits the L2: timed warm AND with an L2 flush
                            between single steps)
      rietkerk_fpe_L1024x32 config 3's share of one GPU: Rietkerk_FPE clamp rule (1e-10 / 0, F/:4460), L=1024, 32 trials per GPU
                            (256 trials at N=8)
      dp_L2048x8            config 4's share of one GPU: DP Dornic step, L=2048, 8 a-values in [1.0524, 1.0554] per GPU
      rietkerk_L4096x1_fringe  the headline workload on a frame with an extinction fringe (an eighth of the columns empty, a
                            sixteenth at densities 1e-5..1e-1: the exact Poisson / Gamma samplers' regime, riet_cell_exact in the
                            Dornic sweep's second walk), which the headline's all-vegetated frame never enters
    Not the headline; same kernels, same ABI."""
    import numpy as np
    import torch
    import oracle_lib as O          # the synthetic frame rule only
    import trophic_maelstorm_b200 as tm

    peak, _ = measured_peaks()

    def maxr(x):       # the only collective of this section (so that a failing rank can be told apart, see run_ours.guarded)
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        v = float(t.item())
        if v >= PEER_FAILED:
            raise PeerFailed("another rank failed in this section")
        return v

    def barrier(lat):
        lat.sync()
        maxr(0.0)

    def timed(lat, sites_per_rank, bytes_per_site, flush):
        lat.step(W)
        barrier(lat)
        lat.step(K)
        lat.sync()
        ms, n_launch = lat.last_step_ms()
        ms = maxr(ms)
        r = {"value": world * sites_per_rank * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
             "ms_per_step": ms / K, "scaling": "weak", "dtype": "f64", "data": "synthetic", "gpu_launches": n_launch,
             "roofline_frac": sites_per_rank * bytes_per_site / (ms * 1e-3 / K) / 1e9 / peak}
        if flush:
            # the state fits the 126 MB L2: single steps with a 256 MB write in between (timing rule: flush or exceed the L2)
            junk = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")
            msf = 0.0
            for _ in range(K):
                junk.fill_(1)
                torch.cuda.synchronize()
                lat.step(1)
                lat.sync()
                msf += lat.last_step_ms()[0]
            del junk
            msf = maxr(msf)
            r["l2_flushed"] = {"value": world * sites_per_rank * K / (msf * 1e-3), "ms_per_step": msf / K,
                               "roofline_frac": sites_per_rank * bytes_per_site / (msf * 1e-3 / K) / 1e9 / peak}
        st = lat.measure()
        r["finite"] = all(s.nonfinite == 0 for s in st)
        return r

    out = {}

    def riet(key, L, trials, fpe, what, fringe=False):
        lat = tm.Lattice(tm.rietkerk_params(L, fpe_clamp=fpe), trials, device=local)
        try:
            lat.seed(20261019, trial0=rank * trials)
            base = [O.synthetic_rietkerk_frame(L, seed=777 + k, fringe=fringe) for k in range(min(trials, 4))]
            frames = np.empty((trials, 3, L, L))
            for t in range(trials):
                frames[t] = base[t % len(base)]     # four distinct speckle frames; every trial has its own Philox stream
            lat.upload(frames)
            state_mib = trials * 3 * L * L * 8 / 2 ** 20
            fits = trials * 3 * L * L * 8 * 2 <= 126e6
            r = timed(lat, float(trials) * L * L, BYTES_PER_SITE_UPDATE, fits)
            r["config"] = {"workload": what % (L, trials),
                           "l2": ("state %.0f MiB per GPU (x2 ping-pong) fits the 126 MB L2: `value` is warm, `l2_flushed` has a 256 MB "
                                  "write between single steps" if fits else
                                  "state %.0f MiB per GPU (x2 ping-pong) exceeds the 126 MB L2; no flush needed") % state_mib}
            out[key] = r
        finally:
            lat.close()

    riet("rietkerk_L%dx%d" % shapes[0], shapes[0][0], shapes[0][1], False, "BASELINE config 2: Rietkerk SPB=1 Dornic+RK4 step, L=%d, %d trials per GPU")
    riet("rietkerk_fpe_L%dx%d" % shapes[1], shapes[1][0], shapes[1][1], True,
         "BASELINE config 3 (one GPU's share; 256 trials at N=8): Rietkerk_FPE SPB=1 step (clamp 1e-10 / 0), L=%d, %d trials per GPU")
    riet("rietkerk_L%dx%d_fringe" % shapes[3], shapes[3][0], shapes[3][1], False,
         "the headline workload (Rietkerk SPB=1 Dornic+RK4 step, L=%d, %d trial per GPU) on a frame with an extinction fringe: 12.5 %% of the "
         "columns empty, 6.25 %% at densities 1e-5..1e-1 (exact Poisson / Gamma samplers)", fringe=True)
    L, n_a = shapes[2]
    lat = tm.Lattice(tm.dp_params(L, a=1.0524, b=1.0, D=0.25, sigma=2 ** 0.5, dt=0.1, dx=1.0), n_a, device=local)
    try:
        for i, av in enumerate(np.linspace(1.0524, 1.0554, n_a)):
            lat.set_trial_a(i, float(av))
        lat.seed(7, trial0=rank * n_a)
        lat.upload(np.ones((n_a, 1, L, L)))
        r = timed(lat, float(n_a) * L * L, 16.0, False)
        r["metric"] = "lattice site-updates/sec (FP64, device-timed), DP Dornic step"
        r["config"] = {"workload": "BASELINE config 4 (one GPU's share): DP Dornic step (percolationDornic_2D), L=%d, %d a-values in "
                                   "[1.0524, 1.0554] per GPU, rho0=1; 16 B per site-update, sampler-bound by construction" % (L, n_a),
                       "l2": "state %.0f MiB per GPU (x2 ping-pong) exceeds the 126 MB L2; no flush needed" % (n_a * L * L * 8 / 2 ** 20)}
        out["dp_L%dx%d" % (L, n_a)] = r
    finally:
        lat.close()
    return out


def slab_measure(a, L, K, W, rank, world, local, dist):
    """BASELINE config 5: ONE Rietkerk lattice of side L cut into `world` row slabs (include/tm_slab.h), strong scaling.
    Returns the dict rank 0 prints (None on the other ranks).  world == 1 is the undivided lattice on one GPU: the base of the
    speed-up.  Three schedules are timed over the same K steps: overlapped (the product), serial (exchange, then the whole slab:
    the round-1 schedule) and the overlapped schedule with the transfers left out (what the compute alone costs)."""
    import numpy as np
    import torch
    import oracle_lib as O
    import trophic_maelstorm_b200 as tm
    from trophic_maelstorm_b200 import slab as S

    p = tm.rietkerk_params(L)
    ring = S.SlabRing(p, 1, local, dist if world > 1 else None)
    # the frame: the L=1024 MFT speckle tile repeated periodically (a valid periodic L x L frame; every rank builds only its rows)
    T = 1024 if L % 1024 == 0 else L
    tile = O.synthetic_rietkerk_frame(T, seed=4242, fringe=False)
    rows = (np.arange(ring.row0, ring.row0 + ring.nrows) % T)
    host = torch.empty((1, 3, ring.nrows, L), dtype=torch.float64, pin_memory=True)
    host.numpy()[0] = np.tile(tile[:, rows, :], (1, 1, L // T))
    ring.seed(20261018)
    if world > 1 and a.slab_reserve_sms > 0:
        ring.reserve_sms(a.slab_reserve_sms)
    if world > 1:
        ring.set_transport(S.PEER_COPY if a.slab_transport == "peer" else S.NCCL)

    def barrier():
        ring.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(mode, steps):
        tm.lib().tm_set_step_counter(ring._ctx(), 0)
        ring.upload(host.numpy())
        ring.run(W, mode)
        barrier()
        ring.run(steps, mode)
        ring.sync()
        ms = maxr(ring.last_ms())
        barrier()
        return ms / steps

    ms_overlap = timed(S.OVERLAP, K)
    st = ring.measure()[0]
    active, mean_p = int(st.active[0]), float(st.mean[0]) * ring.nrows / L
    finite = st.nonfinite == 0
    if world > 1:
        t = torch.tensor([active, mean_p, 0.0 if finite else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        active, mean_p, finite = int(t[0].item()), float(t[1].item()), t[2].item() == 0.0
    ms_serial = timed(S.SERIAL, K) if world > 1 else ms_overlap
    ms_nocomm = timed(S.NOCOMM, K) if world > 1 else ms_overlap
    # the schedule a production ring keeps is the faster of the two on this box (a one-off calibration of a few steps): with the
    # copy-engine transport the overlapped one (nothing competes for CTA slots), with NCCL's send/recv kernels the serial one
    best_mode, ms_best = (S.OVERLAP, ms_overlap) if ms_overlap <= ms_serial else (S.SERIAL, ms_serial)
    # the user's call with host buffers: this rank's rows up, K steps, the reductions back
    barrier()
    t0 = time.perf_counter()
    tm.lib().tm_set_step_counter(ring._ctx(), 0)
    ring.upload(host.numpy())
    ring.run(K, best_mode)
    ring.measure()
    te = maxr(time.perf_counter() - t0)
    comm = ring.comm_bytes_per_step()
    ring.close()
    peak, peak_src = measured_peaks()
    sites = float(L) * L
    value = sites / (ms_best * 1e-3)
    return None if rank != 0 else {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_best,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Rietkerk SPB=1 Dornic+RK4 step, ONE lattice L=%d cut into %d row slab(s) of %d rows, 4 halo rows x 3 fields "
                               "per neighbour per step over %s (libtm_slab.so)" % (L, world, L // world, "CUDA IPC peer copies on the copy engines + stream-ordered flags"
                                                                                 if a.slab_transport == "peer" else "ncclSend/ncclRecv"),
                   "parallelism": "row slabs, periodic ring, halo traffic behind the inner rows' update" if world > 1 else "one GPU, undivided lattice",
                   "l2": "slab state %.0f MiB per GPU (x2 ping-pong) exceeds the 126 MB L2" % (3 * ring.nrows * L * 8 / 2 ** 20)},
        "timing": "CUDA events on each rank's main stream around K steps (the join with the band/comm stream included), max over ranks",
        "finite": bool(finite), "gpu_launches": (6 if world > 1 else 2) * K,
        "check": {"active_sites": active, "mean_P": mean_p, "note": "the same for every number of slabs (bit-identical fields; the mean is summed per rank)"},
        "roofline": {"bound": "hbm", "achieved": sites * BYTES_PER_SITE_UPDATE / world / (ms_best * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": sites * BYTES_PER_SITE_UPDATE / world / (ms_best * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "note": "per GPU: its slab's 48 B per site-update over the step time"},
        "comm": {"transport": a.slab_transport, "schedule": "overlapped" if best_mode == S.OVERLAP else "serial",
                 "bytes_sent_per_step_per_rank": comm, "ms_per_step_overlapped": ms_overlap, "ms_per_step_serial_schedule": ms_serial,
                 "ms_per_step_without_transfers": ms_nocomm, "exposed_ms_per_step": ms_best - ms_nocomm,
                 "note": "value is the faster schedule; exposed = that schedule minus the band/inner schedule with the transfers left out"},
        "e2e": {"value": sites * K / te, "unit": UNIT, "h2d_bytes_per_step": 3 * ring.nrows * L * 8 / K, "d2h_bytes_per_step": 64.0 / K,
                "steps": K, "call": "tm_upload(rows of this rank) + tm_slab_run(K) + tm_measure: the frame goes in once per call "
                                    "(a slab ring keeps its state resident between exchanges), wall clock, max over ranks"},
    }


def run_slab(a):
    import torch
    rank, local, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = a.L if a.L != 4096 else 16384
    sampler = ClockSampler(local)
    sampler.start()
    t0 = time.time()
    out = slab_measure(a, L, a.steps, max(a.warmup, 3), rank, world, local, dist)
    clocks = sampler.stop(t0, time.time())
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        out["clocks"] = clocks
        print(json.dumps(out), flush=True)


def run_dp_extra(a):
    """Not the headline: DP a-sweep (BASELINE config 4 shape) through the same ABI, for the record."""
    import numpy as np
    import trophic_maelstorm_b200 as tm
    L, n_a, K = a.L if a.L != 4096 else 2048, 8, a.steps
    p = tm.dp_params(L, a=1.0524, b=1.0, D=0.25, sigma=2 ** 0.5, dt=0.1, dx=1.0)
    lat = tm.Lattice(p, n_a, device=int(os.environ.get("LOCAL_RANK", "0")))
    for i, av in enumerate(np.linspace(1.0524, 1.0554, n_a)):
        lat.set_trial_a(i, float(av))
    lat.seed(7)
    lat.upload(np.ones((n_a, 1, L, L)))
    lat.step(max(a.warmup, 3)); lat.sync()
    lat.step(K); lat.sync()
    ms, n = lat.last_step_ms()
    peak, src = measured_peaks()
    rate = n_a * L * L * K / (ms * 1e-3)
    print(json.dumps({"metric": "lattice site-updates/sec (FP64, device-timed), DP Dornic step", "value": rate, "unit": UNIT,
                      "n_gpus": 1, "steps": K, "ms_per_step": ms / K, "dtype": "f64", "data": "synthetic", "gpu_launches": n,
                      "config": {"workload": "DP Dornic step (percolationDornic_2D), L=%d, %d a-values in [1.0524,1.0554], rho0=1" % (L, n_a)},
                      "roofline": {"bound": "hbm", "achieved": rate * 16 / 1e9, "peak": peak, "unit": "GB/s", "frac": rate * 16 / 1e9 / peak,
                                   "traffic": None, "peak_source": src, "note": "16 B per site-update; RNG/ALU-bound by construction"}}), flush=True)
    lat.close()


def run_fke_extra(a):
    """Not the headline: the FKE movement step of one consumer species, L x L, through include/tm_fke.h.
    A "site-update" here = one lattice site moved once.  Algorithmic bytes: rho, gamma (8 B each) and v_eff (16 B) read,
    rho written (8 B) = 40 B per site; the reference's own CPU advdiff_FKE_MultiSp is timed beside it."""
    import ctypes as C
    import numpy as np
    import torch
    import trophic_maelstorm_b200 as tm
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    L, K, W = a.L, a.steps, max(a.warmup, 3)
    bD, bV, pref = a.fke_bound_d, a.fke_bound_v, 0.82
    rng = np.random.default_rng(5)
    rho = rng.gamma(2.0, 3.0, L * L); rho[rng.random(L * L) < 0.3] = 0.0
    gamma = rng.random(L * L); veff = (rng.random((L * L, 2)) * 2 - 1) * 1.2
    def stencil(sigma, R):      # a normalised half-Gaussian table [0..R] (set-up of the product arm: numpy only, nothing from oracle/)
        k = np.arange(R + 1, dtype=np.float64)
        g = np.exp(-0.5 * (k / sigma) ** 2)
        return np.ascontiguousarray(g / (g[0] + 2.0 * g[1:].sum()))
    gD, gV = stencil(1.0, bD), stencil(1.0, bV)
    m = tm.FkeMover(L, 1, device=0)
    m.set_terms(0, bD, (bV, bV), (bV, bV), pref, 1e-10, gD, gV)
    dev = torch.device("cuda:0")
    d = [torch.from_numpy(x).to(dev) for x in (rho, gamma, veff)]
    out = torch.empty(L * L, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    for _ in range(W):
        m.move_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), out.data_ptr())
    m.sync()
    tot = 0.0
    for _ in range(K):   # inputs (537 MB) + scratch exceed the L2 at L = 4096
        m.move_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), out.data_ptr())
        ms, n = m.last_ms()
        tot += ms
    ms = tot / K
    t0 = time.perf_counter()
    for _ in range(3):
        res = m.move(rho, gamma, veff)
    te = (time.perf_counter() - t0) / 3
    peak, src = measured_peaks()
    rate = L * L / (ms * 1e-3)
    line = {"metric": "lattice sites moved/sec (FP64, device-timed), FKE movement step", "value": rate, "unit": "sites/s", "n_gpus": 1,
            "steps": K, "warmup": W, "ms_per_step": ms, "dtype": "f64", "data": "synthetic", "gpu_launches": 3 * K,
            "config": {"workload": "FKE movement (advdiff_FKE_MultiSp / updateMovement), 1 species, L=%d, bound_D=%d, bound_vD=%d, "
                                   "displacements within +-1" % (L, bD, bV), "max_displacement": list(m.last_displacement(0))},
            "roofline": {"bound": "hbm", "achieved": rate * 40 / 1e9, "peak": peak, "unit": "GB/s", "frac": rate * 40 / 1e9 / peak,
                         "traffic": None, "peak_source": src, "note": "40 B per site algorithmic; three passes (prepare, columns, rows)"},
            "e2e": {"value": L * L / te, "unit": "sites/s", "h2d_bytes_per_step": 32 * L * L, "d2h_bytes_per_step": 8 * L * L,
                    "call": "tm_fke_move_host (pageable numpy buffers)"}}
    if not a.no_cpu:
        Lc = min(L, 1024)
        sl = lambda x, tail=(): np.ascontiguousarray(x.reshape((L, L) + tail)[:Lc, :Lc]).reshape((Lc * Lc,) + tail)
        r_, g_, v_ = sl(rho), sl(gamma), sl(veff, (2,))
        refso = os.path.join(ROOT, "oracle", "_ref", "libtm_ref_fke.so")
        t0 = time.perf_counter()
        if os.path.exists(refso):
            ref = C.CDLL(refso); kind = "reference"
            ref.ref_fke_override(C.c_int(bD), C.c_int(bV), C.c_int(bV), C.c_double(pref), C.c_int(Lc))
            o = np.zeros(Lc * Lc); dp = C.POINTER(C.c_double); P = lambda x: x.ctypes.data_as(dp)
            t0 = time.perf_counter()
            ref.ref_fke_advdiff(C.c_int(Lc), P(r_), P(g_), P(v_), P(gD), C.c_int(gD.size), P(gV), C.c_int(gV.size), P(o))
        else:
            kind = "port"
            o = O.fke_move(Lc, bD, (bV, bV), (bV, bV), pref, 1e-10, r_, g_, v_, gD, gV)
        tc = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": Lc * Lc / tc, "unit": "sites/s", "cores": 1, "kind": kind,
                                "sample": "advdiff_FKE_MultiSp on the L=%d corner of the same inputs, one call (the reference's loop is serial)" % Lc}
    # the reference's OWN CUDA kernel on the same device buffers (updateMovement, kernel_advdiff_FKE2D.cu:92-180, compiled unmodified
    # for sm_100a into oracle/_ref/libtm_ref_fke_gpu.so): a reported baseline beside the product's three kernels
    refgpu = os.path.join(ROOT, "oracle", "_ref", "libtm_ref_fke_gpu.so")
    if os.path.exists(refgpu):
        try:
            rg = C.CDLL(refgpu)
            rg.ref_fke_gpu_setup.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
            rg.ref_fke_gpu_move.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
            m.move_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), out.data_ptr())
            m.sync()
            ours = out.clone()
            sD, sV = torch.from_numpy(gD).to(dev), torch.from_numpy(gV).to(dev)
            tsar = torch.empty(L * L, dtype=torch.float64, device=dev)
            torch.cuda.synchronize()
            # the reference's set-up prints its constants with printf: keep this program's stdout to the one JSON line
            libc = C.CDLL(None)
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                rc = rg.ref_fke_gpu_setup(bD, bV, bV, pref)
                libc.fflush(None)
            finally:
                os.dup2(saved, 1)
                os.close(saved)
            ms_r, ms_k = C.c_float(0), C.c_float(0)
            if rc == 0:
                rc = rg.ref_fke_gpu_move(L, d[0].data_ptr(), tsar.data_ptr(), d[1].data_ptr(), d[2].data_ptr(), sD.data_ptr(), sV.data_ptr(),
                                         W, C.byref(ms_r), C.byref(ms_k))
            if rc == 0:
                rc = rg.ref_fke_gpu_move(L, d[0].data_ptr(), tsar.data_ptr(), d[1].data_ptr(), d[2].data_ptr(), sD.data_ptr(), sV.data_ptr(),
                                         K, C.byref(ms_r), C.byref(ms_k))
            if rc != 0:
                raise RuntimeError("cudaError_t %d from the reference kernel's harness" % rc)
            diff = float((tsar - ours).abs().max().item()) / max(float(ours.abs().max().item()), 1e-300)
            line["reference_gpu_baseline"] = {
                "kernel": "updateMovement (Rietkerk_FPE/kernel_advdiff_FKE2D.cu:92-180), unmodified, sm_100a, FP64 atomics",
                "ms_per_step": ms_r.value, "ms_kernel_only": ms_k.value, "value": L * L / (ms_r.value * 1e-3), "unit": "sites/s",
                "steps": K, "reference_ms_over_product_ms": ms_r.value / ms,
                "max_abs_diff_over_max": diff,
                "note": "same device buffers; per step the reference zeroes Rho_tsar and launches once (F/:4282-4299 without its host "
                        "copies); the product's step is its three launches"}
        except Exception as e:
            line["reference_gpu_baseline"] = {"error": "%s: %s" % (type(e).__name__, e)}
    print(json.dumps(line), flush=True)
    m.close()


def run_spb2_extra(a):
    """Not the headline: the vegetation + grazer (SPB = 2) time step resident on the device (include/tm_consumer.h
    tm_cons_step2_*): six launches per step.  Algorithmic bytes: four fields read and written once = 64 B per site-update."""
    import numpy as np
    import trophic_maelstorm_b200 as tm
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    L, K, W = a.L, a.steps, max(a.warmup, 3)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "runner_2sp.npz"))
    cons = np.load(os.path.join(ROOT, "tests", "golden", "consumer_ops.npz"))
    par = gold["mixed_par"].astype(np.float64)
    pa, c, gmax, alpha, rW, W0, d, K1, K2, D0, D1, D2, D3, s0, s1, v1, aij, hij, ej, m1, pR1, dt, dx = [float(x) for x in par]
    R = int(pR1 / dx)
    central = np.ascontiguousarray(cons["central_R%d" % R], dtype=np.int32)
    fr = pR1 / (dx * (pR1 / dx))
    veg = tm.rietkerk_params(L, a=pa, dt=dt)
    veg.dx, veg.rc, veg.gmax, veg.alpha, veg.rW, veg.W0, veg.d, veg.K1, veg.K2 = dx, c, gmax, alpha, rW, W0, d, K1, K2
    veg.D0, veg.D1, veg.D2, veg.sigma0 = D0, D2, D3, s0
    graz = tm.GrazParams(L, 0, dt, dx, D1, v1, m1, s1, D0 / (dx * dx))
    cp = tm.ConsParams(L, 0, dt, pa, c, gmax, alpha, rW, W0, K1, K2, D2 / (dx * dx), D3 / (dx * dx), aij, hij, ej)
    rng = np.random.default_rng(7)
    init = gold["mixed_init"]
    f = np.stack([rng.choice(init[s_], L * L) for s_ in range(4)])
    ops = tm.ConsumerOps(L)
    ops.step2_setup(veg, graz, cp, central, fr)
    ops.step2_upload(f, None)
    ops.step2_run(1, nsteps=W)
    ops.step2_download()
    ops.step2_run(1, step0=W, nsteps=K)
    ms, n = ops.last_ms()
    fields = ops.step2_download()[0]
    peak, src = measured_peaks()
    rate = L * L * K / (ms * 1e-3)
    line = {"metric": "lattice site-updates/sec (FP64, device-timed), vegetation + grazer (SPB=2) Dornic+RK4 step", "value": rate,
            "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K, "dtype": "f64", "data": "synthetic", "gpu_launches": n,
            "config": {"workload": "SPB=2 runner step (rietkerk_Dornic_2D_MultiSp -DSPB=2 loop body), L=%d, perception disc of %d sites"
                                   % (L, len(central)), "finite": bool(np.isfinite(fields).all())},
            "roofline": {"bound": "hbm", "achieved": rate * 64 / 1e9, "peak": peak, "unit": "GB/s", "frac": rate * 64 / 1e9 / peak,
                         "traffic": None, "peak_source": src, "note": "64 B per site-update algorithmic; six launches per step"}}
    if not a.no_cpu:
        Lc = 256
        pv, pg, pc, _, _ = O.runner2_params(Lc, par)
        fc = np.stack([rng.choice(init[s_], Lc * Lc) for s_ in range(4)])
        g = np.zeros(Lc * Lc)
        t0 = time.perf_counter()
        for s_ in range(3):
            O.step_2sp(pv, pg, pc, central, fr, O.make_rng(O.SRC_PHILOX, seed=1, step=s_), fc, g)
        tc = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 3 * Lc * Lc / tc, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "oracle tmo_step_2sp, L=%d, 3 steps, one thread" % Lc}
    print(json.dumps(line), flush=True)
    ops.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--L", type=int, default=4096)
    ap.add_argument("--trials", type=int, default=1, help="trials per GPU")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--runner-steps", type=int, default=100)
    ap.add_argument("--ref-L", type=int, default=0, help="lattice side of the CPU sample (0: the workload's own L)")
    ap.add_argument("--ref-steps", type=int, default=2, help="time steps of the cpu_baseline leg of the product arm")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="seconds of back-to-back steps for the sustained figure")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-slab", action="store_true", help="skip the L=16384 row-slab line (BASELINE config 5) the default run carries under 'slab_L16384'")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the BASELINE config 2-4 lines the default run carries under 'other_configs'")
    ap.add_argument("--other-budget-s", type=int, default=180, help="wall-clock budget of the other-configs section of the default run")
    ap.add_argument("--workload", default="rietkerk", choices=["rietkerk", "dp", "fke", "spb2", "slab"],
                    help="dp, fke, spb2: extra, non-headline lines; slab: BASELINE config 5, one L=16384 lattice over the ranks (strong scaling)")
    ap.add_argument("--slab-transport", default="nccl", choices=["peer", "nccl"],
                    help="slab: halo transport (nccl: ncclSend/ncclRecv, measured on 2, 4 and 8 GPUs; peer: CUDA IPC copies on the copy engines, "
                         "measured on 2 and 4 GPUs only)")
    ap.add_argument("--slab-budget-s", type=int, default=150, help="wall-clock budget of the slab section of the default run")
    ap.add_argument("--slab-reserve-sms", type=int, default=0, help="slab: SMs' worth of CTA slots the inner RK4 launch leaves to the band and transport kernels")
    ap.add_argument("--fke-bound-d", type=int, default=1, help="fke: diffusion window half-width (the SPB=2 driver's constants give 1)")
    ap.add_argument("--fke-bound-v", type=int, default=1, help="fke: advection window half-width")
    a = ap.parse_args()
    if a.workload == "dp" and a.impl == "ours":
        return run_dp_extra(a)
    if a.workload == "fke" and a.impl == "ours":
        return run_fke_extra(a)
    if a.workload == "spb2" and a.impl == "ours":
        return run_spb2_extra(a)
    if a.workload == "slab" and a.impl == "ours":
        return run_slab(a)
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
