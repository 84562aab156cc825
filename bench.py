#!/usr/bin/env python
"""bench.py — objective + gradient evaluations/s of the B200 engine (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's engine
    python bench.py --impl reference --gpus N --steps K ...   # the reference arm

One *step* = one objective + gradient evaluation through the C++ host layer
(ObjectiveFunction::calcFunction + calcGradient with the Fi terms BASELINE.json names for the
config: Chi2 + L1 + TSV for configs[1]) over the whole visibility set; for N > 1 the engine
all-reduces the chi2 scalar and the [2][M][N] gradient over NCCL inside those calls. The workload is
BASELINE.json configs[1] (ALMA-like 2048^2 image x 10 M visibilities, 1 channel) on
synthetic data; at N > 1 the visibilities are sharded by contiguous chunk (strong
scaling). `value` is whole-job Mvis*Mpix/s = (Z/1e6)*(M*N/1e6)*evals/s, which is
size-independent, so the reference arm's bounded sample is directly comparable.

Timing: W >= 3 warm-up steps, then exactly K steps bracketed by barrier + synchronize,
CUDA events on the launching stream, max over ranks; inputs (520 MB of visibilities) are
larger than L2. Prints ONE JSON line on rank 0.
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
sys.path.insert(0, ROOT)

METRIC = "obj+grad evals/sec (Mvis*Mpix/s)"
UNIT = "Mvis*Mpix/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"],
                    tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class NativeStdoutToStderr:
    """Native code (the reference prints its progress with printf) must not write into the
    stream that carries the ONE JSON line: fd 1 points at stderr inside the block."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def make_workload(args):
    from gpuvmem_b200 import synth
    if args.config == "c2":
        return synth.config_c2(scale=args.scale), "BASELINE.json configs[1]: ALMA-like synthetic 2048x2048, 10M visibilities, 1 channel"
    if args.config == "c1":
        return synth.config_c1(scale=args.scale), "BASELINE.json configs[0]: co65-shaped 512x512, 2^20 visibilities, 1 channel"
    if args.config == "c3":
        return synth.config_c3(scale=args.scale), "BASELINE.json configs[2]: MFS 64 channels x 1M visibilities, 2048x2048"
    if args.config == "c4":
        return synth.config_c4(scale=args.scale), "BASELINE.json configs[3]: VLBI-like 4096x4096, 50M visibilities"
    if args.config == "c4l":
        return synth.config_c4(scale=args.scale), ("BASELINE.json configs[3] read literally: VLBI-like 4096x4096, 50M UNGRIDDED visibilities, "
                                                   "PSWF_12D 9x9 convolutional degridding in the forward model, exact DFT gradient")
    if args.config == "c5":
        return synth.config_c5(scale=args.scale), "BASELINE.json configs[4]: gridded mode, Briggs R=0, 8192x8192 grid, 200M visibilities"
    raise SystemExit(f"unknown --config {args.config}")


# per workload: the reference command line (initial values, -Z factors with main.cu's index map:
# Entropy 0, L1-Norm 1, TSV 2, Laplacian 3), the Fi terms BASELINE.json names, the optimizer
WORKLOAD_SETUP = {
    "c1": ("-z 0.001 -Z 0.01", "Chi2:-1:0:0,Entropy:0:0:0", "CG-FRPRMN", "Natural", "PillBox2D", (0, 0)),
    "c2": ("-z 0.001 -Z 0.0,0.005,0.002", "Chi2:-1:0:0,L1-Norm:1:0:0,TotalSquaredVariation:2:0:0", "CG-LBFGS",
           "Natural", "PillBox2D", (0, 0)),
    "c3": ("-z 0.001,0.0 -Z 0.01", "Chi2:-1:0:0,Entropy:0:0:0", "CG-FRPRMN", "Natural", "PillBox2D", (0, 0)),
    # PSWF_12D is a GRIDDING kernel in the reference (it only acts under -g; degriddingGPU is dead code)
    "c4": ("-z 0.001 -Z 0.01 -g 1", "Chi2:-1:0:0,Entropy:0:0:0", "CG-FRPRMN", "Natural", "PSWF", (9, 9)),
    # C4 read literally: nothing is gridded; the PSWF table is the DEGRIDDING kernel of the forward model
    # (gvmh_use_ckernel_degridding: degriddingGPU, src/functions.cu:2205-2254, + its GCF), gradient = tcgen05 DFT
    "c4l": ("-z 0.001 -Z 0.01", "Chi2:-1:0:0,Entropy:0:0:0", "CG-FRPRMN", "Natural", "PSWF", (9, 9)),
    # gridded mode: Briggs R = 0 weights, convolutional gridding (-g), then the objective on the gridded samples
    "c5": ("-z 0.001 -Z 0.01 -g 1 -R 0.0", "Chi2:-1:0:0,Entropy:0:0:0", "CG-FRPRMN", "Briggs", "Gaussian2D", (7, 7)),
}


def cpu_baseline(problem, engine_meta, seconds_target=15.0):
    """The oracle port (plain C + OpenMP, all host cores) on a bounded sample of the SAME
    workload: forward chi2 over Zs visibilities + the DFT gradient at npix sampled pixels."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _checkers import Oracle
    o = Oracle()
    cores = o.threads()
    N = problem.N
    Zs = min(20000, len(problem.w[0]))
    sub = problem.subset(Zs)
    cfg = dict(D=problem.antenna_diameter, DELTAX=problem.DELTAX, DELTAY=problem.DELTAY, eta=-1.0)
    meta = dict(engine_meta)
    prep = o.prep(sub.uvw[0], sub.Vo[0], sub.w[0], float(sub.freqs[0]), meta["deltau"], meta["deltav"], N)
    noise = np.zeros((N, N), np.float32)  # nothing masked: every sampled pixel does the full sum
    meta["noise_cut"] = 1.0
    # calibrate npix for ~seconds_target
    rate_guess = 4.0e7 * cores   # pairs/s
    npix = int(max(256, min(N * N, seconds_target * rate_guess / Zs)))
    pix = np.linspace(0, N * N - 1, npix).astype(np.int64)
    Vr = np.ascontiguousarray(prep["Vo"])
    t0 = time.perf_counter()
    o.dchi2(pix, N, prep["uvw"], Vr, prep["w"], noise, None, float(sub.freqs[0]), meta, cfg)
    dt = time.perf_counter() - t0
    value = (Zs / 1e6) * (npix / 1e6) / dt
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle/gvm_oracle.c gvo_dchi2 (fp64, OpenMP): {npix} of {N*N} pixels x {Zs} visibilities in {dt:.1f} s"}


def cpu_gridding_baseline(problem, ckernel, ck_size, scheme, nvis=200000):
    """The reference's OpenMP CPU gridding path (-g: WeightingScheme::apply + do_gridding, host code of
    the unmodified reference inside oracle/_ref/libgvref.so) on a bounded sample of the workload's first
    channel and the workload's own grid, with 1 thread and with every host core; and this engine's
    gvm_weights + gvm_grid_block on the same sample. A reported baseline, not the target."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _checkers import GVREF_SO, GvRef
    if not os.path.exists(GVREF_SO):
        return None
    from gpuvmem_b200 import host
    from gpuvmem_b200.engine import RPDEG_D, grid_block, weights
    sub = problem.subset(nvis)
    for c in range(sub.nchan - 1, 0, -1):       # first channel only
        del sub.uvw[c], sub.Vo[c], sub.w[c]
    sub.freqs = sub.freqs[:1].copy()
    Zs = len(sub.w[0])
    ck, (m, n) = (ckernel, ck_size) if ck_size[0] > 1 else ("Gaussian2D", (7, 7))
    sch = scheme if scheme != "Natural" else ""
    ref = GvRef()
    ref.set_problem(sub)
    cores = os.cpu_count() or 1
    out = {"unit": "Mvis/s", "sample": f"{Zs} visibilities of channel 0 onto the {problem.M}x{problem.N} grid, {ck} {m}x{n}"
                                       f"{', ' + sch + ' R=0 weights' if sch else ''}", "kind": "reference", "cores": cores}
    with NativeStdoutToStderr():
        for label, th in (("threads_1", 1), ("threads_all", cores)):
            ref.cpu_gridding(ck, m, n, scheme=sch, robust=0.0, threads=th)
            tw, tg = ref.cpu_last_seconds()
            out[label] = {"Mvis_per_s": Zs / 1e6 / (tw + tg), "weighting_s": tw, "gridding_s": tg, "threads": th}
    du, dv = 1.0 / (sub.M * RPDEG_D * sub.DELTAX), 1.0 / (sub.N * RPDEG_D * sub.DELTAY)
    table, support = host.ckernel_table(ck, m, n, np.float32(abs(du)), np.float32(abs(dv)))
    for rep in range(2):                         # second pass: buffers allocated, context warm
        t0 = time.perf_counter()
        w = sub.w[0].copy()
        if sch:
            w = weights(sch, sub.M, sub.N, du, dv, [sub.uvw[0]], [float(sub.freqs[0])], [w], robust=0.0)[0]
        grid_block(sub.M, sub.N, du, dv, float(sub.freqs[0]), sub.uvw[0], sub.Vo[0], w, table, support)
        dt = time.perf_counter() - t0
    out["value"] = out["threads_all"]["Mvis_per_s"]
    out["this_engine_Mvis_per_s_same_sample"] = Zs / 1e6 / dt
    return out


def workload_config(name, problem, wl):
    """The `config` object of a bench line: what was measured, independent of which arm measured it (both
    arms print exactly this for the same --config, so the driver can tell they ran the same workload)."""
    cli, fi_spec, optimizer, scheme, ckernel, ck_size = WORKLOAD_SETUP[name]
    return {"workload": wl, "image": f"{problem.M}x{problem.N}", "visibilities": problem.total_vis(),
            "channels": problem.nchan, "terms": fi_spec, "cli": cli, "optimizer": optimizer, "weighting": scheme,
            "ckernel": f"{ckernel} {ck_size[0]}x{ck_size[1]}" if ("-g" in cli or name == "c4l") else None}


REF_RECON_HANDOFF = "/tmp/gvm_b200_reference_recon_{cfg}.npz"   # reference arm -> this arm, same box


def reference_recon(args):
    """Full reconstruction by the reference's own CUDA build (MFS::run through gvref_run: optimizer from the flat
    start image) on the WHOLE workload of --config; meant for C1 (CG, 50 iterations ~ tens of seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _checkers import GvRef
    problem, wl = make_workload(args)
    cli, fi_spec, optimizer, scheme, ckernel, ck_size = WORKLOAD_SETUP[args.config]
    with NativeStdoutToStderr():
        ref = GvRef()
        ref.set_problem(problem)
        ref.lib.gvref_set_verbose(0)
        ref.init(f"-X 16 -Y 16 -V 256 {cli} -t {args.recon_iters} -i synth.ms -o out.ms -m hdr.fits",
                 optimizer=optimizer, scheme=scheme, ckernel=ckernel, ck_m=max(ck_size[0], 1), ck_n=max(ck_size[1], 1))
        if optimizer == "CG-LBFGS":
            ref.lib.gvref_set_lbfgs_k(args.lbfgs_k)
        t0 = time.perf_counter()
        img, iters, ms = ref.run()
        wall = time.perf_counter() - t0
    out = {"config": args.config, "optimizer": optimizer, "iterations": int(iters), "seconds": ms / 1e3,
           "wall_seconds": wall, "terms": "main.cu's five Fi with -Z " + cli.split("-Z")[1].split()[0],
           "visibilities": problem.total_vis(), "image": f"{problem.M}x{problem.N}"}
    try:
        np.savez(REF_RECON_HANDOFF.format(cfg=args.config), image=img, **{k: v for k, v in out.items() if not isinstance(v, str)})
    except Exception:
        pass
    return out


def run_reference(args):
    """Reference arm: gpuvmem's own CUDA implementation of the path (its only implementation — the reference
    has no CPU objective/gradient), compiled unmodified for sm_100a into oracle/_ref/libgvref.so. The K timed
    steps run on a bounded visibility sample of the workload (DChi2 is a serial loop over the visibilities per
    pixel, so the metric is size-independent); `full_workload_check` backs that with ONE evaluation of the
    whole workload, and a three-point sweep over the sample size is committed under profiles/. Falls back to
    the CPU oracle port when that library or a GPU is missing."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.recon_only:
        print(json.dumps({"impl": "reference", "recon": reference_recon(args)}), flush=True)
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _checkers import GVREF_SO, GvRef
    problem, wl = make_workload(args)
    N = problem.N
    Ztot = problem.total_vis()
    Zs = min(args.ref_sample, len(problem.w[0]))
    sub = problem.subset(Zs)
    Zsub = sub.total_vis()
    warm = max(1, min(args.warmup, 2))    # a reference step takes seconds; the arm has to end within minutes
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.config, problem, wl),
            "sample": {"visibilities_per_step": Zsub, "visibilities_in_workload": Ztot, "warmup_steps_run": warm}}
    have_gpu = False
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    if have_gpu and os.path.exists(GVREF_SO):
        cli, fi_spec, optimizer, scheme, ckernel, ck_size = WORKLOAD_SETUP[args.config]
        unit = lambda Z, ms: (Z / 1e6) * (N * N / 1e6) / (ms / 1e3)

        def init(p):
            ref = GvRef()
            ref.set_problem(p)
            ref.lib.gvref_set_verbose(0)
            ref.init(f"-X 16 -Y 16 -V 256 {cli} -t {max(args.recon_iters, 1)} -i synth.ms -o out.ms -m hdr.fits",
                     optimizer=optimizer, scheme=scheme, ckernel=ckernel, ck_m=max(ck_size[0], 1), ck_n=max(ck_size[1], 1))
            return ref

        with NativeStdoutToStderr():
            ref = init(sub)
            sampler = ClockSampler(0)
            for _ in range(warm):
                ref.time_evals(1, iteration=1)
            sampler.start()
            ms = ref.time_evals(args.steps, iteration=1)
            clocks = sampler.stop()
        value = unit(Zsub, ms)
        line.update({"value": value, "ms_per_step": ms, "clocks": clocks, "gpu_launches": None,
                     "evals_per_s_extrapolated_to_workload": value / ((Ztot / 1e6) * (N * N / 1e6)),
                     "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "cpu_baseline": {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                                      "sample": f"reference CUDA build (unmodified src/*.cu, sm_100a) calcFunction()+calcGradient() on "
                                                f"{Zsub} of {Ztot} visibilities, full {N}x{N} image, 1 GPU; "
                                                "gpuvmem has no CPU implementation of this path"}})
        # ONE evaluation of the whole workload in a fresh process (the reference keeps its problem in process
        # globals): the sample-based figure above must agree with it
        if args.full_check and args.gpus == 1 and Zsub < Ztot:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", args.config,
                                    "--scale", str(args.scale), "--full-eval-only"], capture_output=True, text=True, timeout=1500)
                full = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
                full["ratio_to_sampled_value"] = full["value"] / value
                line["full_workload_check"] = full
            except Exception as exc:
                line["full_workload_check"] = {"error": repr(exc)}
        # the second half of the BASELINE metric for the config the reference can finish: C1, CG 50 iterations
        if args.ref_recon and args.gpus == 1:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", "c1",
                                    "--recon-only", "--recon-iters", "50"], capture_output=True, text=True, timeout=900)
                line["recon"] = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])["recon"]
            except Exception as exc:
                line["recon"] = {"error": repr(exc)}
    else:
        from gpuvmem_b200.engine import beam_model
        pbf, pbc, pb = beam_model(problem.telescope, problem.antenna_diameter, float(problem.freqs.min()))
        meta = dict(deltau=1.0 / (N * np.deg2rad(problem.DELTAX)), deltav=1.0 / (N * np.deg2rad(problem.DELTAY)),
                    fg_scale=1.0, pb_factor=pbf, pb_cutoff=pbc, primary_beam=pb, xpix=N / 2, ypix=N / 2,
                    nu_0=float(problem.freqs[0]), noise_cut=1.0)
        t0 = time.perf_counter()
        cb = cpu_baseline(problem, meta, seconds_target=10.0)
        line.update({"value": cb["value"], "ms_per_step": (time.perf_counter() - t0) * 1e3, "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


def reference_full_eval(args):
    """ONE calcFunction + calcGradient of the reference CUDA build over the whole workload (C2: ~65 s)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _checkers import GvRef
    problem, wl = make_workload(args)
    N, Ztot = problem.N, problem.total_vis()
    cli, fi_spec, optimizer, scheme, ckernel, ck_size = WORKLOAD_SETUP[args.config]
    with NativeStdoutToStderr():
        ref = GvRef()
        ref.set_problem(problem)
        ref.lib.gvref_set_verbose(0)
        ref.init(f"-X 16 -Y 16 -V 256 {cli} -t 1 -i synth.ms -o out.ms -m hdr.fits", optimizer=optimizer, scheme=scheme,
                 ckernel=ckernel, ck_m=max(ck_size[0], 1), ck_n=max(ck_size[1], 1))
        ms = ref.time_evals(1, iteration=1)
        v, fi = ref.calc_function(iteration=1)
    print(json.dumps({"visibilities": Ztot, "evaluations": 1, "ms_per_step": ms, "unit": UNIT,
                      "value": (Ztot / 1e6) * (N * N / 1e6) / (ms / 1e3), "objective": float(v), "half_chi2": float(fi[0])}), flush=True)


class Ctx:
    """Rank / world / torch handles shared by the measurements of one bench.py process."""

    def __init__(self, gpus):
        import torch
        import torch.distributed as dist
        from gpuvmem_b200 import dist as gdist
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local = gdist.init_from_env(gpus)
        torch.cuda.set_device(self.local)

    def nccl_id(self):
        from gpuvmem_b200 import host
        if self.world <= 1:
            return None
        box = [host.nccl_unique_id() if self.rank == 0 else None]   # rank 0 creates it; the launcher's rendezvous carries it
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def max_over_ranks(self, x):
        if self.world <= 1:
            return x
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def measure(ctx, name, scale, steps, warmup, recon_iters, grad_mode=0, lbfgs_k=10, sample_clocks=False, cpu=False,
            recon_reference=False):
    """One workload through the public API (C++ host layer over the C ABI): K timed objective + gradient
    evaluations with the image resident in HBM (`value`), the same end to end from pinned host memory (`e2e`),
    the full reconstruction, the dominant kernel's roofline entry, and a `check` block (chi2, gradient norm and
    sum at the start image) that must agree between runs at different GPU counts."""
    from gpuvmem_b200 import host
    torch, dist, rank, world, local = ctx.torch, ctx.dist, ctx.rank, ctx.world, ctx.local
    ns = argparse.Namespace(config=name, scale=scale)
    t_gen = time.perf_counter()
    problem, wl = make_workload(ns)
    t_gen = time.perf_counter() - t_gen
    M, N = problem.M, problem.N
    Ztot, MN = problem.total_vis(), problem.M * problem.N
    cli, fi_spec, optimizer, scheme, ckernel, ck_size = WORKLOAD_SETUP[name]
    host.set_quiet(True)   # stdout carries the JSON line only
    # Every rank builds the same Session; MFS::setDevice uploads this rank's shard (host.shard_plan).
    s = host.Session(problem, args=f"{cli} -t {max(recon_iters, 1)} -G {local} -K {grad_mode}", optimizer=optimizer,
                     scheme=scheme, ckernel=ckernel, ck_size=ck_size, fi_spec=fi_spec, rank=rank, world=world,
                     nccl_id=ctx.nccl_id())
    if optimizer == "CG-LBFGS":
        s.set_lbfgs_k(lbfgs_k)
    if name == "c4l":
        s.use_ckernel_degridding(True)
    stream = torch.cuda.ExternalStream(s.eng.gvm_get_stream(s.engine_handle()), device=local)
    I_host = torch.from_numpy(s.get_image()).pin_memory()
    grad_host = torch.empty(2 * MN).pin_memory()
    s.set_iteration(1)   # priors are gated off at iteration 0 in the reference (src/functions.cu:4643)

    def step():          # ObjectiveFunction::calcFunction + calcGradient, image resident in HBM
        s.eval_device(1)

    def step_e2e():      # host image in, objective value + gradient out
        s.eval_host(I_host.data_ptr(), grad_host.data_ptr(), 1)

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(k):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return ctx.max_over_ranks(a.elapsed_time(b))

    for _ in range(warmup):
        step()
    sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    l0, c0 = s.launch_count(), s.collectives()
    ms_total = timed(step, steps)
    # dominant-kernel time of the LAST timed step: CUDA events recorded by the engine on its own
    # stream around each gradient-kernel launch inside the timed region
    kern_ms, kern_n = s.last_grad_kernel_ms()
    launches = s.launch_count() - l0
    collectives = s.collectives() - c0
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / steps
    # gridded mode: the objective runs over the gridded samples, not the raw visibilities
    Zwork = sum(s.h.gvmh_nvis(s.s, c) for c in range(problem.nchan)) if "-g" in cli else Ztot
    unit_work = (Zwork / 1e6) * (MN / 1e6)
    value = unit_work / (ms_step / 1e3)

    step_e2e()
    ms_e2e = timed(step_e2e, steps) / steps
    e2e_value = unit_work / (ms_e2e / 1e3)

    # what this run computed at the start image, iteration 1: identical (to fp32 summation order) at every N
    v_obj, fi_vals = s.calc_function()
    g_chk = s.calc_gradient(1).astype(np.float64)
    check = {"objective": float(v_obj), "half_chi2": float(fi_vals[0]), "grad_l2": float(np.sqrt((g_chk * g_chk).sum())),
             "grad_sum": float(g_chk.sum()), "at": "flat start image, iteration 1 (priors active)"}
    del g_chk

    # full reconstruction (the second half of the BASELINE metric): optimizer from the flat
    # starting image, setup/H2D excluded, wall time of Synthesizer::run
    recon = None
    if recon_iters > 0:
        s.clear_run()
        s.set_iteration(0)
        st0 = s.stats()
        img, sec = s.run()
        st = s.stats()
        sec = ctx.max_over_ranks(sec)
        recon = {"optimizer": optimizer, "iterations": int(s.scalars()["iterations_done"]), "seconds": sec,
                 "function_evals": st["function_evals"] - st0["function_evals"],
                 "gradient_evals": st["gradient_evals"] - st0["gradient_evals"],
                 "seconds_in_function_evals": st["function_s"] - st0["function_s"],
                 "seconds_in_gradient_evals": st["gradient_s"] - st0["gradient_s"],
                 "exit": s.exit_reason(), "lbfgs_k": lbfgs_k if optimizer == "CG-LBFGS" else None,
                 "setup_seconds": st["setup_s"]}
        path = REF_RECON_HANDOFF.format(cfg=name)
        if recon_reference and rank == 0 and os.path.exists(path):
            # left on this box by `bench.py --impl reference` (the driver runs that arm first): the reference CUDA
            # build's own reconstruction of the same workload
            try:
                r = np.load(path)
                same = int(r["iterations"]) == recon["iterations"] and r["image"].shape == img.shape
                recon["reference"] = {"seconds": float(r["seconds"]), "iterations": int(r["iterations"]),
                                      "speedup": float(r["seconds"]) / sec,
                                      "final_image_rel_l2": float(np.linalg.norm(img[0].astype(np.float64) - r["image"][0]) /
                                                                  np.linalg.norm(r["image"][0].astype(np.float64))) if same else None}
            except Exception as exc:
                recon["reference"] = {"error": repr(exc)}

    out = None
    if rank == 0:
        pk = peaks()
        mode = s.last_grad_mode()
        Zloc = s.local_nvis()
        ntiles, npx = s.grad_plan()
        if mode != 1:
            npx = MN
        flops = 4.0 * npx * Zloc                     # algorithmic: 2 FMA per (computed pixel, visibility) pair
        ach = flops / (kern_ms / 1e3) / 1e12 if kern_ms > 0 else None
        peak = pk["tflops_sustained"]
        st_all = s.stats()
        fp16x3 = os.environ.get("GVM_UMMA_SPLIT", "") == "fp16x3"
        umma_name = ("k_grad_umma (tcgen05 cta_group::2, three fp16 products)" if fp16x3 else
                     "k_grad_umma (tcgen05 cta_group::2, fp16 product + two 8-bit-float correction products)")
        umma_note = ("the fp16x3 split issues 3 fp16 MMAs per useful product, so frac <= 1/3 by construction" if fp16x3 else
                     "per useful product the kernel issues one fp16 MMA and one E4M3/E5M2 MMA of twice the K (both corrections), "
                     "i.e. 2 fp16-MMA times, so frac <= 1/2 by construction")
        roof = {"bound": "tensor", "kernel": {1: umma_name, 2: "k_grad_sep (CUDA cores fp32)",
                                              3: "k_grad_exact (CUDA cores)"}.get(mode, str(mode)),
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": (ach / peak) if ach else None,
                "traffic": None, "share_of_step": kern_ms / ms_step if ms_step > 0 else None,
                "plan": {"tiles": ntiles, "pixels_computed": npx, "pixels_image": MN},
                "note": f"algorithmic flops 4*P*Z per step on this rank = {flops:.3e}, P = pixels of the tiles that cover the "
                        f"unmasked part of the image (masked pixels are skipped, as DChi2 does); {kern_n} launch(es) per step "
                        f"(one per channel), {kern_ms:.2f} ms in total; peak = bf16/fp16 dense ({pk['source']}, sustained); "
                        + umma_note}
        if mode == 4:
            # gridded samples: scatter (28 B/sample in + 8 B RMW) + memset of the half plane (4 B/px) + cuFFT C2R
            # (two passes: 4 B/px in + ~8 B/px intermediate r/w + 4 B/px real out) — HBM-bound
            gbytes = 36.0 * Zloc + (4.0 + 16.0) * MN
            ach = gbytes / (kern_ms / 1e3) / 1e9 if kern_ms > 0 else None
            roof = {"bound": "hbm", "kernel": "k_gridfft_scatter + cuFFT C2R (gridded samples: the DFT is an FFT of the Hermitian half plane)",
                    "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": (ach / pk["hbm_gbs"]) if ach else None,
                    "traffic": None, "share_of_step": kern_ms / ms_step if ms_step > 0 else None,
                    "note": f"algorithmic bytes 36*Z + 20*M*N = {gbytes:.3e} per gradient on this rank, {kern_ms:.3f} ms; "
                            f"peak = measured HBM copy bandwidth ({pk['source']})"}
        if mode == 1 and name == "c2" and scale == 1.0 and world == 1:
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel on this exact workload,
            # from the committed ncu --set full capture (profiles/r2f_traffic.json; r1b_traffic.json for the fp16x3 split)
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", "r1b_traffic.json" if fp16x3 else "r2f_traffic.json")))
                roof["traffic"] = next(iter(tr.values()))["dram_bytes_total"]
                roof["traffic_note"] = ("bytes per launch (ncu); algorithmic minimum 28 B x Z per tile pass x 32 tiles, "
                                        "served from L2, + the split-K scratch slices (340 MB written once); DRAM at 0.03 % of peak")
            except Exception:
                pass
        cfg = workload_config(name, problem, wl)
        engine = {"samples_in_objective": Zwork,
                  "api": "C++ host layer (ObjectiveFunction::calcFunction + calcGradient) over the C ABI",
                  "sharding": ("replicated objective on the gridded samples (image-sized work does not shard); weighting + gridding "
                               f"distributed over {world} ranks" if ("-g" in cli and world > 1) else
                               f"visibility chunks over {world} rank(s)" if problem.nchan < world or world == 1
                               else f"channels over {world} rank(s) (i % world)"),
                  "l2": "inputs (>= 52 B/vis x Z) exceed the 126 MB L2; no flush needed", "grad_mode": mode}
        out = {"value": value, "ms_per_step": ms_step, "evals_per_s": 1e3 / ms_step,
               # the metric counts image pixels; masked pixels (noise >= noise_cut) are skipped by the
               # reference's DChi2 and by this engine alike — the same figure over the computed pixels only:
               "value_computed_pixels": value * (npx / MN), "config": cfg, "engine": engine,
               "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                       "h2d_bytes_per_step": 2 * MN * 4, "d2h_bytes_per_step": 2 * MN * 4 + 4,
                       "copies": "rank 0 only; the other ranks receive the image by ncclBroadcast" if world > 1 else "rank 0"},
               "gpu_launches": int(launches), "collectives": int(collectives), "clocks": clocks, "roofline": roof,
               "check": check, "recon": recon, "synth_seconds": t_gen,
               "preprocessing": {"raw_visibilities": Ztot, "samples_after_gridding": int(Zloc) if "-g" in cli else None,
                                 "weighting_seconds": st_all["weighting_s"], "gridding_seconds": st_all["gridding_s"],
                                 "scheme": scheme, "ckernel": f"{ckernel} {ck_size[0]}x{ck_size[1]}" if "-g" in cli else None,
                                 "gridding_Mvis_per_s": (Ztot / 1e6 / (st_all["weighting_s"] + st_all["gridding_s"]))
                                 if "-g" in cli and st_all["gridding_s"] > 0 else None}}
        if cpu:
            sc = s.scalars()
            from gpuvmem_b200.engine import beam_model
            pbf, pbc, pb = beam_model(problem.telescope, problem.antenna_diameter, float(problem.freqs.min()))
            meta = dict(deltau=sc["deltau"], deltav=sc["deltav"], fg_scale=sc["fg_scale"], pb_factor=pbf, pb_cutoff=pbc,
                        primary_beam=pb, xpix=sc["xobs_pix"], ypix=sc["yobs_pix"], nu_0=sc["nu_0"], noise_cut=sc["noise_cut"])
            out["cpu_baseline"] = cpu_baseline(problem, meta)
            try:
                out["cpu_gridding"] = cpu_gridding_baseline(problem, ckernel, ck_size, scheme)
            except Exception as exc:      # a reported baseline must not take the bench line down
                out["cpu_gridding"] = {"error": repr(exc)}
    s.close()
    del I_host, grad_host, problem
    return out


# the other named shapes of BASELINE.json, measured in the same process after the headline workload so that
# the driver's BENCH/SCALE records carry them at every GPU count: (config, scale, steps, recon iterations)
SIDE_CONFIGS = [("c1", 1.0, 20, 50), ("c3", 1.0, 3, 0), ("c4l", 0.1, 2, 0), ("c5", 0.25, 10, 10)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2")
    ap.add_argument("--scale", type=float, default=1.0, help="scale the visibility count (tests)")
    ap.add_argument("--grad-mode", type=int, default=0)
    ap.add_argument("--ref-sample", type=int, default=1000000, help="reference arm: visibilities per timed step")
    ap.add_argument("--no-full-check", dest="full_check", action="store_false",
                    help="reference arm: skip the single evaluation of the whole workload")
    ap.add_argument("--no-ref-recon", dest="ref_recon", action="store_false",
                    help="reference arm: skip the C1 reconstruction by the reference CUDA build")
    ap.add_argument("--full-eval-only", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--recon-only", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline workload only (skip the C1/C3/C4/C5 side measurements)")
    ap.add_argument("--recon-iters", type=int, default=10, help="optimizer iterations of the full-reconstruction leg (0: skip)")
    ap.add_argument("--lbfgs-k", type=int, default=10)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        if args.full_eval_only:
            return reference_full_eval(args)
        return run_reference(args)

    ctx = Ctx(args.gpus)
    main_out = measure(ctx, args.config, args.scale, args.steps, args.warmup, args.recon_iters, grad_mode=args.grad_mode,
                       lbfgs_k=args.lbfgs_k, sample_clocks=True, cpu=not args.no_cpu_baseline,
                       recon_reference=args.config == "c1")
    sides = {}
    if args.config == "c2" and args.scale == 1.0 and not args.no_configs:
        for name, scale, k, recon_it in SIDE_CONFIGS:
            key = name if scale == 1.0 else f"{name}_x{scale:g}"
            try:
                r = measure(ctx, name, scale, k, 3, recon_it, recon_reference=name == "c1")
            except Exception as exc:          # a side measurement must not take the headline line down
                r = {"error": repr(exc)}
            if ctx.rank == 0 and r is not None:
                r.pop("clocks", None)
                sides[key] = r
    if ctx.rank == 0:
        line = {"metric": METRIC, "value": main_out.pop("value"), "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main_out.pop("ms_per_step"), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
        line.update(main_out)
        if sides:
            line["configs"] = sides
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
