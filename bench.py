#!/usr/bin/env python
"""Benchmark of the power-spectrum hot path (BASELINE.json metric: P(k) cubes/s at 1024^3 fp32).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 1024]

A step is one ``power_spectrum_1d`` (15 log bins, return_n_modes) over one synthetic toy-ionised
21-cm brightness-temperature cube.  With N > 1 (torchrun, one rank per GPU) every rank transforms its
own cube -- the "batches of coeval boxes, one per GPU" partition of the north star, no data-path
collective -- and ``value`` is all cubes divided by the slowest rank's time ("scaling": "weak").

Printed JSON (one line, rank 0): the driver contract plus
  roofline      dominant kernel: algorithmic bytes per launch / CUDA-event duration vs the measured HBM peak
  passes        the same for every pass, and for the whole pipeline (B_auto = 4N^3 + 4*8N^2(N/2+1) bytes)
  e2e           the same call fed from pinned HOST memory, H2D copy and result read-back inside the timing
  cpu_baseline  the CPU oracle (numpy/scipy port of the reference, 1 thread like the reference) on a bounded sample

``--impl reference`` times that CPU port alone (rank 0 only) on the same metric/config.
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

BOX_MPC = 500.0 / 0.7
KBINS = 15


def algorithmic_bytes(n):
    c = 8 * n * n * (n // 2 + 1)
    return {"z": 4 * n ** 3 + c, "y": 2 * c, "x": c, "pipeline": 4 * n ** 3 + 4 * c}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_reference():
    """The UNMODIFIED reference's power_spectrum module from baseline/_ref (installed by baseline/install_reference.sh,
    git-ignored, shipped to the GPU box), imported through a stub package: tools21cm/__init__.py pulls in astropy /
    scikit-image, which this image lacks and which the power-spectrum path does not need.  None when absent."""
    import types
    d = os.path.join(ROOT, "baseline", "_ref", "tools21cm")
    if not os.path.isfile(os.path.join(d, "power_spectrum.py")):
        return None
    if "tools21cm" not in sys.modules:
        pkg = types.ModuleType("tools21cm")
        pkg.__path__ = [d]
        sys.modules["tools21cm"] = pkg
    try:
        import tools21cm.power_spectrum as ref
        return ref
    except Exception:
        return None


def cpu_reference_run(n, cube=None):
    """One power_spectrum_1d of the bench workload on an n^3 cube on the host: the reference itself when
    baseline/_ref is there ("reference"), else the oracle port ("port").  The two halves of the path are timed
    through the reference's own public functions: power_spectrum_nd (FFT, O(N^3 log N)) and radial_average
    (k grids + two np.histogram passes, O(N^3)).  Returns (kind, seconds_fft, seconds_binning, result)."""
    import numpy as np
    from tools21cm_b200 import synthetic
    ref = load_reference()
    kind = "reference"
    if ref is None:
        from oracle import ps_oracle as ref          # the checker, timed as the CPU baseline -- never on the product path
        kind = "port"
    if cube is None:
        cube = synthetic.toy_dt_cube((n, n, n), seed=1235)
    with np.errstate(all="ignore"):                   # tools21cm/__init__.py:94-95 does the same globally
        t0 = time.perf_counter()
        nd = ref.power_spectrum_nd(cube, box_dims=[BOX_MPC] * 3)
        t1 = time.perf_counter()
        ps, ks, nm = ref.radial_average(nd, [BOX_MPC] * 3, kbins=KBINS, binning="log")
        t2 = time.perf_counter()
    return kind, t1 - t0, t2 - t1, (ps, ks, nm)


def scale_to(n, sample, t_fft, t_bin):
    """seconds for an n^3 cube from a sample^3 measurement: the FFT part scales as N^3 log N, the binning part
    (k grids, histograms: 2/3 of the time) as N^3"""
    import math
    r = (n / sample) ** 3
    return t_fft * r * math.log2(n) / math.log2(sample) + t_bin * r


CPU_SAMPLE = 512     # ONE sample rule for both legs (the 1024^3 cube needs ~70 GB and minutes on the CPU path)


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores (one thread:
    scipy.fftpack and np.histogram are single-threaded).  Each step is one power_spectrum_1d on a 512^3 sample of
    the workload (the size is in `config` and `cpu_baseline.sample`); `value` scales it to the metric's 1024^3."""
    if rank != 0:
        return
    n = args.size
    sample = min(CPU_SAMPLE, n)
    for _ in range(args.warmup):
        cpu_reference_run(sample)
    tf = tb = 0.0
    kind = "port"
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        kind, a, b, _res = cpu_reference_run(sample)
        tf += a
        tb += b
    steps = max(1, args.steps)
    wall = (time.perf_counter() - t0) / steps
    tf /= steps
    tb /= steps
    est = scale_to(n, sample, tf, tb)
    value = 1.0 / est
    desc = ("%s power_spectrum_1d on one %d^3 cube per step: %.2f s measured (FFT %.2f s + binning %.2f s; %.2f s with "
            "input generation); scaled to %d^3 as FFT x N^3 log N + binning x N^3 = %.1f s"
            % ("tools21cm v2.4.0 (baseline/_ref)" if kind == "reference" else "oracle port of tools21cm",
               sample, tf + tb, tf, tb, wall, n, est))
    cfg = workload_config(n)
    cfg["reference_sample_cube"] = sample
    line = {
        "impl": "reference", "metric": "P(k) cubes/sec at %d^3 fp32" % n, "value": value, "unit": "cubes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": est * 1e3,
        "measured_ms_per_sample_step": (tf + tb) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "cubes/s", "cores": 1, "kind": kind, "sample": desc,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": "cubes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def nccl_summary():
    """What NCCL logged at init (NCCL_DEBUG_FILE of this process): communicator size and whether NVLS is in use."""
    import glob, re
    out = {"nranks": None, "nvls": None, "log": None}
    pat = os.environ.get("NCCL_DEBUG_FILE", "").replace("%h", "*").replace("%p", str(os.getpid()))
    found = glob.glob(pat)
    if not found:
        print("no NCCL log at %s (NCCL_DEBUG=%s)" % (pat, os.environ.get("NCCL_DEBUG")), file=sys.stderr)
    for path in found:
        try:
            txt = open(path, errors="replace").read()
        except OSError:
            continue
        out["log"] = path
        m = re.findall(r"nranks (\d+)", txt)
        if m:
            out["nranks"] = max(int(v) for v in m)
        out["nvls"] = bool(re.search(r"NVLS", txt))
        for ln in txt.splitlines():
            if "nranks" in ln or "NVLS" in ln or "NCCL version" in ln:
                print(ln, file=sys.stderr)
    return out


def gather_slabs(slab, dist, world, rank):
    """The full cube on rank 0 from the ranks' x slabs (for the parity figures of the sharded legs)."""
    import torch
    parts = [torch.empty_like(slab) for _ in range(world)] if rank == 0 else None
    dist.gather(slab, parts, dst=0)
    return torch.cat(parts, dim=0) if rank == 0 else None


def sharded_leg(args, D, t2c, lib, synthetic, dist, dev, rank, world, barrier, e0, e1):
    import ctypes as C
    import numpy as np
    import torch
    ns = args.slab_size or args.size
    nxl = ns // world
    slabs = [synthetic.toy_dt_cube_torch((nxl, ns, ns), seed=4000 + rank, device=dev)]
    cross = args.slab_cross or args.slab_nd
    if cross:
        slabs.append(synthetic.gaussian_cube_torch((nxl, ns, ns), seed=5000 + rank, device=dev))

    def slab_step():
        if args.slab_nd:
            rows, yidx = D.sharded_cross_power_spectrum_nd(slabs[0], slabs[1], BOX_MPC, copy=False)
            return None, None, np.array([float(rows.numel())])
        if cross:
            return D.sharded_cross_power_spectrum_1d(slabs[0], slabs[1], kbins=KBINS, box_dims=BOX_MPC, return_n_modes=True)
        return D.sharded_power_spectrum_1d(slabs[0], kbins=KBINS, box_dims=BOX_MPC, return_n_modes=True)

    def timed(n_steps):
        for _ in range(3):
            r = slab_step()
        barrier()
        e0.record()
        for _ in range(n_steps):
            r = slab_step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n_steps, r

    nf = 2 if cross else 1
    cbytes = 8 * ns * ns * (ns // 2 + 1)
    a_bytes = (world - 1) / world ** 2 * cbytes * nf           # per GPU and direction, all fields
    out = {"workload": "%s of ONE %d^3 fp32 cube held as %d x slabs" %
                       ("cross_power_spectrum_nd (output sharded by y rows)" if args.slab_nd else
                        "cross_power_spectrum_1d" if cross else "power_spectrum_1d", ns, world),
           "unit": "cubes/s", "scaling": "strong", "nvlink_bytes_per_gpu_per_direction": a_bytes}
    results = {}
    for mode in ("p2p", "nccl"):
        os.environ["T21_EXCHANGE"] = mode
        ms, res = timed(args.steps)
        results[mode] = res
        entry = {"ms_per_step": ms, "value": 1e3 / ms}
        if mode == "p2p":
            # stage 1 of the fused path alone: Z pass + Y pass whose stores ARE the exchange (CUDA events of the library)
            buf = (C.c_float * 5)()
            if True:
                lib.t21_profile(1)
                acc = np.zeros(5)
                for _ in range(3):
                    slab_step()
                    lib.t21_profile_last(buf, 5)
                    acc += np.array(list(buf))
                lib.t21_profile(0)
                acc /= 3.0
                t = torch.tensor(acc, device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                acc = t.cpu().numpy()
                if True:
                    gbs1 = a_bytes / nf / 1e9 / (acc[2] / 1e3) if acc[2] > 0 else None      # (the profile holds the last field)
                    entry["fused_y_pass"] = {
                        "ms_per_field": float(acc[2]), "z_ms_per_field": float(acc[1]), "achieved_gbs": gbs1,
                        "frac_of_770_measured_peer": gbs1 / 770.0 if gbs1 else None,
                        "frac_of_900_nominal": gbs1 / 900.0 if gbs1 else None,
                        "what": "Y pass with SEND=2 (its stores go straight to the owners' y slabs over NVLink), one field, "
                                "CUDA events of the library around the pass; bytes = A(N,P) of SURVEY.md 8d per field"}
                    entry["x_pass_ms"] = float(acc[3])
        elif mode == "nccl":
            # the NCCL exchange alone (one field, no overlap with compute)
            be = D.CudaBackend()
            send = be.stage1(slabs[0], None, send_world=world)
            for _ in range(2):
                D.all_to_all_packed(send)
            barrier()
            e0.record()
            for _ in range(args.steps):
                D.all_to_all_packed(send)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            a2a = float(t.item())
            del send
            gbs = a_bytes / nf / 1e9 / (a2a / 1e3)
            entry["all_to_all_alone"] = {"ms": a2a, "achieved_gbs": gbs, "frac_of_770_measured_peer": gbs / 770.0,
                                         "frac_of_900_nominal": gbs / 900.0,
                                         "what": "all_to_all_single over NCCL of ONE field's packed send buffer, timed alone"}
        out[mode] = entry
    os.environ["T21_EXCHANGE"] = "p2p"
    out["value"] = out["p2p"]["value"]
    out["ms_per_step"] = out["p2p"]["ms_per_step"]
    # parity: the same cube on ONE GPU (rank 0 gathers the slabs) against both exchanges
    if not args.slab_nd and ns <= 1024:
        full = [gather_slabs(sl, dist, world, rank) for sl in slabs]
        if rank == 0:
            if cross:
                ref = t2c.cross_power_spectrum_1d(full[0], full[1], kbins=KBINS, box_dims=BOX_MPC, return_n_modes=True)
            else:
                ref = t2c.power_spectrum_1d(full[0], kbins=KBINS, box_dims=BOX_MPC, return_n_modes=True)
            par = {}
            for mode in ("p2p", "nccl"):
                got = results[mode]
                with np.errstate(all="ignore"):
                    rel = np.abs(got[0] - ref[0]) / np.abs(ref[0])
                par[mode] = {"n_modes_equal": bool(np.array_equal(got[2], ref[2])), "max_rel": float(np.nanmax(rel))}
            out["parity"] = dict(par, against="the single-GPU estimator on the gathered cube (rank 0)")
        del full
    elif not args.slab_nd:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import lattice                                   # exact per-bin mode counts without an FFT (a checker)
        want = lattice.shell_counts((ns,) * 3, [BOX_MPC] * 3, lattice.reference_edges((ns,) * 3, [BOX_MPC] * 3, KBINS))
        out["parity"] = {m: {"n_modes_equal": bool(np.array_equal(results[m][2], want))} for m in ("p2p", "nccl")}
        out["parity"]["against"] = "exact lattice mode counts (tests/lattice.py); the cube is too large to gather"
    out["n_modes_total"] = float(np.sum(results["p2p"][2]))
    return out


def config3_leg(args, D, t2c, synthetic, dist, dev, rank, world, barrier):
    """BASELINE config 3: cross_power_spectrum_1d of 64 pairs of 512^3 (density, xfrac) cubes, dealt to the ranks."""
    import numpy as np
    import torch
    npairs, n = 64, 512
    mine = D.partition(npairs, world, rank)
    pairs = {}
    for i in mine:
        g = torch.Generator(device=dev).manual_seed(2000 + i)
        delta = 0.3 * torch.randn((n, n, n), generator=g, device=dev)
        xf = (torch.rand((n, n, n), generator=g, device=dev) < 0.3).to(torch.float32)
        pairs[i] = ((1.0 + delta), xf)

    def fn(i):
        a, b = pairs[i]
        return t2c.cross_power_spectrum_1d(a, b, kbins=KBINS, box_dims=244 / 0.7, return_n_modes=True)
    D.batched(fn, list(range(npairs)))
    barrier()
    t0 = time.perf_counter()
    res = D.batched(fn, list(range(npairs)))
    torch.cuda.synchronize()
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    b_pair = 2 * (4 * n ** 3 + 4 * 8 * n * n * (n // 2 + 1))
    return {"workload": "cross_power_spectrum_1d, 64 pairs of 512^3 fp32 (1+delta, mask) cubes, pair i on rank i mod %d, "
                        "results gathered on every rank" % world,
            "pairs_per_s": npairs / dt, "ms_total": dt * 1e3, "ms_per_pair_per_gpu": dt * 1e3 / (npairs / world),
            "hbm_gbs_per_gpu": b_pair * (npairs / world) / dt / 1e9,
            "n_modes_total_pair0": float(np.sum(res[0][2])), "timing": "wall clock around the whole batch, max over ranks"}


def config5_leg(args, D, lib, synthetic, dist, dev, rank, world, barrier):
    """BASELINE config 5: cross_power_spectrum_nd of a 2048^3 pair held as x slabs, output left sharded by y rows."""
    import ctypes as C
    import numpy as np
    import torch
    n = args.c5_size
    nxl = n // world
    free, _ = torch.cuda.mem_get_info()
    need = 2 * nxl * n * n * 4 * 6
    if free < need:
        return {"skipped": "needs %.0f GB per GPU, %.0f free" % (need / 2 ** 30, free / 2 ** 30)}
    g = torch.Generator(device=dev).manual_seed(7000 + rank)
    a = torch.randn((nxl, n, n), generator=g, device=dev)
    b = torch.randn((nxl, n, n), generator=g, device=dev)
    out = {"workload": "cross_power_spectrum_nd, one pair of %d^3 fp32 cubes as %d x slabs, fftshifted fp32 output left sharded "
                       "by rows of y" % (n, world)}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cbytes = 8 * n * n * (n // 2 + 1)
    a_bytes = 2 * (world - 1) / world ** 2 * cbytes
    for mode in ("p2p", "nccl"):
        os.environ["T21_EXCHANGE"] = mode
        try:
            for _ in range(2):
                rows, yidx = D.sharded_cross_power_spectrum_nd(a, b, 714.0, copy=False)
            barrier()
            e0.record()
            reps = 3
            for _ in range(reps):
                rows, yidx = D.sharded_cross_power_spectrum_nd(a, b, 714.0, copy=False)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            # Parseval over the sharded output: sum of the fftshifted power == N^3 * sum(a * b) * scale
            s = rows.double().sum().reshape(1)
            dot = (a.double() * b.double()).sum().reshape(1)
            dist.all_reduce(s)
            dist.all_reduce(dot)
            scale = (714.0 ** 3 / n ** 3) ** 2 / 714.0 ** 3
            expect = float(dot.item()) * n ** 3 * scale
            ref = float((a.double() ** 2).sum().item()) * world * n ** 3 * scale       # auto-power level
            out[mode] = {"ms_per_pair": ms, "pairs_per_s": 1e3 / ms,
                         "parseval_rel_err_vs_auto_level": abs(float(s.item()) - expect) / ref,
                         "rows_shape": list(rows.shape)}
            del rows
            if mode == "p2p":
                # per-stage CUDA events of the library (max over ranks): stage 1 of the LAST field (Z pass, then the Y
                # pass whose stores are the exchange) and the X pass of stage 2 over both fields
                buf = (C.c_float * 5)()
                lib.t21_profile(1)
                acc = np.zeros(5)
                for _ in range(2):
                    D.sharded_cross_power_spectrum_nd(a, b, 714.0, copy=False)
                    lib.t21_profile_last(buf, 5)
                    acc += np.array(list(buf))
                lib.t21_profile(0)
                tt = torch.tensor(acc / 2.0, device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                acc = tt.cpu().numpy()
                if acc[2] > 0:
                    gbs = a_bytes / 2 / 1e9 / (acc[2] / 1e3)
                    out[mode]["stages_ms"] = {"z_per_field": float(acc[1]), "y_fused_exchange_per_field": float(acc[2]),
                                              "barrier_plus_x_both_fields": float(acc[3]),
                                              "rest_assembly_and_barriers": float(ms - 2 * (acc[1] + acc[2]) - acc[3])}
                    out[mode]["all_to_all_phase"] = {
                        "ms_both_fields": float(2 * acc[2]), "achieved_gbs": gbs, "frac_of_770_measured_peer": gbs / 770.0,
                        "frac_of_900_nominal": gbs / 900.0,
                        "what": "the Y pass with SEND=2 IS the all-to-all (its stores go to the owners' y slabs over NVLink); "
                                "bytes = A(N,P) per field (SURVEY.md 8d)"}
        except Exception as exc:                       # keep the bench line alive; say what failed
            out[mode] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
        torch.cuda.empty_cache()
    os.environ["T21_EXCHANGE"] = "p2p"
    out["nvlink_bytes_per_gpu_per_direction"] = a_bytes
    out["target"] = "all-to-all phase <= %.1f ms for 50 %% of 900 GB/s (SURVEY.md 8d)" % (a_bytes / 1e9 / 0.45)
    return out


def workload_config(n):
    return {"workload": "power_spectrum_1d, %d^3 fp32 toy-ionised dT cube, box_dims=500/0.7 cMpc, "
                        "15 log kbins, return_n_modes" % n,
            "cube": n, "kbins": KBINS, "binning": "log", "partition": "one cube per GPU, no collective",
            "l2": "inputs (%.2f GB per cube) are larger than the 126 MB L2; no explicit flush" % (4 * n ** 3 / 1e9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-slab", action="store_true", help="skip the slab-sharded leg of multi-GPU runs")
    ap.add_argument("--slab-size", type=int, default=0, help="cube size of the slab-sharded leg (default: --size)")
    ap.add_argument("--slab-cross", action="store_true", help="slab-sharded leg: cross spectrum of two cubes")
    ap.add_argument("--slab-nd", action="store_true",
                    help="slab-sharded leg: cross_power_spectrum_nd, output left sharded (BASELINE.json config 5)")
    ap.add_argument("--configs35", action="store_true", help="run the config 3 / config 5 legs with fewer than 8 ranks too")
    ap.add_argument("--c5-size", type=int, default=2048, help="cube size of the config 5 leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and args.impl == "ours":
        # NCCL's log goes to a file per rank (stdout carries the JSON line), set before torch loads NCCL; rank 0
        # echoes the communicator lines to stderr at the end and reports nranks / NVLS in the line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "INFO"             # the communicator lines (nranks, NVLS) are INFO level
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/t21_nccl_%p.log")

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import tools21cm_b200 as t2c
    from tools21cm_b200 import _lib, synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.size
    lib = _lib.load()
    x = synthetic.toy_dt_cube_torch((n, n, n), seed=1235 + rank, device=dev)

    def step(inp):
        return t2c.power_spectrum_1d(inp, kbins=KBINS, box_dims=BOX_MPC, return_n_modes=True)

    for _ in range(args.warmup):
        res = step(x)
    barrier()

    # ---- timed region: inputs resident in HBM ------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.t21_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        res = step(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = lib.t21_launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = world * args.steps / (ms / 1e3)

    # ---- per-pass CUDA-event timing (same steps, events recorded between passes) ----------
    import ctypes as C
    lib.t21_profile(1)
    acc = np.zeros(5)
    buf = (C.c_float * 5)()
    for _ in range(args.steps):
        step(x)
        lib.t21_profile_last(buf, 5)
        acc += np.array(list(buf))
    lib.t21_profile(0)
    pass_ms = dict(zip(["mean", "z", "y", "x", "reduce"], (acc / args.steps).tolist()))
    ab = algorithmic_bytes(n)
    peak, peak_src = measured_peak()
    passes = {}
    for k in ("z", "y", "x"):
        gbs = ab[k] / 1e9 / (pass_ms[k] / 1e3) if pass_ms[k] > 0 else 0.0
        passes[k] = {"ms": pass_ms[k], "bytes": ab[k], "achieved_gbs": gbs, "frac": gbs / peak}
    passes["mean"] = {"ms": pass_ms["mean"]}
    passes["reduce"] = {"ms": pass_ms["reduce"]}
    pipe_gbs = ab["pipeline"] / 1e9 / (ms_per_step / 1e3)
    passes["pipeline"] = {"ms": ms_per_step, "bytes": ab["pipeline"], "achieved_gbs": pipe_gbs,
                          "frac": pipe_gbs / peak, "frac_of_8TBs_nominal": pipe_gbs / 8000.0}
    dom = max(("z", "y", "x"), key=lambda k: pass_ms[k])
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("%s_%d" % (dom, n))
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": {"z": "ZPass (tile_kernel)", "y": "YPass (tile_kernel)",
                                           "x": "XPass (xpass_kernel)"}[dom],
                "achieved": passes[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": passes[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "share_of_step": pass_ms[dom] / max(1e-9, sum(pass_ms.values()))}

    # ---- end to end through the public API with HOST buffers: H2D + estimator + D2H inside the timing ----
    # numpy_in: the reference's calling convention (a pageable numpy array; staged through the library's pinned ring);
    # pinned_in: a pinned torch tensor (one DMA).  The headline e2e is numpy_in.
    e2e = None
    if not args.no_e2e:
        host = torch.empty((n, n, n), dtype=torch.float32, pin_memory=True)
        host.copy_(x)
        host_np = np.empty((n, n, n), dtype=np.float32)
        host_np[...] = host.numpy()
        del x
        torch.cuda.empty_cache()

        def timed_e2e(inp):
            for _ in range(2):
                step(inp)
            barrier()
            t0 = time.perf_counter()
            e0.record()
            for _ in range(args.steps):
                step(inp)                  # H2D of the cube + kernels + D2H of (ps, n_modes)
            e1.record()
            torch.cuda.synchronize()
            ms_e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
            if world > 1:
                t = torch.tensor([ms_e], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_e = float(t.item())
            return ms_e
        ms_np = timed_e2e(host_np)
        ms_pin = timed_e2e(host)
        e2e = {"value": world * args.steps / (ms_np / 1e3), "unit": "cubes/s",
               "h2d_bytes_per_step": 4 * n ** 3, "d2h_bytes_per_step": 2 * KBINS * 8,
               "ms_per_step": ms_np / args.steps, "input": "pageable numpy array (the reference's calling convention)",
               "pinned_in": {"value": world * args.steps / (ms_pin / 1e3), "ms_per_step": ms_pin / args.steps,
                             "input": "pinned torch tensor"},
               "note": "every rank copies its own cube over the host's PCIe links and memory: with several ranks the host "
                       "side is shared, so per-GPU end-to-end throughput drops"}
        del host_np

    # ---- one cube sharded over all ranks as x slabs (BASELINE config 4 "1 vs 2/4/8 B200 slab-sharded") ----
    sharded = None
    if world > 1 and not args.no_slab:
        from tools21cm_b200 import distributed as D
        try:
            del host
        except NameError:
            pass
        try:
            del x
        except NameError:
            pass
        torch.cuda.empty_cache()
        sharded = sharded_leg(args, D, t2c, lib, synthetic, dist, dev, rank, world, barrier, e0, e1)

    # ---- BASELINE configs 3 and 5 (only with 8 ranks, or when asked for) ----
    config3 = config5 = None
    if world > 1 and (world == 8 or args.configs35) and not args.no_slab:
        from tools21cm_b200 import distributed as D
        torch.cuda.empty_cache()
        config3 = config3_leg(args, D, t2c, synthetic, dist, dev, rank, world, barrier)
        torch.cuda.empty_cache()
        config5 = config5_leg(args, D, lib, synthetic, dist, dev, rank, world, barrier)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # the reference on the host cores next to the GPU number, and the GPU result of the SAME cube beside it
        sample = min(CPU_SAMPLE, n)
        cube = synthetic.toy_dt_cube((sample,) * 3, seed=1235)
        kind, tf, tb, (ps_c, ks_c, nm_c) = cpu_reference_run(sample, cube)
        ps_g, ks_g, nm_g = t2c.power_spectrum_1d(cube, kbins=KBINS, box_dims=BOX_MPC, return_n_modes=True)
        with np.errstate(all="ignore"):
            rel = np.abs(ps_g - ps_c) / np.abs(ps_c)
        est = scale_to(n, sample, tf, tb)
        cpu = {"value": 1.0 / est, "unit": "cubes/s", "cores": 1, "kind": kind,
               "sample": "%s power_spectrum_1d on one %d^3 cube: %.2f s measured (FFT %.2f s + binning %.2f s), scaled to "
                         "%d^3 as FFT x N^3 log N + binning x N^3 = %.1f s (the full cube needs ~70 GB and minutes on the "
                         "CPU path)" % ("tools21cm v2.4.0 (baseline/_ref)" if kind == "reference" else "oracle port",
                                        sample, tf + tb, tf, tb, n, est),
               "measured_seconds_at_sample": tf + tb, "sample_cube": sample, "host_cores": os.cpu_count(),
               "parity": {"cube": sample, "n_modes_equal": bool(np.array_equal(nm_g, nm_c)),
                          "k_centres_equal": bool(np.array_equal(ks_g, ks_c)),
                          "max_rel_err": float(np.nanmax(rel)), "tolerance": 1e-5}}

    if rank == 0:
        line = {
            "metric": "P(k) cubes/sec at %d^3 fp32" % n, "value": value, "unit": "cubes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(n), "clocks": clocks, "gpu_launches": int(launches),
            "roofline": roofline, "passes": passes, "e2e": e2e, "cpu_baseline": cpu, "sharded": sharded,
            "config3": config3, "config5": config5, "nccl": nccl_summary() if world > 1 else None,
            "n_modes_total": float(np.sum(res[2])),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
