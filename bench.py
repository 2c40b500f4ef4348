#!/usr/bin/env python
"""bench.py -- depth-maps/s of the PatchmatchNet cascade with the B200-native PatchMatch hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo (one JSON line on rank 0)
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on the host CPU

Workload (BASELINE.json configs[1]): 1 reference + 4 source views, 640x512, full 3-stage cascade with
64/32 -> 16/16 -> 8 hypotheses, one reference view per GPU (weak scaling: N GPUs -> N depth maps per step,
no collective on the data path).  A "step" is one full forward over one batch of synthetic input.

Numbers on the JSON line
    value      depth-maps/s, inputs resident in HBM, forward replayed from a CUDA graph, L2 flushed
               between timed steps, device time by CUDA events, max over ranks
    e2e        same metric through the public API (DepthEngine.infer_stream): pinned HOST inputs copied
               to the device and results copied back inside the timed region, every step
    roofline   the dominant hand-written kernel (fused warp+correlation, all source views per launch):
               algorithmic bytes per launch / CUDA-event duration of that launch, against the measured
               HBM copy bandwidth in MEASURED_PEAKS.json
    cpu_baseline  the oracle (CPU restatement of the reference, bit-identical to it) timed on the host
    gpu_eager_reference  the same oracle in eager PyTorch on the GPU, issuing the reference's own op sequence
               (what the unmodified reference would do on this B200; SURVEY.md 8d "honest bar")
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from patchmatchnet_b200 import synthetic  # noqa: E402

WEIGHTS = os.path.join(REPO, "tests", "golden", "weights_000007.pt")
METRIC = "depth-maps/sec (1ref+4src, 640x512)"
UNIT = "depth-maps/s"


def build_net(patchmatch_cls=None):
    from patchmatchnet_b200.net import PatchmatchNet, load_reference_state

    kw = dict(synthetic.DEFAULT_NET_KWARGS)
    net = PatchmatchNet(**kw) if patchmatch_cls is None else PatchmatchNet(**kw, patchmatch_cls=patchmatch_cls)
    if os.path.exists(WEIGHTS):
        load_reference_state(net, torch.load(WEIGHTS, map_location="cpu"))
        weights = "reference checkpoint params_000007 (fixture)"
    else:  # random init; give the zero-initialised offset convs something to do
        g = torch.Generator().manual_seed(0)
        for m in (net.patchmatch_1, net.patchmatch_2, net.patchmatch_3):
            for conv in (m.propa_conv, m.eval_conv):
                conv.weight.data.normal_(0, 0.05, generator=g)
        weights = "random init"
    return net.eval(), weights


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.

    The timed region of the default run is ~20 ms, far shorter than the start-up of an `nvidia-smi -lms` process, so the
    primary sampler is an NVML thread (pynvml / nvidia-ml-py: one clock + one reasons query every ~2 ms, GIL released
    inside the library calls); `nvidia-smi -lms 20`, started BEFORE the warm-up so that it is already printing, is the
    fallback when NVML cannot be loaded.  Only samples taken between mark_begin() and mark_end() are reported; when the
    window caught none (nvidia-smi fallback on a very short run) the samples of the warm-up just before it are used and
    the line says so."""

    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, gpu_index: int, device_uuid: str = ""):
        self.gpu = gpu_index
        self.uuid = device_uuid
        self.proc = None
        self.path = None
        self.thread = None
        self.samples = []  # (perf_counter, sm_mhz, reasons tuple)   [NVML thread]
        self.max_mhz = None
        self._stop = False
        self.t_begin = self.t_end = None
        self.wall_begin = self.wall_end = None

    # ---- NVML thread -------------------------------------------------------------------------------------------
    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        if self.uuid:
            for cand in (self.uuid, "GPU-" + self.uuid):
                for arg in (cand, cand.encode()):  # str or bytes, depending on the nvidia-ml-py version
                    try:
                        return pynvml, pynvml.nvmlDeviceGetHandleByUUID(arg)
                    except Exception:  # noqa: BLE001
                        continue
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.gpu)

    def _run_nvml(self, nv, h):
        bits = ((nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"), (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"), (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap"))
        while not self._stop:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.samples.append((time.perf_counter(), mhz, tuple(n for b, n in bits if mask & b)))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import threading

            nv, h = self._nvml_handle()
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run_nvml, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL,
            )
        except Exception:  # noqa: BLE001
            self.proc = None

    def mark_begin(self):
        self.t_begin, self.wall_begin = time.perf_counter(), time.time()

    def mark_end(self):
        self.t_end, self.wall_end = time.perf_counter(), time.time()

    # ---- results -----------------------------------------------------------------------------------------------
    @staticmethod
    def _summary(rows, max_mhz, how, window):
        """rows: (sm_mhz, reasons tuple)"""
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": max_mhz, "reasons": ["no samples"], "how": how}
        reasons = sorted({r for _, rs in rows for r in rs})
        return {"sm_mhz": statistics.median(m for m, _ in rows), "sm_max_mhz": max_mhz, "reasons": reasons, "samples": len(rows),
                "how": how, "window": window}

    def stop(self):
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            lo = self.t_begin if self.t_begin is not None else -1.0
            hi = self.t_end if self.t_end is not None else float("inf")
            inside = [(m, r) for t, m, r in self.samples if lo <= t <= hi]
            if inside:
                return self._summary(inside, self.max_mhz, "NVML thread, 2 ms period", "timed region")
            return self._summary([(m, r) for _, m, r in self.samples], self.max_mhz, "NVML thread, 2 ms period", "warm-up + timed region")
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML and nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        rows, mx = [], []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    stamp = time.mktime(time.strptime(f[0].split(".")[0], "%Y/%m/%d %H:%M:%S")) + float("0." + f[0].split(".")[1])
                    mhz = float(f[1])
                    mx.append(float(f[2]))
                except (ValueError, IndexError):
                    continue
                rows.append((stamp, mhz, tuple(nm for nm, val in zip(self.NAMES, f[5:9]) if val.lower().startswith("active"))))
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        top = max(mx) if mx else None
        if self.wall_begin is not None and self.wall_end is not None:
            inside = [(m, r) for t, m, r in rows if self.wall_begin - 0.02 <= t <= self.wall_end + 0.02]
            if inside:
                return self._summary(inside, top, "nvidia-smi -lms 20", "timed region")
        return self._summary([(m, r) for _, m, r in rows], top, "nvidia-smi -lms 20", "warm-up + timed region")


# ----------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference, host cores)
# ----------------------------------------------------------------------------------------------


def cpu_forward_timer(height: int, width: int, n_views: int):
    """The oracle port of the reference on the host CPU.  torch's default is one thread per core; on a
    many-core host the small ops of this network get SLOWER with every core added, so the thread count is
    calibrated first (one forward per candidate, ascending, stop when it stops paying) and the reference is
    given its best configuration."""
    from oracle.pm_oracle import PatchMatchOracle  # allowed here: cpu_baseline / --impl reference legs only

    cores = os.cpu_count() or 1
    net, _ = build_net(PatchMatchOracle)
    net.stack_views = False  # the reference runs FeatureNet view by view (net.py:203-208)
    inp = synthetic.make_inputs(1, n_views, height, width, seed=0)

    def step():
        with torch.no_grad():
            t0 = time.perf_counter()
            net([i.clone() for i in inp["images"]], inp["intrinsics"].clone(), inp["extrinsics"].clone(), inp["depth_min"], inp["depth_max"])
            return time.perf_counter() - t0

    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores})
    best_t, best_c, tried = None, cands[0], {}
    for c in cands:
        torch.set_num_threads(c)
        step()  # warm-up at this thread count
        t = step()
        tried[c] = round(t, 3)
        if best_t is None or t < best_t:
            best_t, best_c = t, c
        elif t > 1.5 * best_t:
            break
    torch.set_num_threads(best_c)
    step.calibration = tried
    return step, best_c


def gpu_eager_reference(height: int, width: int, n_views: int, device, warmup: int = 3, iters: int = 10):
    """The reference's algorithm as the reference would run it on THIS GPU: the oracle port (bit-identical to the
    reference on CPU) in eager PyTorch, the caller-side shell issuing the reference's own op sequence (conv, BatchNorm,
    ReLU separately, view by view; no folding, no stacking), cudnn.benchmark on as in eval.py:301, TF32 as torch's
    defaults leave it.  A reported baseline (SURVEY.md 8d "honest bar"), never part of the product path."""
    from oracle.pm_oracle import PatchMatchOracle  # allowed here: baseline legs only
    import patchmatchnet_b200.net as net_mod

    net, _ = build_net(PatchMatchOracle)
    net = net.to(device).eval()
    net.stack_views = False
    inp = synthetic.make_inputs(1, n_views, height, width, seed=0)
    imgs = [i.to(device) for i in inp["images"]]
    K, E = inp["intrinsics"].to(device), inp["extrinsics"].to(device)
    dmin, dmax = inp["depth_min"].to(device), inp["depth_max"].to(device)
    net_mod.LIBRARY_FAST_PATH = False
    try:
        with torch.no_grad():
            for _ in range(warmup):
                net([i.clone() for i in imgs], K.clone(), E.clone(), dmin, dmax)
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            for _ in range(iters):
                net([i.clone() for i in imgs], K.clone(), E.clone(), dmin, dmax)
            torch.cuda.synchronize(device)
            dt = time.perf_counter() - t0
    finally:
        net_mod.LIBRARY_FAST_PATH = True
    return {"value": iters / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / iters, "kind": "port",
            "how": f"oracle port of the reference in eager PyTorch on the same GPU, reference op sequence, inputs resident, "
                   f"{warmup} warm-up + {iters} timed forwards, synchronize-bracketed wall clock, cudnn.benchmark=True"}


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, cores = cpu_forward_timer(args.height, args.width, args.views)
    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    total = sum(times)
    value = args.steps / total
    sample = (f"{args.steps} full forwards of the workload (1 depth map each) after {args.warmup} warm-up, {cores} torch threads "
              f"(best of calibration {step.calibration}, host has {os.cpu_count()} cores)")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"1ref+{args.views - 1}src {args.width}x{args.height} 3-stage cascade (64/32,16/16,8 hyp), batch 1, host CPU",
                   "weights": "reference checkpoint (fixture)" if os.path.exists(WEIGHTS) else "random init"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# this repo
# ----------------------------------------------------------------------------------------------


class LaunchCounter:
    """Counts native (libpmb200.so) kernel launches issued through patchmatchnet_b200.ops."""

    NAMES = ("relative_projection", "pack_nhwc", "warp_corr", "aggregate_views", "offset_corr", "init_propagate", "adaptive_eval",
             "warp_corr_score", "warp_corr_view_weights", "offset_corr_weight", "aggregate_views_score",
             "conv2d_nhwc", "upsample2x_add_nhwc", "photometric_confidence")

    def __init__(self):
        from patchmatchnet_b200 import _native

        self.n = 0
        self.by_name = {}
        self._lib = _native.lib()
        self._orig = {}

    def __enter__(self):
        for nm in self.NAMES:
            fn = getattr(self._lib, "pmb200_" + nm)
            self._orig[nm] = fn

            def wrapped(*a, _fn=fn, _nm=nm):
                self.n += 1
                self.by_name[_nm] = self.by_name.get(_nm, 0) + 1
                return _fn(*a)

            setattr(self._lib, "pmb200_" + nm, wrapped)
        return self

    def __exit__(self, *exc):
        for nm, fn in self._orig.items():
            setattr(self._lib, "pmb200_" + nm, fn)


def warp_corr_algorithmic_bytes(V, B, C, G, H, W, D):
    """per source view 4*B*H*W*(2C + D + G*D) bytes; one launch processes V views."""
    return V * 4 * B * H * W * (2 * C + D + G * D)


def warp_corr_minimum_bytes(V, B, C, G, H, W, D):
    """What the fused launch must actually move: ref once, V source maps, depth, V weights, one output."""
    return 4 * B * H * W * (C * (1 + V) + D + V + G * D)


KA_ENTRIES = ("warp_corr", "warp_corr_score", "warp_corr_view_weights")


def _ka_row(n, a, k, times_s, peak_gbs):
    ref, src, depth, G = a[0], a[1], a[3], a[4]
    B, H, W, C = ref.shape
    V, D = src.shape[0], depth.shape[1]
    t = statistics.mean(times_s)
    alg = warp_corr_algorithmic_bytes(V, B, C, G, H, W, D)
    per_view_out = n == "warp_corr" and (len(a) < 6 or a[5] is None) and k.get("view_weights") is None
    keeps_sims = n == "warp_corr_view_weights" and (k.get("keep_sims") or (len(a) > 6 and a[6]))
    out_floats = {"warp_corr": G * D * (V if per_view_out else 1), "warp_corr_score": D,
                  "warp_corr_view_weights": V + (G * D * V if keeps_sims else 0)}[n]
    return {"entry": n, "shape": f"C{C} G{G} D{D} {H}x{W} V{V} B{B}", "us": 1e6 * t, "us_min": 1e6 * min(times_s),
            "algorithmic_bytes": alg, "minimum_bytes": 4 * B * H * W * (C * (1 + V) + D + V + out_floats),
            "achieved_gbs": alg / t / 1e9, "frac": alg / t / 1e9 / peak_gbs}


def time_warp_corr_in_step(net, dev_inputs, peak_gbs, flush, steps=10, warmup=3):
    """Launch durations of the fused warp+correlation kernel measured LIVE inside timed steps: eager forwards
    (no graph, so events can bracket a launch), L2 flushed between steps -- not between kernels, the producer
    kernels of the same step leave their outputs in L2 exactly as in production -- one CUDA-event pair per K-A
    launch on the launching stream.  Covers warp_corr / warp_corr_score / warp_corr_view_weights."""
    from patchmatchnet_b200 import ops

    origs = {n: getattr(ops, n) for n in KA_ENTRIES}
    record = []  # (name, args, kwargs, start_event, end_event) per launch of the current step

    def make_spy(n):
        def spy(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = origs[n](*a, **k)
            e1.record()
            record.append((n, a, k, e0, e1))
            return out
        return spy

    for n in KA_ENTRIES:
        setattr(ops, n, make_spy(n))
    per_call = []
    try:
        with torch.no_grad():
            for it in range(warmup + steps):
                record.clear()
                flush()
                # eager Python enqueues slower than the GPU drains (an eager forward is ~90 launches through Python, a few
                # ms of host time, for < 1.5 ms of device time): park the GPU (~12 ms) so the whole step is queued before
                # it starts, otherwise an event pair would also time the host's enqueue latency
                torch.cuda._sleep(24_000_000)
                torch.manual_seed(0)
                net(*dev_inputs())
                torch.cuda.synchronize()
                if it >= warmup:
                    for i, (n, a, k, e0, e1) in enumerate(record):
                        if len(per_call) <= i:
                            per_call.append((n, a, k, []))
                        per_call[i][3].append(e0.elapsed_time(e1) * 1e-3)
    finally:
        for n in KA_ENTRIES:
            setattr(ops, n, origs[n])
    return [_ka_row(n, a, k, ts, peak_gbs) for (n, a, k, ts) in per_call]


def time_warp_corr_isolated(net, dev_inputs, peak_gbs, flush, iters=10):
    """The same launches re-run one at a time with L2 flushed before EACH launch (cold inputs from HBM)."""
    from patchmatchnet_b200 import ops

    calls = []
    origs = {n: getattr(ops, n) for n in KA_ENTRIES}

    def make_spy(n):
        def spy(*a, **k):
            calls.append((n, a, k))
            return origs[n](*a, **k)
        return spy

    for n in KA_ENTRIES:
        setattr(ops, n, make_spy(n))
    try:
        with torch.no_grad():
            torch.manual_seed(0)
            net(*dev_inputs())
    finally:
        for n in KA_ENTRIES:
            setattr(ops, n, origs[n])
    torch.cuda.synchronize()
    rows = []
    for (n, a, k) in calls:
        fn = origs[n]
        for _ in range(2):
            fn(*a, **k)
        ts = []
        for _ in range(iters):
            flush()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*a, **k)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        rows.append(_ka_row(n, a, k, ts, peak_gbs))
    return rows


def _recorded(stream):
    ev = torch.cuda.Event()
    ev.record(stream)
    return ev


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--views", type=int, default=5, help="1 reference + (views-1) sources")
    ap.add_argument("--batch", type=int, default=1, help="reference views per GPU")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the eager-PyTorch reference timing on the GPU")
    ap.add_argument("--no-tf32", action="store_true", help="run the library convolutions in full fp32 instead of torch's default TF32")
    ap.add_argument("--slots", type=int, default=3, help="independent requests in flight per GPU (each its own stream + CUDA graph)")
    ap.add_argument("--cpu-samples", type=int, default=4)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = True  # as the reference's eval.py:301 does; autotuned during warm-up
    if args.no_tf32:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod

    from patchmatchnet_b200.engine import DepthEngine

    B, N, H, W = args.batch, args.views, args.height, args.width
    net, weights = build_net()
    eng = DepthEngine(net, B, N, H, W, device=str(dev), use_graph=not args.no_graph, n_slots=args.slots)

    host = synthetic.make_inputs(B, N, H, W, seed=rank)
    host_pinned = dict(
        images=[im.pin_memory() for im in host["images"]],
        intrinsics=host["intrinsics"].pin_memory(), extrinsics=host["extrinsics"].pin_memory(),
        depth_min=host["depth_min"].pin_memory(), depth_max=host["depth_max"].pin_memory(),
    )
    d_in = dict(images=[im.to(dev) for im in host["images"]], intrinsics=host["intrinsics"].to(dev),
                extrinsics=host["extrinsics"].to(dev), depth_min=host["depth_min"].to(dev), depth_max=host["depth_max"].to(dev))
    for s in range(eng.n_slots):
        eng.set_device_inputs(s, d_in["images"], d_in["intrinsics"], d_in["extrinsics"], d_in["depth_min"], d_in["depth_max"])
    torch.cuda.synchronize()

    with LaunchCounter() as lc:
        eng.prepare()  # warm-up forwards + graph capture
    forwards_counted = eng.n_slots * (eng._warmup + (1 if eng.use_graph else 0))
    launches_per_step = lc.n // max(1, forwards_counted)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def flush():
        flush_buf.zero_()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- kernel-side throughput: inputs resident in HBM ----------------
    # K steps run as rounds of up to `slots` concurrent forwards (one per slot stream); L2 is flushed before every
    # round, outside the event pair that times the round.
    stream = torch.cuda.current_stream()
    S = eng.n_slots
    rounds = [min(S, args.steps - r) for r in range(0, args.steps, S)]
    try:
        uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:  # noqa: BLE001
        uuid = ""
    sampler = ClockSampler(local, uuid)
    sampler.start()  # before the warm-up: the sampler is already running when the timed region starts
    for _ in range(max(1, (args.warmup + S - 1) // S)):
        flush()
        eng.run_round(S, _recorded(stream), torch.cuda.Event())
    barrier()
    sampler.mark_begin()
    starts = [torch.cuda.Event(enable_timing=True) for _ in rounds]
    ends = [torch.cuda.Event(enable_timing=True) for _ in rounds]
    for i, n_now in enumerate(rounds):
        flush()
        starts[i].record(stream)
        eng.run_round(n_now, starts[i], ends[i])
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    dev_seconds = sum(s.elapsed_time(e) for s, e in zip(starts, ends)) * 1e-3

    # ---------------- end to end: pinned host in, pinned host out, every step ----------------
    reqs = [host_pinned] * (args.steps)
    eng.infer_stream(reqs[: max(args.warmup, eng.n_slots)])
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    eng.copy_stream.wait_event(t0)
    for sl in eng._slots:
        sl["stream"].wait_event(t0)
    h2d, d2h = eng.infer_stream(reqs)  # returns after every slot stream has drained
    t1.record(stream)
    barrier()
    e2e_seconds = t0.elapsed_time(t1) * 1e-3

    if dist is not None:
        t = torch.tensor([dev_seconds, e2e_seconds], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_seconds, e2e_seconds = float(t[0]), float(t[1])

    maps = args.steps * B * world
    value = maps / dev_seconds
    e2e_value = maps / e2e_seconds

    # ---------------- roofline of the dominant kernel + CPU baseline (rank 0, N=1 detail) ----------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    roofline, detail, detail_cold, cpu_baseline, gpu_eager = None, None, None, None, None
    if rank == 0:
        dev_inputs = lambda: ([im.clone() for im in d_in["images"]], d_in["intrinsics"].clone(), d_in["extrinsics"].clone(),
                              d_in["depth_min"], d_in["depth_max"])
        detail = time_warp_corr_in_step(net, dev_inputs, peak_gbs, flush, steps=max(5, args.steps // 2))
        detail_cold = time_warp_corr_isolated(net, dev_inputs, peak_gbs, flush)
        if detail:
            # the dominant hand-written kernel is the fused warp+correlation kernel: average over its launches of one step
            n_l = len(detail)
            alg_mean = sum(r["algorithmic_bytes"] for r in detail) / n_l
            us_mean = sum(r["us"] for r in detail) / n_l
            traffic = None
            tfile = os.path.join(REPO, "profiles", "warp_corr_traffic.json")
            if os.path.exists(tfile):
                try:
                    traffic = json.load(open(tfile)).get("mean_dram_bytes_per_launch")
                except Exception:
                    traffic = None
            achieved = alg_mean / (us_mean * 1e-6) / 1e9
            roofline = {"bound": "hbm", "kernel": "warp_corr3_kernel (fused warp + bilinear gather + group correlation + view aggregation + head), "
                                                  f"{n_l} launches per step",
                        "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic,
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_mean, "us_per_launch": us_mean,
                        "how": "mean over the kernel's launches of one step; each launch timed by a CUDA-event pair inside timed eager steps "
                               "(L2 flushed between steps, GPU parked so the host's enqueue latency is not measured); "
                               "algorithmic bytes = V * 4*B*H*W*(2C + D + G*D) per launch (SURVEY.md 8d)",
                        "best_launch_frac": max(r["frac"] for r in detail), "worst_launch_frac": min(r["frac"] for r in detail)}
        if world == 1 and not args.no_gpu_baseline:
            gpu_eager = gpu_eager_reference(H, W, N, dev)
        if world == 1 and not args.no_cpu_baseline:
            step, cores = cpu_forward_timer(H, W, N)
            step()
            ts = [step() for _ in range(args.cpu_samples)]
            cpu_baseline = {"value": len(ts) / sum(ts), "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{len(ts)} full forwards of the same workload (after 1 warm-up) with the oracle port of the reference, "
                                      f"{cores} torch threads (best of calibration {step.calibration}, host has {os.cpu_count()} cores)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_seconds / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"1ref+{N - 1}src {W}x{H} 3-stage cascade (64/32,16/16,8 hyp), batch {B} per GPU (BASELINE.json configs[1])",
                       "parallelism": f"{world} x (1 process/GPU, reference views sharded by rank, no data-path collective)",
                       "weights": weights, "cuda_graph": eng.use_graph, "requests_in_flight": eng.n_slots,
                       "l2": "flushed before every timed round of <= requests_in_flight concurrent steps (256 MiB write, outside the events)",
                       "library_convs": ("cuDNN, full fp32 (--no-tf32)" if args.no_tf32 else
                                         "cuDNN under torch's default flags (TF32 allowed, as the unmodified reference would run on this GPU) for "
                                         "FeatureNet/Refinement/offset convs; every hand-written kernel is full fp32")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_seconds / args.steps,
                    "how": "DepthEngine.infer_stream: pinned host inputs -> device (copy stream) -> CUDA graph on the request's slot stream -> pinned host outputs; "
                           "requests_in_flight independent requests overlap"},
            "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "native_kernels_per_step": {k: v // max(1, forwards_counted) for k, v in sorted(lc.by_name.items())},
            "clocks": clocks,
            "roofline": roofline, "roofline_detail": detail, "roofline_detail_cold_isolated": detail_cold if rank == 0 else None,
            "cpu_baseline": cpu_baseline,
            "gpu_eager_reference": gpu_eager,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
