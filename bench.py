#!/usr/bin/env python
"""Headline benchmark: Shopformer pose windows/sec on N B200s (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config A|A1|B|C] [--windows 65536] [--precision fp32|bf16]

A *step* is one pass of the scoring hot path (windows -> ST-GCN tokenizer -> transformer ->
per-window reconstruction-error score) over one batch of synthetic windows per GPU.  At N=1 the
workload is BASELINE.json configs[1]: 65,536 synthetic COCO-17 windows (T=24), config A
(`shopformer/` train.py defaults), deterministic synthetic weights.  N>1 is weak scaling: every rank
scores its own 65,536 windows and the scores are all-gathered over NCCL (the path's only collective).

One JSON line is printed by rank 0:
  value      windows/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        the same metric through the C-ABI host-buffer call (`sf_runner_score`: H2D of the poses
             from pinned memory, kernels, D2H of the scores inside the timed region; at N > 1 also the
             all-gather); `e2e.from_tracks` is the same from packed host TRACKS at every N
             (`sf_runner_score_tracks`: half the bytes per window, windowing on the device)
  roofline   dominant kernel (tokenizer) useful FLOP/s against the measured bf16 tensor peak
  cpu_baseline  the CPU oracle port on the host cores on a bounded sample of the same workload
`--impl reference` times the CPU restatement of the reference path (oracle/, torch CPU ops, all host
threads) -- the reference itself is Python and cannot travel to the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import warnings
from pathlib import Path

REPO = Path(__file__).resolve().parent
PKG = REPO / "computer-vision-shoplifting-detection_b200"
for p in (str(REPO), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "shopformer_pose_windows_per_sec"
UNIT = "windows/s"


def load_peaks():
    path = REPO / "MEASURED_PEAKS.json"
    if path.exists():
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


# DRAM bytes per launch of the dominant kernel, from the committed ncu capture (profiles/r1_bf16_summary.md)
def ncu_traffic(config: str, precision: str, kernel: str):
    """DRAM bytes per launch of a kernel (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture),
    parsed from the committed ncu CSV by profiles/ncu_traffic.py into profiles/r2_traffic.json; None when absent."""
    return load_ncu_traffic("batch").get(f"{config}/{precision}/{kernel}")


def build_model(name: str):
    """Drop-in facade of config `name` with deterministic synthetic weights (CPU, eval)."""
    import importlib
    from shopformer_b200 import configs as CFG
    from shopformer_b200.synthetic import synth_state_dict
    which = "shopformer" if CFG.variant_of(name) == 1 else "shopformer_2"
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("models", "data", "utils")}
    sys.path.insert(0, str(PKG / which))
    try:
        models = importlib.import_module("models")
    finally:
        sys.path.remove(str(PKG / which))
        for k in list(sys.modules):
            if k.split(".")[0] in ("models", "data", "utils"):
                sys.modules.pop(k)
        sys.modules.update(saved)
    args = CFG.ctor_args(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = models.Shopformer(**args) if CFG.variant_of(name) == 1 else models.Shopformer(args)
    model.load_state_dict(synth_state_dict(model.state_dict(), seed=0), strict=True)
    return model.eval()


def oracle_runner(model, name: str):
    """Closure scoring a float32 CPU tensor with the CPU oracle port (fp32, all host threads)."""
    import oracle.scoring_oracle as O
    from shopformer_b200 import configs as CFG
    enc = model.gcae.encoder
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    kw = dict(variant=CFG.variant_of(name), strides=list(enc.strides), nhead=model.transformer.nhead,
              pool_tokens=(enc.num_tokens if enc._needs_pooling else None))

    def run(x: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            return O.score_windows(sd, x, dtype=torch.float32, **kw)["score"]
    return run


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(a):
    """CPU restatement of the reference path, timed on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from shopformer_b200 import configs as CFG
    from shopformer_b200.synthetic import synth_windows
    model = build_model(a.config)
    run = oracle_runner(model, a.config)
    C, T, V = CFG.input_shape(a.config)
    sample = a.ref_windows
    x = torch.from_numpy(synth_windows(sample, T, V, seed=1234)[0])
    # all the host threads the box has: torchrun exports OMP_NUM_THREADS=1 to its workers, which would throttle this arm
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    for _ in range(a.warmup):
        run(x)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        run(x)
    dt = time.perf_counter() - t0
    v = sample * a.steps / dt
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"shopformer config {a.config}: {sample}-window sample of the {a.windows}-window batch, "
                                   f"T={T}, V={V}, CPU oracle port (torch CPU ops), batch {sample}", "windows_per_step": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} windows x {a.steps} steps, fp32, torch {torch.__version__} CPU, {os.cpu_count()} logical cpus"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_streaming(a):
    """Config #5: per-tick latency of scoring one new window on each of `--streams` resident streams."""
    from shopformer_b200 import configs as CFG
    from shopformer_b200.streaming import StreamScorer
    torch.cuda.set_device(0)
    model = build_model(a.config).to("cuda")
    C, T, V = CFG.input_shape(a.config)
    stride = T // 2
    sc = StreamScorer(model._sf_engine(), a.streams, T, stride, precision=a.precision, use_graph=True)
    rs = np.random.RandomState(0)
    new = (rs.uniform(100, 900, (a.streams, stride, 17, 3))).astype(np.float32)
    for _ in range(20):
        sc.tick(new)
    lat = []
    for _ in range(a.ticks):
        t0 = time.perf_counter()
        sc.tick(new)
        lat.append(time.perf_counter() - t0)
    lat = np.asarray(lat) * 1e3
    line = {"metric": "shopformer_streaming_tick_latency_ms", "value": float(np.percentile(lat, 50)), "unit": "ms",
            "higher_is_better": False, "n_gpus": 1, "p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)),
            "mean_ms": float(lat.mean()), "windows_per_sec": a.streams / (float(lat.mean()) * 1e-3), "ticks": a.ticks,
            "dtype": "f32" if a.precision == "fp32" else "/".join(sorted(set(sc.eng.tc_formats(T)))), "data": "synthetic",
            "config": {"workload": f"streaming config #5: {a.streams} streams, 1 new window (T={T}, stride {stride}) per stream per tick, "
                                   f"config {a.config}; latency = pinned H2D of the new frames + CUDA-graph replay "
                                   f"(ring shift, normalise, tokenizer, transformer, score) + D2H of the scores, host wall clock"}}
    print(json.dumps(line), flush=True)


def run_windowing(a):
    """HBM-bound kernels of the path: `sf_window_normalize` on packed synthetic tracks (SURVEY 8d recipe) and the
    stand-alone `sf_normality_score`, CUDA-event timed, against the measured HBM copy bandwidth."""
    from shopformer_b200 import configs as CFG
    from shopformer_b200.engine import DeviceTracks, PackedTracks, window_normalize
    from shopformer_b200.synthetic import synth_tracks
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    peaks = load_peaks()
    C, T, V = CFG.input_shape(a.config)
    stride = T // 2
    # ~1500 frames per track -> ~124 windows per track at stride 12; generated on the device (the default 1 M windows are
    # 12.6 M detections = 2.6 GB of keypoints: launch overheads of the four kernels are amortised as in a real sweep)
    from shopformer_b200.synthetic import synth_tracks_device
    n_win_target = a.windows if a.windows != 65536 else 1_048_576
    n_tracks = max(8, int(n_win_target / ((1512 - T) // stride + 1)))
    kp, fno, off, vid, gt, gto = synth_tracks_device(n_tracks, 1512, dev, seed=1234, channels=3)
    dt = DeviceTracks.__new__(DeviceTracks)
    dt.host, dt.device, dt.kp, dt.frame_no, dt.gt = None, dev, kp, fno, gt
    dt._off, dt._vid, dt._gto = off, vid, gto
    out = window_normalize(dt, T, stride, num_keypoints=V)
    nwin = int(out["n_windows"])

    def timed(fn, steps):
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    sampler = ClockSampler(0)
    sampler.start()
    ms_win = timed(lambda: window_normalize(dt, T, stride, num_keypoints=V, sync=False), a.steps)
    # stand-alone MSE score on tokens / recon of the same number of windows
    model = build_model(a.config).to(dev)
    eng = model._sf_engine()
    S, D = eng.token_shape(T)
    tok = torch.randn(nwin, S, D, device=dev)
    rec = torch.randn(nwin, S, D, device=dev)
    ms_score = timed(lambda: eng.normality_score(tok, rec), a.steps)
    clocks = sampler.stop()
    # algorithmic bytes: every raw frame once (stride * K * 12 B per window) + the (2, T, V) fp32 window written;
    # score: both (S, D) fp32 token tensors read + 4 B written
    b_win = stride * 17 * 12 + 2 * T * V * 4
    b_score = 2 * S * D * 4 + 4
    gb_win = b_win * nwin / (ms_win * 1e-3) / 1e9
    gb_score = b_score * nwin / (ms_score * 1e-3) / 1e9
    traffic = load_ncu_traffic("windowing")
    line = {"metric": "shopformer_windowing_windows_per_sec", "value": nwin / (ms_win * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_win, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"windowing + normalisation of {n_tracks} synthetic packed tracks ({dt.kp.shape[0]} detections, T={T}, stride {stride}, V={V}) "
                                   f"-> {nwin} windows; stand-alone score on {nwin} x ({S},{D}) tokens",
                       "l2": f"inputs larger than L2 ({dt.kp.numel() * 4 / 1e6:.0f} MB of keypoints + {nwin * 2 * T * V * 4 / 1e6:.0f} MB of windows vs 126 MB L2)"},
            "clocks": clocks, "gpu_launches": 4 * a.steps,
            "roofline": {"bound": "hbm", "kernel": "k_flag+k_scan+k_compact+k_gather (sf_window_normalize)", "achieved": gb_win, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": gb_win / peaks["hbm_gbs"], "traffic": traffic.get("sf_window_normalize"),
                         "bytes_per_window": b_win, "peak_source": f"{peaks['source']} HBM copy bandwidth (MEASURED_PEAKS.json)"},
            "score_kernel": {"ms": ms_score, "windows_per_sec": nwin / (ms_score * 1e-3),
                             "roofline": {"bound": "hbm", "kernel": "score_kernel (sf_normality_score)", "achieved": gb_score, "peak": peaks["hbm_gbs"],
                                          "unit": "GB/s", "frac": gb_score / peaks["hbm_gbs"], "traffic": traffic.get("sf_normality_score"),
                                          "bytes_per_window": b_score}}}
    print(json.dumps(line), flush=True)


def load_ncu_traffic(tag: str):
    """DRAM bytes per launch parsed from the committed ncu CSVs (profiles/r2_traffic.json, written by
    profiles/ncu_traffic.py from `ncu --page raw --csv` of the captures); {} when absent."""
    path = REPO / "profiles" / "r2_traffic.json"
    if not path.exists():
        return {}
    try:
        return json.load(open(path)).get(tag, {})
    except Exception:
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="A")
    ap.add_argument("--windows", type=int, default=65536, help="windows per GPU per step")
    ap.add_argument("--precision", default="tc", choices=["fp32", "tc", "bf16"],
                    help="tc (alias bf16): the 16-bit tensor-core kernels, operand format per --tc-format; fp32: CUDA-core kernels")
    ap.add_argument("--tc-format", default="auto", choices=["auto", "fp16", "bf16"],
                    help="16-bit operand format of the tensor-core path (auto = fp16 when the weights fit its range)")
    ap.add_argument("--total-windows", type=int, default=10_000_000, help="--mode tracks: windows of the whole sweep (strong scaling)")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra legs (config C @ 262,144, model-call loop, from-tracks e2e)")
    ap.add_argument("--ref-windows", type=int, default=4096, help="--impl reference / cpu_baseline sample per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="batch", choices=["batch", "streaming", "windowing", "tracks"],
                    help="streaming: BASELINE config #5 (512 streams, one window per stream per tick, p50/p99 latency); "
                         "windowing: the HBM-bound windowing / stand-alone score kernels against the HBM roofline; "
                         "tracks: BASELINE config #4 (--total-windows windows cut from packed tracks sharded by track over the "
                         "ranks, NCCL all-gather of the scores; strong scaling)")
    ap.add_argument("--streams", type=int, default=512)
    ap.add_argument("--ticks", type=int, default=2000)
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.precision == "bf16":
        a.precision = "tc"
    if a.tc_format != "auto":
        os.environ["SHOPFORMER_B200_TC_FORMAT"] = a.tc_format
    if a.mode == "streaming":
        return run_streaming(a)
    if a.mode == "windowing":
        return run_windowing(a)
    if a.mode == "tracks":
        return run_tracks(a)
    if a.impl == "reference":
        return run_reference(a)

    from shopformer_b200 import configs as CFG
    from shopformer_b200.sharding import ShardedScorer
    from shopformer_b200.synthetic import synth_windows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the scoring path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    model = build_model(a.config)
    cpu_run = oracle_runner(model, a.config) if rank == 0 else None
    model = model.to(dev)
    eng = model._sf_engine()
    C, T, V = CFG.input_shape(a.config)
    S, D = eng.token_shape(T)
    fmt_tok, fmt_xf = eng.tc_formats(T)
    dtype = "f32" if a.precision == "fp32" else (fmt_tok if fmt_tok == fmt_xf else f"{fmt_tok}+{fmt_xf}")
    n = a.windows
    xs = synth_windows(n, T, V, seed=1234 + rank)[0]
    x = torch.from_numpy(xs).to(dev)
    sharded = ShardedScorer(eng, n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step():
        sharded.score(x, precision=a.precision)

    # ---- parity guard on rank 0: a bench number from wrong scores is worthless
    if rank == 0:
        sub = np.random.RandomState(0).choice(n, 256, replace=False)
        ref = cpu_run(torch.from_numpy(xs[sub])).numpy()
        got = eng.score_windows(x[torch.from_numpy(sub).to(dev)], precision=a.precision).cpu().numpy()
        err = float(np.max(np.abs(got - ref) / np.abs(ref)))
        tol = 1e-3 if a.precision == "fp32" else 1e-2
        if not err < tol:
            raise SystemExit(f"bench.py: scores differ from the CPU oracle by {err:.3e} (> {tol}); refusing to report")
    else:
        err = None

    # ---- device-resident throughput
    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # per-kernel timing of the two kernels of the path (same stream, same inputs): K back-to-back launches each,
    # CUDA events on the launching stream.  (The step itself is ONE native call that launches both kernels, so no
    # event can be placed between them there; tok_ms + xf_ms is checked against the step time below.)
    tok = eng.tokenize(x, precision=a.precision)
    rec = eng.reconstruct_tokens(tok, precision=a.precision)

    def _loop_ms(fn):
        fn()
        torch.cuda.synchronize(dev)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(a.steps):
            fn()
        k1.record()
        torch.cuda.synchronize(dev)
        return k0.elapsed_time(k1) / a.steps

    tok_ms = _loop_ms(lambda: eng.tokenize(x, precision=a.precision, out=tok))
    xf_ms = _loop_ms(lambda: eng.reconstruct_tokens(tok, precision=a.precision, out=rec))
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n * a.steps / (ms * 1e-3)

    # ---- end to end through the C-ABI host-buffer call: inputs in page-locked host memory, H2D + kernels + D2H
    # of every chunk pipelined on a copy + a compute stream inside sf_runner_score
    chunk = 8192              # profiles/r2_e2e_sweep*.txt: 8,192-window uploads hide best behind the kernels
    xs_pinned = torch.from_numpy(xs).pin_memory()
    xs = xs_pinned.numpy()
    eng.score_host(xs, precision=a.precision, chunk=chunk)      # allocate the runner and touch every page once (the first device
    eng.score_host(xs, precision=a.precision, chunk=chunk)      # read of a freshly pinned buffer pays a one-time mapping cost)
    barrier()
    e2e_steps = max(3, min(a.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_scores = eng.score_host(xs, precision=a.precision, chunk=chunk)
        if world > 1:
            # the path's one collective is part of the end-to-end step: every rank ends up with every rank's scores on the host
            sharded.my_slice().copy_(torch.from_numpy(host_scores), non_blocking=True)
            all_host = sharded.gather().cpu()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(t.item())
    h2d = int(xs.nbytes)
    d2h = int(host_scores.nbytes)

    # ---- extra end-to-end legs (rank 0's view, N = 1 only; the headline `e2e` above stays the pre-cut-window call)
    extras = {}
    if not a.no_extras:
        extras["from_tracks"] = tracks_e2e(a, eng, n, T, dev, world, rank)
    if world == 1 and not a.no_extras:
        extras.update(extra_legs(a, model, eng, xs, T, V, dev))

    # ---- the other precision of BASELINE configs[1] ("fp32 and bf16"), device resident, 3 steps
    other = "fp32" if a.precision != "fp32" else "tc"
    other_line = None
    try:
        eng.score_windows(x, precision=other)
        torch.cuda.synchronize(dev)
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record()
        for _ in range(3):
            eng.score_windows(x, precision=other)
        o1.record()
        torch.cuda.synchronize(dev)
        other_line = {"precision": other, "value_per_gpu": n * 3 / (o0.elapsed_time(o1) * 1e-3), "ms_per_step": o0.elapsed_time(o1) / 3}
    except Exception as exc:  # e.g. a shape the tensor-core path does not cover
        other_line = {"precision": other, "error": str(exc)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # launches of OUR kernels per step: tokenizer + transformer, twice when sf_score_windows runs the pass as two halves on
    # two streams (tensor-core path, >= sm_count * 160 windows, SF_SPLIT_STREAMS != 0; csrc/api.cu)
    split_on = a.precision != "fp32" and os.environ.get("SF_SPLIT_STREAMS", "1") != "0"
    split_min = torch.cuda.get_device_properties(dev).multi_processor_count * 160
    launches_per_step = sum(4 if (split_on and min(131072, n - o) >= split_min) else 2 for o in range(0, n, 131072))
    # ---- roofline of the dominant kernel (tokenizer): useful FLOPs per launch / its duration
    tok_flops = {"A": 7_042_080, "A1": 26_988_384, "B": 7_729_344, "C": 17_057_120}.get(a.config)
    dom = "tokenizer" if tok_ms >= xf_ms else "transformer"
    path_flops = CFG.USEFUL_FLOPS.get(a.config)
    if tok_flops is not None:
        dom_flops = tok_flops if dom == "tokenizer" else path_flops - tok_flops
        achieved = dom_flops * n / ((tok_ms if dom == "tokenizer" else xf_ms) * 1e-3) / 1e12
        peak = peaks["bf16_sustained"]
        roofline = {"bound": "tensor", "kernel": f"{dom}_{a.precision}", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": ncu_traffic(a.config, a.precision, dom) if n == 65536 else None,
                    "traffic_note": ("dram__bytes_read.sum + dram__bytes_write.sum of this kernel per 65,536-window launch, bytes, parsed from the "
                                     "committed ncu --set full CSV (profiles/r2_traffic.json; algorithmic: 213.9 MB poses in + 106.9 MB tokens out)")
                    if n == 65536 and ncu_traffic(a.config, a.precision, dom) is not None else None,
                    "peak_source": f"{peaks['source']} bf16 sustained (MEASURED_PEAKS.json)",
                    "flops_per_window": dom_flops, "kernel_ms": {"tokenizer": tok_ms, "transformer": xf_ms, "sum_vs_step": (tok_ms + xf_ms) / (ms / a.steps)},
                    "path_frac": path_flops * n / ((tok_ms + xf_ms) * 1e-3) / 1e12 / peak,
                    "note": "fp32 path runs on the FFMA pipe, not the tensor pipe; frac is vs the bf16 tensor peak"
                            if a.precision == "fp32" else ""}
    else:
        roofline = None

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        sample = min(a.ref_windows, n)
        try:
            torch.set_num_threads(len(os.sched_getaffinity(0)))
        except (AttributeError, OSError):
            torch.set_num_threads(os.cpu_count() or 1)
        xc = torch.from_numpy(xs[:sample])
        cpu_run(xc)
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 200):
            cpu_run(xc)
            reps += 1
        dt = time.perf_counter() - t0
        cpu = {"value": sample * reps / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{sample} windows x {reps} passes of the CPU oracle (torch {torch.__version__} CPU fp32), "
                         f"{os.cpu_count()} logical cpus"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype, "data": "synthetic",
        "config": {"workload": f"shopformer config {a.config}{' (BASELINE configs[1])' if a.config == 'A' else ''}: {n} synthetic windows per GPU, "
                               f"T={T}, V={V}, S={S}, d={D}, deterministic synthetic weights",
                   "windows_per_gpu_per_step": n, "precision": a.precision,
                   "tc_operand_format": {"tokenizer": fmt_tok, "transformer": fmt_xf} if a.precision != "fp32" else None,
                   "l2": (f"inputs larger than L2 ({xs.nbytes / 1e6:.0f} MB of windows per step vs 126 MB L2)" if xs.nbytes > 126e6 else
                          f"inputs ({xs.nbytes / 1e6:.0f} MB) + tokens smaller than the 126 MB L2: not a headline configuration"),
                   "collective": "NCCL all-gather of fp32 scores" if world > 1 else "none (1 GPU)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "sf_runner_score (C ABI, host buffers): 8,192-window chunks of the page-locked pose buffer are DMA-ed on a copy stream "
                       "while the previous chunk is scored on the compute stream, scores copied back per chunk; pageable sources are staged "
                       "through a 4-slot pinned ring" + ("; followed by the NCCL all-gather of the scores and a D2H of the gathered vector" if world > 1 else ""),
                "steps": e2e_steps, **extras},
        "gpu_launches": launches_per_step * a.steps, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "parity": {"max_rel_err_vs_cpu_oracle": err, "checked_windows": 256, "tolerance": 1e-3 if a.precision == "fp32" else 1e-2},
        "other_precision": other_line,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def tracks_e2e(a, eng, n, T, dev, world, rank):
    """`e2e.from_tracks`, every N: each rank scores its OWN packed tracks (about n windows) from page-locked host memory
    through sf_runner_score_tracks -- x, y only, 12 new frames per stride-12 window = half the PCIe / host-memory bytes of
    pre-cut windows; windowing + normalisation + scoring on the device -- then (N > 1) the ragged NCCL all-gather of the
    scores and a D2H of the gathered vector.  Wall clock, max over ranks; windows = sum over ranks."""
    from shopformer_b200.engine import PackedTracks
    from shopformer_b200.sharding import gather_ragged
    from shopformer_b200.synthetic import synth_tracks
    stride = T // 2
    try:
        n_tracks = max(8, int(n / ((1515 - T) / stride)))
        tr = synth_tracks(n_tracks, seed=4321 + rank)
        pin = lambda arr: torch.from_numpy(np.ascontiguousarray(arr)).pin_memory().numpy()      # every bulk array page-locked
        kp2 = pin(tr["kp"][:, :, :2])                                                            # the confidence channel is dropped at ingest
        host = PackedTracks(kp=kp2, frame_no=pin(tr["frame_no"]), track_offsets=tr["track_offsets"], track_video=tr["track_video"],
                            gt=pin(tr["gt"]), gt_offsets=tr["gt_offsets"])
        kw = dict(add_neck=False, precision=a.precision, chunk=16384)

        def one():
            res = eng.score_tracks_host(host, T, stride, **kw)
            if world > 1:
                allsc, _ = gather_ragged(torch.from_numpy(res["scores"][:int(res["n_windows"])]).to(dev, non_blocking=True))
                allsc.cpu()
            return res
        for _ in range(2):
            res = one()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            res = one()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        nw = int(res["n_windows"])
        tt = torch.tensor([dt, float(nw)], dtype=torch.float64, device=dev)
        if world > 1:
            t_max = tt[:1].clone()
            dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
            n_sum = tt[1:].clone()
            dist.all_reduce(n_sum, op=dist.ReduceOp.SUM)
            dt, nw_all = float(t_max.item()), int(n_sum.item())
        else:
            nw_all = nw
        return {"value": nw_all * reps / dt, "unit": UNIT, "windows_per_step": nw_all,
                "h2d_bytes_per_step_per_gpu": int(kp2.nbytes + tr["frame_no"].nbytes), "d2h_bytes_per_step_per_gpu": 16 * nw,
                "api": "sf_runner_score_tracks: host packed tracks (x, y) -> windowing + normalisation + scoring on the device -> "
                       "host scores, labels and window index; groups of whole tracks uploaded under the previous group's kernels"
                       + ("; then the ragged NCCL all-gather of the scores and a D2H of the gathered vector" if world > 1 else "")}
    except Exception as exc:
        return {"error": str(exc)[:200]}


def extra_legs(a, model, eng, xs, T, V, dev):
    """Extra keys of the N=1 line: (1) the same windows scored FROM PACKED TRACKS on the host (sf_runner_score_tracks:
    x, y only, 12 new frames per stride-12 window = half the PCIe bytes of pre-cut windows), (2) the reference's own loop
    shape -- `model(x)` on host tensors at batch 32 (train.py:evaluate) and 4096 -- so per-call host overhead is visible,
    (3) BASELINE configs[2]: config C at 262,144 windows, device resident."""
    from shopformer_b200 import configs as CFG
    from shopformer_b200.engine import PackedTracks
    from shopformer_b200.synthetic import synth_tracks, synth_windows
    out = {}
    n = xs.shape[0]
    stride = T // 2
    # (2) the reference's loop: model(x)['normality_score'] per batch, host tensors in, host scores out
    try:
        loop = {}
        xt = torch.from_numpy(xs)
        with torch.no_grad():
            for bs, nb in ((32, 256), (4096, 16)):
                tot = bs * nb
                def run():
                    acc = []
                    for i in range(0, tot, bs):
                        acc.append(model(xt[i:i + bs].to(dev))["normality_score"].cpu().numpy())
                    return acc
                run()
                t0 = time.perf_counter()
                run()
                dt = time.perf_counter() - t0
                loop[f"batch_{bs}"] = {"windows_per_sec": tot / dt, "us_per_call": 1e6 * dt / nb}
        out["model_call_loop"] = {"api": "model(poses.to(device))['normality_score'].cpu().numpy() per batch (shopformer/train.py:311-319, "
                                         "evaluate.py:89-99), pageable host tensors", **loop}
    except Exception as exc:
        out["model_call_loop"] = {"error": str(exc)[:200]}
    # (4) BASELINE configs[4]: 512 streams, one new window per stream per tick (CUDA-graph tick), p50 / p99 latency
    try:
        if a.config == "A":
            from shopformer_b200.streaming import StreamScorer
            sc = StreamScorer(eng, 512, T, stride, precision=a.precision, use_graph=True)
            new = np.random.RandomState(0).uniform(100, 900, (512, stride, 17, 3)).astype(np.float32)
            for _ in range(20):
                sc.tick(new)
            lat = []
            for _ in range(1000):
                t0 = time.perf_counter()
                sc.tick(new)
                lat.append(time.perf_counter() - t0)
            lat = np.asarray(lat) * 1e3
            out["streaming_512"] = {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)), "mean_ms": float(lat.mean()),
                                    "windows_per_sec": 512 / (float(lat.mean()) * 1e-3), "ticks": 1000,
                                    "workload": "BASELINE configs[4]: 512 streams x 1 new window per tick; pinned H2D of the new frames + CUDA-graph replay "
                                                "(ring shift, normalise, tokenizer, transformer, score) + D2H of the scores, host wall clock per tick"}
            del sc
    except Exception as exc:
        out["streaming_512"] = {"error": str(exc)[:200]}
    # (5) the incumbent GPU number: eager PyTorch (ATen composition of the same modules, eval + no_grad, TF32 off) on this GPU
    try:
        os.environ["SHOPFORMER_B200_EAGER_BASELINE"] = "1"
        tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        eager = {}
        with torch.no_grad():
            for bs in (32, 4096, 16384):
                xb = torch.from_numpy(xs[:bs]).to(dev)
                for _ in range(2):
                    ref_scores = model(xb)["normality_score"]
                torch.cuda.synchronize(dev)
                reps = 20 if bs == 32 else 5
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(reps):
                    ref_scores = model(xb)["normality_score"]
                g1.record()
                torch.cuda.synchronize(dev)
                eager[f"batch_{bs}"] = {"windows_per_sec": bs * reps / (g0.elapsed_time(g1) * 1e-3), "ms_per_call": g0.elapsed_time(g1) / reps}
            del os.environ["SHOPFORMER_B200_EAGER_BASELINE"]
            native = model(xb)["normality_score"]
        eager["max_rel_diff_vs_native"] = float(((native - ref_scores).abs() / ref_scores.abs()).max())
        out["eager_aten_gpu"] = {"what": "ATen composition of the same nn.Modules (the drop-in's training path in eval mode) on this GPU, device-resident "
                                         "inputs, fp32, TF32 off: stands in for the reference's eager PyTorch on 1xB200 (the reference itself cannot travel to "
                                         "the GPU box); it also computes the GCAE decoder output, as the reference's forward does", **eager}
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    except Exception as exc:
        os.environ.pop("SHOPFORMER_B200_EAGER_BASELINE", None)
        out["eager_aten_gpu"] = {"error": str(exc)[:200]}
    # (3) BASELINE configs[2]
    try:
        if a.config == "A":
            mc = build_model("C").to(dev)
            ec = mc._sf_engine()
            Cc, Tc, Vc = CFG.input_shape("C")
            nC = 262144
            xc = torch.from_numpy(synth_windows(nC, Tc, Vc, seed=99)[0]).to(dev)
            ec.score_windows(xc, precision=a.precision)
            torch.cuda.synchronize(dev)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(3):
                ec.score_windows(xc, precision=a.precision)
            c1.record()
            torch.cuda.synchronize(dev)
            ms = c0.elapsed_time(c1) / 3
            out["config_C_262144"] = {"value": nC / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "tc_operand_format": list(ec.tc_formats(Tc)),
                                      "path_frac": CFG.USEFUL_FLOPS["C"] * nC / (ms * 1e-3) / 1e12 / load_peaks()["bf16_sustained"],
                                      "workload": "BASELINE configs[2]: shopformer_2 default config (V=17 T=24 H=64, 12 heads, 4+4 layers, ff 512), "
                                                  "262,144 windows device resident, 3 steps"}
            del xc, ec, mc
    except Exception as exc:
        out["config_C_262144"] = {"error": str(exc)[:200]}
    # (4) the other canonical configs at 65,536 windows, device resident, 5 steps: B = shopformer_2 paper config, A1 = shopformer/
    # class defaults (hidden 64)
    if a.config == "A":
        for name in ("B", "A1"):
            try:
                mo = build_model(name).to(dev)
                eo = mo._sf_engine()
                _, To, Vo = CFG.input_shape(name)
                xo = torch.from_numpy(synth_windows(65536, To, Vo, seed=98)[0]).to(dev)
                eo.score_windows(xo, precision=a.precision)
                torch.cuda.synchronize(dev)
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record()
                for _ in range(5):
                    eo.score_windows(xo, precision=a.precision)
                c1.record()
                torch.cuda.synchronize(dev)
                ms = c0.elapsed_time(c1) / 5
                out[f"config_{name}_65536"] = {"value": 65536 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "tc_operand_format": list(eo.tc_formats(To)),
                                               "path_frac": CFG.USEFUL_FLOPS[name] * 65536 / (ms * 1e-3) / 1e12 / load_peaks()["bf16_sustained"]}
                del xo, eo, mo
            except Exception as exc:
                out[f"config_{name}_65536"] = {"error": str(exc)[:200]}
    return out


def run_tracks(a):
    """BASELINE configs[3]: a sweep of --total-windows windows cut from packed tracks, sharded by track over the ranks with the
    host prefix sum of per-track window counts (global window index = reference order), every rank windows + normalises +
    scores its shard from device-resident tracks, the scores are all-gathered over NCCL.  Strong scaling."""
    from shopformer_b200 import configs as CFG
    from shopformer_b200.engine import DeviceTracks, PackedTracks
    from shopformer_b200.sharding import gather_ragged, shard_tracks
    from shopformer_b200.synthetic import synth_tracks_device
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    model = build_model(a.config).to(dev)
    eng = model._sf_engine()
    C, T, V = CFG.input_shape(a.config)
    stride = T // 2
    track_len = 1512
    per_track = (track_len - T) // stride + 1
    n_tracks = (a.total_windows + per_track - 1) // per_track
    offsets = np.arange(n_tracks + 1, dtype=np.int64) * track_len
    lo, hi, first = shard_tracks(offsets, T, stride, world)[rank]
    # every rank generates only its own tracks (seeded by the first track index: the sweep is the same at every N)
    kp, fno, off, vid, gt, gto = synth_tracks_device(hi - lo, track_len, dev, seed=1234 + lo, channels=2)
    dt = DeviceTracks.__new__(DeviceTracks)
    dt.host, dt.device, dt.kp, dt.frame_no, dt.gt = None, dev, kp, fno, gt
    dt._off, dt._vid, dt._gto = off, vid, gto

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step():
        r = eng.score_tracks(dt, T, stride, add_neck=False, precision=a.precision)
        allv, counts = gather_ragged(r["scores"])
        return r, allv, counts

    for _ in range(max(a.warmup, 1)):
        r, allv, counts = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    steps = max(1, min(a.steps, 5))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r, allv, counts = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    total = int(sum(counts))
    if rank == 0:
        fmt = eng.tc_formats(T)
        line = {"metric": METRIC, "value": total * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(a.warmup, 1),
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32" if a.precision == "fp32" else "/".join(sorted(set(fmt))), "data": "synthetic",
                "config": {"workload": f"BASELINE configs[3]: sweep of {n_tracks} synthetic packed tracks x {track_len} detections (x, y; "
                                       f"{n_tracks * track_len * V * 8 / 1e9:.1f} GB) -> {total} valid windows (T={T}, stride {stride}) of config {a.config}, "
                                       f"sharded by track with the host prefix sum of per-track window counts, scores all-gathered",
                           "windows_total": total, "windows_per_rank": counts, "precision": a.precision,
                           "l2": "inputs larger than L2 (GBs of keypoints per rank)", "collective": "NCCL all-gather of fp32 scores (ragged: counts first)" if world > 1 else "none (1 GPU)"},
                "gpu_launches": None, "clocks": clocks,
                "note": "device-resident tracks; per step: k_flag/k_scan/k_compact once, then per 131,072-window pass k_gather + tokenizer + transformer"}
        passes = (int(counts[0]) + 131071) // 131072
        line["gpu_launches"] = steps * (3 + 3 * passes)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
