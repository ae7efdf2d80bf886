#!/usr/bin/env python
"""Benchmark of the dspeed ProcessingChain hot path on B200 (BASELINE.json metric:
waveforms/s for the HPGe DSP chain; % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nproc-per-node N bench.py --gpus N ...     (one rank per GPU)

Workload (``config.workload``): BASELINE.json configs[1] -- the full LEGEND ICPC HPGe
chain (dspeed_b200/configs/hpge_icpc.yaml, 34 outputs) on 1 M synthetic 8192-sample
uint16 waveforms per GPU (weak scaling: every rank owns its own shard of events; there
is no collective on the hot path).  A step is one pass of the chain over the rank's
1 M-row batch.

``value``  : waveforms/s, inputs resident in HBM (device tensors), outputs left in HBM.
``e2e``    : same metric through the public API (build_dsp's ProcessingChain call) with
             the raw table in pinned HOST memory and the output table on the host: H2D
             and D2H copies are inside the timed region.
``roofline``: the dominant kernel's achieved algorithmic HBM bytes/s against the
             measured copy bandwidth in MEASURED_PEAKS.json.
``cpu_baseline`` / ``--impl reference``: the CPU oracle chain (oracle/, a C restatement of
             the reference's numba processors; the reference is Python and cannot travel
             to the GPU box) on the host cores, bounded sample.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

ROWS_PER_GPU = 1_000_000
WF_LEN = 8192
N_OUT = 34
ALGO_BYTES_PER_WF = WF_LEN * 2 + 2 + 4 * N_OUT  # raw u16 + baseline + 34 float32 outputs
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=int(os.environ.get("DSPB_BENCH_ROWS", ROWS_PER_GPU)))
    ap.add_argument("--block-width", type=int, default=int(os.environ.get("DSPB_BENCH_BLOCK", 0)) or None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE.json configurations (C1, C4, C5)")
    # bounded CPU sample: ~10-15 s of work on a 16-thread host at ~10 k wf/s (the chain is O(rows))
    ap.add_argument("--cpu-rows", type=int, default=int(os.environ.get("DSPB_BENCH_CPU_ROWS", 98304)))
    return ap.parse_args()


def load_config():
    import yaml

    return yaml.safe_load(open(os.path.join(REPO, "dspeed_b200", "configs", "hpge_icpc.yaml")))


def hbm_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------
# CPU baseline: the oracle chain on the host cores
# ----------------------------------------------------------------------------------------
CPU_CHUNK = 8192   # rows per oracle call (the reference also walks its input in buffer_len chunks, build_dsp.py:399-407)


def _cpu_pass(chains, vals, bl, consts, cores, keep=None):
    """one pass of the CPU oracle chain over the sample; `keep` (a dict) collects its output columns"""
    parts = []
    for lo in range(0, len(vals), CPU_CHUNK):
        o = chains.icpc_chain(vals[lo:lo + CPU_CHUNK], bl[lo:lo + CPU_CHUNK], consts=consts, keep_waveforms=False,
                              conv="library", threads=cores)
        if keep is not None:
            parts.append(o)
    if keep is not None and parts:
        for k in parts[0]:
            keep[k] = np.concatenate([np.asarray(p[k]) for p in parts])


def cpu_chain_throughput(n_rows: int, repeats: int = 1, vals=None, bl=None, keep=None):
    """waveforms/s of the CPU oracle ICPC chain on `n_rows` synthetic waveforms using all
    host threads (OpenMP over rows, like LEGEND production parallelises over files)."""
    from dspeed_b200 import synth
    from oracle import chains
    from oracle import oracle as O

    cores = O.set_threads(os.cpu_count() or 1)
    if vals is None:
        d = synth.hpge_waveforms(n_rows, seed=2026, stress=True)
        vals, bl = d["values"].numpy(), d["baseline"].numpy()
    consts = chains.icpc_constants()
    chains.icpc_chain(vals[:64], bl[:64], consts=consts, keep_waveforms=False, conv="library", threads=cores)  # warm-up
    best = None
    for _ in range(max(1, repeats)):
        t = time.perf_counter()
        _cpu_pass(chains, vals, bl, consts, cores, keep=keep)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return n_rows / best, cores, best


def genuine_reference():
    """The unmodified reference (`dspeed` + its lgdo / lh5 / pint dependencies), when this box can import it: from
    baseline/_ref (the offline pip install of /root/reference) or the environment.  Returns (module, None) or
    (None, why not).  The reference is pure Python + numba; in the build container its dependencies have no wheels,
    so this normally reports why and the arm times the C restatement instead (kind "port")."""
    ref = os.path.join(REPO, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.append(ref)
    try:
        import dspeed  # noqa: F401
        import lgdo  # noqa: F401
        from dspeed import build_dsp  # noqa: F401
        return dspeed, None
    except Exception as e:   # ImportError, or a broken partial install
        return None, f"{type(e).__name__}: {e}"


def run_genuine_reference(dspeed_mod, vals, bl, steps):
    """times dspeed.build_dsp (numba gufuncs, stock code path) on an in-memory lgdo table of the sample"""
    import lgdo

    from dspeed import build_dsp

    n = len(vals)
    wf = lgdo.WaveformTable(size=n, t0=0, t0_units="ns", dt=16, dt_units="ns", values=vals)
    tb = lgdo.Table(col_dict={"waveform": wf, "baseline": lgdo.Array(bl)}, size=n)
    cfg = load_config()
    db = {"pz": {"tau": 27460.5}}
    build_dsp(raw_in=tb, dsp_config=cfg, database=db, n_entries=min(n, 256))   # numba warm-up
    t = time.perf_counter()
    for _ in range(steps):
        build_dsp(raw_in=tb, dsp_config=cfg, database=db)
    return (time.perf_counter() - t) / steps


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    n = args.cpu_rows
    from dspeed_b200 import synth
    from oracle import chains
    from oracle import oracle as O

    cores = O.set_threads(os.cpu_count() or 1)
    d = synth.hpge_waveforms(n, seed=2026, stress=True)
    vals, bl = d["values"].numpy(), d["baseline"].numpy()
    ref_mod, why_not = genuine_reference()
    if ref_mod is not None:
        try:
            dt = run_genuine_reference(ref_mod, vals, bl, steps)
            v = n / dt
            sample = (f"{n} synthetic 8192-sample waveforms per step through the unmodified dspeed.build_dsp "
                      f"(numba gufuncs, in-memory lgdo table, block_width 16), 1 process")
            line = {
                "impl": "reference", "metric": "waveforms/s for HPGe DSP chain", "value": v, "unit": "waveforms/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "full LEGEND ICPC HPGe chain (34 outputs), 8192-sample uint16 waveforms",
                           "rows_per_step": n},
                "cpu_baseline": {"value": v, "unit": "waveforms/s", "cores": 1, "kind": "reference", "sample": sample},
                "e2e": {"value": v, "unit": "waveforms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            }
            print(json.dumps(line), flush=True)
            return
        except Exception as e:
            why_not = f"dspeed imported but build_dsp failed: {type(e).__name__}: {e}"
    consts = chains.icpc_constants()
    for _ in range(min(warm, 1)):
        chains.icpc_chain(vals[: min(n, 256)], bl[: min(n, 256)], consts=consts, keep_waveforms=False,
                          conv="library", threads=cores)
    t = time.perf_counter()
    for _ in range(steps):
        _cpu_pass(chains, vals, bl, consts, cores)
    dt = (time.perf_counter() - t) / steps
    v = n / dt
    sample = (f"{n} synthetic 8192-sample waveforms per step (bounded sample of the 1M-row workload; "
              f"the reference is O(N), one row at a time), CPU oracle chain (C restatement of the numba "
              f"processors; convolutions through the reference's own numpy.convolve / scipy fftconvolve calls), "
              f"{cores} threads")
    line = {
        "impl": "reference", "metric": "waveforms/s for HPGe DSP chain", "value": v, "unit": "waveforms/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "full LEGEND ICPC HPGe chain (34 outputs), 8192-sample uint16 waveforms",
                   "rows_per_step": n},
        "cpu_baseline": {"value": v, "unit": "waveforms/s", "cores": cores, "kind": "port", "sample": sample,
                         "genuine_reference_unavailable": why_not},
        "e2e": {"value": v, "unit": "waveforms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# clocks during the timed region
# ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                 "-i", str(self.gpu)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t0=None, t1=None):
        """median SM clock / throttle reasons of the samples taken between wall-clock times t0 and t1 (the timed
        region); nvidia-smi needs a few hundred ms to start, so the sampler is started before the warm-up steps and
        the samples are selected by their timestamps (all samples under load if the region was too short for one)"""
        import datetime

        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for ln in open(self.path):
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(p[1]), float(p[2]),
                                 [name for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                                             "sw_power_cap"), p[5:9]) if val.lower().startswith("active")]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        sel = [r for r in rows if t0 is not None and t1 is not None and t0 <= r[0] <= t1]
        window = "timed region"
        if not sel:      # region shorter than the sampling period: the warm-up steps ran the same kernel
            sel, window = rows, "warm-up + timed region"
        if sel:
            out.update(sm_mhz=float(np.median([r[1] for r in sel])), sm_max_mhz=float(max(r[2] for r in sel)),
                       reasons=sorted({x for r in sel for x in r[3]}), samples=len(sel), window=window)
        return out


# ----------------------------------------------------------------------------------------
# host side of the end-to-end leg: where a rank runs and what the box can deliver
# ----------------------------------------------------------------------------------------
def bind_rank_to_its_gpu(local_rank: int, world: int) -> dict:
    """Pin this rank to a private slice of the host cores that are local to its GPU (NVML CPU affinity = the cores of
    the GPU's NUMA node) BEFORE any pinned buffer is allocated: first touch then places the staging pages on that node,
    and the ranks' copy / launch threads do not migrate across each other.  Returns what was done (reported)."""
    info = {"cpus": None, "numa_cpus_of_gpu": None}
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        local = [c for c in range(n_cpu) if (words[c // 64] >> (c % 64)) & 1]
        allowed = sorted(set(local) & set(os.sched_getaffinity(0))) or sorted(os.sched_getaffinity(0))
        info["numa_cpus_of_gpu"] = f"{allowed[0]}-{allowed[-1]} ({len(allowed)})" if allowed else None
        # ranks whose GPUs share these cores split them evenly (contiguous slices)
        sharers = []
        for g in range(world):
            try:
                w = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(g), (n_cpu + 63) // 64)
                if list(w) == list(words):
                    sharers.append(g)
            except Exception:
                pass
        sharers = sharers or [local_rank]
        k, m = sharers.index(local_rank) if local_rank in sharers else 0, len(sharers)
        per = max(1, len(allowed) // m)
        mine = allowed[k * per:(k + 1) * per] or allowed
        os.sched_setaffinity(0, mine)
        info["cpus"] = f"{mine[0]}-{mine[-1]} ({len(mine)})"
    except Exception as e:      # no NVML / not permitted: run unbound, and say so
        info["error"] = f"{type(e).__name__}: {e}"
    return info


def h2d_ceiling(torch, dist, dev, world, nbytes=1 << 30, seconds=1.0):
    """What the box delivers when every rank does nothing but copy pinned host memory to its GPU at the same time (plain
    cudaMemcpyAsync of `nbytes` blocks, two in flight): GB/s of the slowest rank.  The end-to-end rate is reported as a
    fraction of this, so a value that stops scaling with the number of ranks shows up as the HOST's limit."""
    src = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    for s_ in src:
        s_.fill_(1)
    dst = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        for k in range(2):
            with torch.cuda.stream(streams[k]):
                dst[k].copy_(src[k], non_blocking=True)
        n += 2
        for st in streams:
            st.synchronize()
    dt = time.perf_counter() - t0
    rate = n * nbytes / dt / 1e9
    if world > 1:
        t = torch.tensor([rate], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        rate = float(t.item())
    del src, dst
    return rate


# ----------------------------------------------------------------------------------------
# the B200 arm
# ----------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from dspeed_b200 import synth, tables
    from dspeed_b200.processing_chain import build_processing_chain

    binding_info = bind_rank_to_its_gpu(local_rank, world)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.rows
    cfg = load_config()
    steps, warm = max(1, args.steps), max(3, args.warmup)

    # ---- synthetic shard of this rank, generated on the device --------------------------
    data = synth.hpge_waveforms(n, seed=1000 + rank, device=dev, stress=True)
    vals_d, bl_d = data["values"], data["baseline"]

    def table(values, baseline, t0, dt, size=None):
        size = n if size is None else size
        wf = tables.WaveformTable(size=size, t0=tables.Array(t0, attrs={"units": "ns"}),
                                  dt=tables.Array(dt, attrs={"units": "ns"}), values=values)
        return tables.Table({"waveform": wf, "baseline": tables.Array(baseline)}, size=size)

    tb_dev = table(vals_d, bl_d, data["t0"], data["dt"])
    chain, _, tb_out_host = build_processing_chain(cfg, tb_dev, block_width=args.block_width, device=dev)
    out_names = list(tb_out_host.keys())
    tb_out_dev = tables.Table({k: tables.Array(torch.empty(n, dtype=torch.float32, device=dev),
                                               attrs=dict(tb_out_host[k].attrs)) for k in out_names}, size=n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ---------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(warm):
        chain(tb_dev, tb_out_dev)
    chain.enable_event_timing(True)
    launches0 = chain.stats["launches"]
    barrier()
    wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        chain(tb_dev, tb_out_dev)
    e1.record()
    barrier()
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    t_dev = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    launches = chain.stats["launches"] - launches0
    timing = chain.get_timing()  # resolves the per-processor CUDA events
    chain.enable_event_timing(False)
    if rank == 0 and os.environ.get("DSPB_BENCH_TIMING"):
        tot_t = sum(timing.values()) or 1.0
        for name, t in sorted(timing.items(), key=lambda kv: -kv[1]):
            print(f"  {t / steps * 1e3:9.3f} ms/step {100 * t / tot_t:5.1f}%  {name}", file=sys.stderr)
    value = world * n * steps / t_dev

    # ---- the one collective of the path: gather of the output tables (outside the hot path) -------
    gather_ms = None
    if world > 1:
        from dspeed_b200 import parallel

        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        del_me = parallel.gather_table(tb_out_dev, world * n, dst=0)    # untimed: NCCL sets up its point-to-point channels
        del del_me
        barrier()
        g0.record()
        full = parallel.gather_table(tb_out_dev, world * n, dst=0)
        g1.record()
        barrier()
        gather_ms = max_over_ranks(g0.elapsed_time(g1))
        if rank == 0:
            assert len(full) == world * n
        del full

    # ---- dominant kernel and its roofline ------------------------------------------------
    peak, peak_src = hbm_peak()
    fused = chain._fused is not None and chain._fused.can_run(chain)
    if fused:
        k_time = chain._fused.device_time / max(1, chain._fused.device_calls)   # seconds per launch
        rows_per_launch = chain._fused.rows_per_launch
        dom_name = chain._fused.kernel_name
        share = chain._fused.device_time / t_dev
        algo = ALGO_BYTES_PER_WF * rows_per_launch
    else:
        tot = sum(timing.values()) or 1.0
        dom_name, dom_t = max(timing.items(), key=lambda kv: kv[1])
        pm = next(p for p in chain._proc_managers if str(p) == dom_name)
        k_time = dom_t / max(1, getattr(pm, "device_calls", 1))
        rows_per_launch = min(chain._block_width, n)
        share = dom_t / tot
        algo = ALGO_BYTES_PER_WF * rows_per_launch
    achieved = algo / k_time / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": _traffic_from_profile(dom_name, rows_per_launch),
        "traffic_source": "profiles/dominant_kernel_traffic.json (ncu --set full capture of this kernel, bytes per row x rows)",
        "kernel": dom_name, "kernel_share_of_step": share,
        "kernel_ms_per_launch": k_time * 1e3, "rows_per_launch": rows_per_launch,
        "algorithmic_bytes_per_waveform": ALGO_BYTES_PER_WF, "peak_source": peak_src,
        "chain_frac_of_hbm_roofline": (value / world) * ALGO_BYTES_PER_WF / 1e9 / peak,
    }

    # ---- end to end from pinned host memory ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # the whole batch in pinned host memory (16.4 GB per rank); if the host cannot pin that much (8 ranks on one
        # node), the end-to-end leg runs on the largest power-of-two fraction that fits -- it is PCIe bound, so the
        # rate does not depend on the batch size -- and says so in `rows_per_gpu`
        m = min(n, int(os.environ.get("DSPB_BENCH_E2E_ROWS", n)))
        while True:
            try:
                vals_h = torch.empty((m, WF_LEN), dtype=torch.uint16, pin_memory=True)
                break
            except RuntimeError:
                if m <= 65536:
                    raise
                m //= 2
        if world > 1:
            mt = torch.tensor([m], dtype=torch.int64, device=dev)
            dist.all_reduce(mt, op=dist.ReduceOp.MIN)
            m = int(mt.item())
            vals_h = vals_h[:m]
        vals_h.copy_(vals_d[:m])
        bl_h = torch.empty((m,), dtype=torch.uint16, pin_memory=True)
        bl_h.copy_(bl_d[:m])
        tb_host = table(vals_h.numpy(), bl_h.numpy(), data["t0"][:m].cpu().numpy(), data["dt"][:m].cpu().numpy(), size=m)
        h0, d0 = chain.stats["h2d_bytes"], chain.stats["d2h_bytes"]
        chain(tb_host, tb_out_host)  # warm-up (and first-touch of the pinned output columns)
        h1, d1 = chain.stats["h2d_bytes"], chain.stats["d2h_bytes"]
        barrier()
        t = time.perf_counter()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(steps):
            chain(tb_host, tb_out_host)
        f1.record()
        barrier()
        wall = time.perf_counter() - t
        t_e2e = max_over_ranks(max(f0.elapsed_time(f1) * 1e-3, wall))
        ceiling = h2d_ceiling(torch, dist, dev, world)
        h2d_rate = (h1 - h0) * steps / t_e2e / 1e9        # GB/s per GPU actually moved inside the timed region
        e2e = {"value": world * m * steps / t_e2e, "unit": "waveforms/s", "h2d_bytes_per_step": h1 - h0,
               "d2h_bytes_per_step": d1 - d0, "ms_per_step": t_e2e / steps * 1e3, "rows_per_gpu": m,
               "h2d_gbs_per_gpu": h2d_rate, "h2d_ceiling_gbs_per_gpu": ceiling,
               "frac_of_h2d_ceiling": h2d_rate / ceiling if ceiling else None,
               "h2d_ceiling_how": f"all {world} rank(s) copying 1 GiB pinned blocks to their GPUs simultaneously "
                                  f"(cudaMemcpyAsync, 2 in flight), slowest rank",
               "rank0_host_binding": binding_info}
        checksum = float(np.nansum(np.asarray(tb_out_host["trapEmax"].nda, np.float64)[:m]))
    else:
        checksum = None

    # ---- CPU baseline (rank 0, single-GPU run only) --------------------------------------------
    cpu, parity, other = None, None, None
    if rank == 0 and world == 1:
        mc = min(args.cpu_rows, n)   # the first rows of the very batch the GPU processed
        vals_c, bl_c = vals_d[:mc].cpu().numpy(), bl_d[:mc].cpu().numpy()
        oracle_out = {}
        v, cores, secs = cpu_chain_throughput(mc, vals=vals_c, bl=bl_c, keep=oracle_out)
        cpu = {"value": v, "unit": "waveforms/s", "cores": cores, "kind": "port",
               "sample": f"{mc} of the same synthetic waveforms ({secs:.1f} s of CPU work), CPU oracle "
                         f"chain (C restatement of the reference's numba processors; convolutions through the "
                         f"reference's own numpy.convolve / scipy fftconvolve calls), {cores} threads"}
        # how the reference itself runs: ONE thread (numba gufunc target "cpu"), on a smaller sample of the same rows
        from oracle import chains as _chains
        from oracle import oracle as _O

        m1 = min(4096, mc)
        _O.set_threads(1)
        try:
            t1 = time.perf_counter()
            _chains.icpc_chain(vals_c[:m1], bl_c[:m1], consts=_chains.icpc_constants(), keep_waveforms=False,
                               conv="library", threads=1)
            t1 = time.perf_counter() - t1
        finally:
            _O.set_threads(cores)
        cpu["one_thread"] = {"value": m1 / t1, "unit": "waveforms/s", "cores": 1,
                             "sample": f"the first {m1} of those rows ({t1:.1f} s)"}
        # ---- the benchmarked launch checks itself: the device-resident outputs of the last timed step against the
        # oracle outputs of the same rows (tolerances of tests/parity.py; oracle/parity_check.py) ----------------
        from oracle import chains
        from oracle import parity_check

        got = {k: tb_out_dev[k].nda[:mc].cpu().numpy() for k in out_names}
        parity = parity_check.icpc_report(
            got, oracle_out, waves_of=lambda rows: chains.icpc_chain(vals_c[rows], bl_c[rows], keep_waveforms=True))
        parity["path"] = "device-resident one-launch path, rows [0, %d) of the timed batch" % mc
    if rank == 0 and world == 1 and not args.no_configs:
        del tb_dev, vals_d, bl_d, data
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(REPO, "scripts"))
        import bench_configs

        other = bench_configs.run_all(dev)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": "waveforms/s for HPGe DSP chain", "value": value, "unit": "waveforms/s", "n_gpus": world,
        "steps": steps, "warmup": warm, "ms_per_step": t_dev / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": "full LEGEND ICPC HPGe chain (hpge_icpc.yaml, 34 outputs) on synthetic 8192-sample "
                        "uint16 waveforms",
            "rows_per_gpu": n, "wf_len": WF_LEN, "block_width": chain._block_width,
            "rows_per_launch": {"device_resident": int(rows_per_launch), "host_staged": int(min(chain._block_width, n))},
            "sharding": f"events x{world} (no collective on the hot path)",
            "l2_policy": "inputs (16 GB per step) are far larger than the 126 MB L2",
            "fused_kernel": bool(fused),
            "output_gather_ms": gather_ms,
        },
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "parity": parity, "configs": other, "output_checksum_trapEmax": checksum,
    }
    print(json.dumps(line), flush=True)
    if parity is not None and not parity["ok"]:
        print("bench.py: the benchmarked launch does not match the oracle: " + "; ".join(parity["violations"]), file=sys.stderr)
        sys.exit(3)


def _traffic_from_profile(kernel_name: str, rows_per_launch: int):
    """per-launch DRAM bytes (read + write) of the dominant kernel from the committed `ncu --set full`
    capture (profiles/dominant_kernel_traffic.json), scaled to the rows of one bench launch"""
    p = os.path.join(REPO, "profiles", "dominant_kernel_traffic.json")
    try:
        d = json.load(open(p))
        if d.get("kernel", "") not in kernel_name:
            return None
        return d["dram_bytes_per_row"] * rows_per_launch
    except Exception:
        return None


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
