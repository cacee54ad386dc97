#!/usr/bin/env python
"""bench.py -- headline benchmark: audio samples/s denoised (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload at N=1 = BASELINE.json configs[1]: 64 ch x 10 min synthetic 48 kHz, stationary reduce_noise,
n_fft=1024 hop=256 (1.8432 G samples per step).  A step is one full pass of the hot path
(k1_analyze -> k_rowfloor -> k_smooth -> k2_synthesize over all 3072 (chunk, channel) units).

  value      samples/s with input and output resident in HBM (CUDA events, max over ranks)
  e2e        the same metric through the C-ABI call with HOST (pinned) buffers: H2D of the input and
             D2H of the result inside the timed region
  roofline   dominant kernel (k2_synthesize: STFT -> mask apply -> iSTFT): 8 algorithmic bytes per
             output sample / its CUDA-event time, against the measured HBM copy bandwidth
  cpu_baseline  the reference's CPU path (oracle/ref_port.py: same scipy.signal calls, joblib over
             chunks) on this box's host cores, on a bounded sample of the same workload

--impl reference times that CPU path alone (the reference is pure Python on third-party scipy and is
not present on the GPU box; see oracle/ref_port.py).
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 48000
C_PER_GPU = 64
N_SAMPLES = 28_800_000            # 10 min @ 48 kHz
METRIC = "audio samples/sec denoised"
UNIT = "samples/s"
ALGO_BYTES_PER_SAMPLE = 8          # float32 in + float32 out


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes():
    """DRAM bytes per k2 launch from the committed ncu capture, if any (profiles/k2_traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "k2_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = float(rows[0][2])
        out["power_w_max"] = max(float(r[3]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, nm in enumerate(names):
            if any(r[5 + i].strip().lower().startswith("active") for r in rows):
                out["reasons"].append(nm)
        out["samples"] = len(rows)
        return out


def synth_device(torch, C, n, c0, device):
    """SURVEY.md section 8d synthetic signal: white noise floor + a gated tone per channel."""
    g = torch.Generator(device=device).manual_seed(1000 + c0)
    y = torch.empty((C, n), dtype=torch.float32, device=device)
    t = torch.arange(n, device=device, dtype=torch.float64) / SR
    gate = ((t % 2.0) < 0.5).to(torch.float32)
    for c in range(C):
        f = 440.0 * 2.0 ** (((c0 + c) % 24) / 12.0)
        y[c] = 0.05 * torch.randn(n, device=device, generator=g)
        y[c] += 0.25 * gate * torch.sin(2 * torch.pi * torch.remainder(t * f, 1.0)).to(torch.float32)
    return y


def cpu_reference_run(steps, warmup, quiet=False):
    """The reference's CPU path on a bounded sample of config 2 (all host cores via joblib)."""
    import numpy as np
    from oracle import ref_port
    from oracle import spectral_gate_oracle as O
    cores = os.cpu_count() or 1
    n_chunks = max(2, min(32, cores))          # bounded sample: <= 32 of config 2's 48 chunks (one per worker)
    C = 64
    n = n_chunks * 600000
    rng = np.random.default_rng(1000)
    base = rng.standard_normal(n + 64 * 997, dtype=np.float32) * np.float32(0.05)
    t = np.arange(n, dtype=np.float64) / SR
    gate = 0.25 * ((t % 2.0) < 0.5)
    y = np.empty((C, n), dtype=np.float32)
    for c in range(C):                         # same distribution as the device workload; shifted noise per channel
        y[c] = base[c * 997: c * 997 + n]
        y[c] += (gate * np.sin(2 * np.pi * 440.0 * 2 ** ((c % 24) / 12) * t)).astype(np.float32)
    cfg = O.GateConfig(sr=SR, stationary=True, n_fft=1024, hop_length=256)
    jobs = min(cores, n_chunks)
    for _ in range(max(0, min(warmup, 1))):           # one warm-up spawns the loky pool
        ref_port.reduce_noise(y[:, : 2 * 600000], SR, cfg, n_jobs=jobs)
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        ref_port.reduce_noise(y, SR, cfg, n_jobs=jobs)
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return dict(value=C * n / dt, unit=UNIT, cores=jobs, kind="port",
                sample=f"64 ch x {n_chunks} chunks of 600000 samples ({C * n / 1e6:.0f} Msamples) per step, "
                       f"joblib n_jobs={jobs} over chunks, scipy.signal stft/fftconvolve/istft",
                seconds_per_step=dt, host_cores=cores)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--minutes", type=float, default=10.0, help="signal length per channel (default config 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)
    config = {"workload": "64ch x 10min synthetic 48kHz stationary reduce_noise n_fft=1024 hop=256 (configs[1])"
                          + (f", x{n_gpus} GPUs: 64 channels per GPU" if n_gpus > 1 else ""),
              "channels_per_gpu": C_PER_GPU, "samples_per_channel": int(args.minutes * 60 * SR),
              "chunk_size": 600000, "padding": 30000, "l2_policy": "inputs (7.4 GB) larger than L2"}
    if n_gpus > 1:
        config["all_gather_bytes_received_per_rank"] = (n_gpus - 1) * C_PER_GPU * int(args.minutes * 60 * SR) * 4

    if args.impl == "reference":
        if rank != 0:
            return
        steps = min(args.steps, 3)
        r = cpu_reference_run(steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": n_gpus,
                "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": r["seconds_per_step"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from noisereduce_b200.device import DeviceGate

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    n = int(args.minutes * 60 * SR)
    C = C_PER_GPU

    x = synth_device(torch, C, n, rank * C, device)
    # multi-GPU transport for the path's one collective (the all-gather of the final waveform):
    #   peer  results pushed into every peer's symmetric-memory buffer by the copy engines (no collective kernels)
    #   nccl  all-gather kernels per 8-channel group on a side stream, 16 SMs reserved for them
    pg = None
    if world > 1 and os.environ.get("B200GATE_GATHER", "peer") == "peer":
        ok = 1
        try:
            from noisereduce_b200.parallel import PeerGather
            pg = PeerGather(world, rank, (world, C, n), torch.float32, device,
                            splits=int(os.environ.get("B200GATE_PUSH_SPLITS", "1")))
        except Exception as exc:                             # e.g. no P2P mapping in this sandbox
            ok = 0
            print(f"[bench] rank {rank}: peer-memory gather unavailable ({exc!r}); using NCCL", file=sys.stderr)
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            pg = None
    out = torch.empty_like(x) if pg is None else None
    # workspace: bits + mask numerators + cached spectra of all 3072 units (41 GB) in one batch -> one launch per kernel
    dg = DeviceGate(sr=SR, stationary=True, n_fft=1024, hop_length=256,
                    reserve_sms=16 if (world > 1 and pg is None) else 0, workspace_limit_bytes=64e9)
    # noise statistics once (stationary.py:61-81): the reference's sequential channel mean, chained over ranks
    if world == 1:
        dg.noise_stats(x)
    else:
        from noisereduce_b200.parallel import chained_noise_stats
        chained_noise_stats(dg, x, rank, world)
    gathered = torch.empty((world * C, n), dtype=torch.float32, device=device) if (world > 1 and pg is None) else None
    comm_stream = torch.cuda.Stream() if world > 1 else None
    acc_stats = {"k1_ms": 0.0, "smooth_ms": 0.0, "k2_ms": 0.0, "fused_ms": 0.0, "kernel_launches": 0}
    if n_gpus > 1:
        config["collective"] = (
            "all-gather of the final [64*N, 28.8M] float32 waveform: each rank's kernels write into its slice of a "
            "symmetric-memory buffer and the copy engines push every finished 8-channel group into all peers' buffers over "
            "NVLink while the next group is computed (no collective kernels, no reserved SMs); device-side barrier per step"
            if pg is not None else
            "all-gather of the final [64*N, 28.8M] float32 waveform, issued as each 8-channel group finishes so NVLink "
            "traffic overlaps the next group's kernels (NCCL kernels; 16 SMs reserved for them)")
        config["gather_transport"] = "peer-copy-engine" if pg is not None else "nccl"

    def step():
        if world == 1:
            dg.run(x, out)
            s_ = dg.gate.stats()
            for k_ in acc_stats:
                acc_stats[k_] += s_[k_]
        elif pg is not None:
            from noisereduce_b200.parallel import sharded_run_peer_push
            sharded_run_peer_push(dg, x, pg, groups=8)
        else:
            # (measured on 2 and 4 B200s: this beats both one monolithic all-gather after the kernels and
            #  64 per-channel zero-copy gathers -- profiles/r01_scaling_notes.md)
            from noisereduce_b200.parallel import sharded_run_overlapped
            sharded_run_overlapped(dg, x, out, gathered, world, comm_stream, groups=8)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k_ in acc_stats:
        acc_stats[k_] = 0
    for _ in range(args.steps):
        step()
    e1.record()
    if world > 1:                                       # per-kernel times of one group set, scaled to the step
        s_ = dg.gate.stats()
        for k_ in acc_stats:
            acc_stats[k_] = s_[k_] * 8 * args.steps
    k1, sm, k2, launches = acc_stats["k1_ms"], acc_stats["smooth_ms"], acc_stats["k2_ms"], acc_stats["kernel_launches"]
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    if world > 1:
        tmax = torch.tensor([ms], device=device)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    ms_per_step = ms / args.steps
    value = world * C * n / (ms_per_step * 1e-3)
    stats = dg.gate.stats()
    # the gathered waveform every rank holds must be what each owner computed: compare per-rank checksums
    gather_verified = None
    if world > 1:
        full = pg.buf if pg is not None else gathered.view(world, C, n)
        mine = full[rank].sum(dtype=torch.float64).view(1)
        owners = torch.empty(world, dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(owners, mine)
        seen = torch.stack([full[r].sum(dtype=torch.float64) for r in range(world)])
        okf = torch.tensor([1.0 if bool((seen == owners).all().item()) else 0.0], device=device)
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        gather_verified = bool(okf.item() > 0)
    ref_out = out if out is not None else pg.buf[rank]

    # ---- end to end through the C ABI with host buffers ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            hx = torch.empty((C, n), dtype=torch.float32, pin_memory=True)
            hy = torch.empty((C, n), dtype=torch.float32, pin_memory=True)
            hx.copy_(x)
            st = torch.cuda.current_stream().cuda_stream

            def e2e_step():
                dg.gate._check(dg.gate.lib.dll.b200gate_run(dg.gate._h, hx.data_ptr(), hy.data_ptr(), 0, C, n, n, n, 0, st))

            e2e_step()
            barrier()
            ksteps = max(2, min(args.steps, 4))
            t0 = time.perf_counter()
            for _ in range(ksteps):
                e2e_step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / ksteps
            ok = 1.0
        except Exception as exc:                      # e.g. pinned-memory limits with many ranks on one host
            dt, ok, ksteps = 0.0, 0.0, 0
            e2e_err = repr(exc)[:200]
        if world > 1:
            tt = torch.tensor([dt, ok], device=device)
            dist.all_reduce(tt[0:1], op=dist.ReduceOp.MAX)
            dist.all_reduce(tt[1:2], op=dist.ReduceOp.MIN)
            dt, ok = float(tt[0].item()), float(tt[1].item())
        if ok > 0 and dt > 0:
            e2e = {"value": world * C * n / dt, "unit": UNIT, "h2d_bytes_per_step": C * n * 4,
                   "d2h_bytes_per_step": C * n * 4, "ms_per_step": dt * 1e3, "steps": ksteps,
                   "parity_vs_device_path": float((hy.to(device) - ref_out).abs().max().item()),
                   "note": "per-rank host buffers (pinned), slab-pipelined H2D / kernels / D2H; no collective in this leg"}
        else:
            e2e = {"value": None, "unit": UNIT, "error": locals().get("e2e_err", "failed on another rank")}
        hx = hy = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_hbm_peak()
    k2_ms = k2 / args.steps
    achieved = ALGO_BYTES_PER_SAMPLE * C * n / (k2_ms * 1e-3) / 1e9 if k2_ms > 0 else None
    pipe_ach = ALGO_BYTES_PER_SAMPLE * world * C * n / (ms_per_step * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "roofline": {"bound": "hbm", "kernel": "k2_synthesize", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic_bytes(),
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SAMPLE * C * n,
                     "kernel_ms": {"k1_analyze+k_rowfloor": k1 / args.steps, "k_smooth": sm / args.steps, "k2_synthesize": k2_ms},
                     "whole_step": {"achieved": pipe_ach, "frac": pipe_ach / (peak * world)},
                     "traffic_note": "ncu dram bytes of k2 (cached spectra 31.7 GB + mask numerators 8.6 GB + waveform) -- profiles/k2_traffic.json",
                     "note": "instruction-issue / dependency bound, not HBM bound (k2: ~2000 warp-instructions per frame pair, 60 % issue-active)"},
        "e2e": e2e,
        "gpu_launches": launches,
        "gather_verified": gather_verified,
        "clocks": clocks,
        "exactness": {k: stats[k] for k in ("bins_rechecked_fp64", "bins_unresolved", "rowfloor_flags", "rowfloor_ambiguous")},
    }
    if n_gpus == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(1, 1)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["cpu_baseline"]["host_cores"] = r["host_cores"]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
