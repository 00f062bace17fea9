#!/usr/bin/env python
"""bench.py - latent vectors quantised per second on the VQ bottleneck (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg3|cfg2|cfg1]

A "step" is one training-mode forward pass of the bottleneck over one batch of synthetic latents resident in HBM:
codebook prep + bf16 latent copy + tcgen05 distance/argmin shortlist + fp32 rescoring + codeword gather + losses +
straight-through output + per-code statistics (+ the NCCL statistics all-reduce when N > 1) + finalize.
Default workload = BASELINE.json configs[2] ("cfg3": N = 2^24 latents per GPU, K = 8192, D = 256), weak scaling:
every rank quantises its own 2^24-frame batch shard, the codebook is replicated.

One JSON line on stdout (rank 0).  `value` = device-resident throughput; `e2e` = the same metric through the
host-buffer C-ABI entry point (pinned host latents -> H2D -> quantise -> D2H indices inside the timed region);
`roofline` = the tcgen05 shortlist kernel against the measured bf16 tensor peak; `cpu_baseline` = a torch-CPU port of the reference's ops
on the host cores over a bounded sample.  `--impl reference` times that CPU port alone (the reference is a Python
module that cannot travel to the GPU box; oracle/vq_oracle.py is its pinned restatement).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write there too (NCCL prints its version banner to stdout under
    NCCL_DEBUG=VERSION, which the GPU boxes set).  Keep a private handle on the real stdout for the JSON line and point
    file descriptor 1 at stderr for everything else."""
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return out

WORKLOADS = {
    # name: (B per GPU, D, W, K, description)
    "cfg3": (1024, 256, 16384, 8192, "quantizer-only N=2^24/GPU K=8192 D=256 (BASELINE.json configs[2])"),
    "cfg2": (64, 64, 16384, 1024, "quantizer-only N=2^20 K=1024 D=64 (BASELINE.json configs[1])"),
    "cfg1": (2, 64, 11000, 512, "VQ bottleneck at debug-batch shape N=22000 K=512 D=64 (BASELINE.json configs[0])"),
    # index export for the BERT stage: indices only (no quantized output, no statistics) + 512-token windows with masks
    "cfg5": (64, 64, 11000, 512, "index export B=64 clips/GPU x 11000 frames, K=512 D=64, 22 windows of 512 (BASELINE.json configs[4])"),
    # full VQ-VAE training step around the bottleneck (stock cuDNN convolutions either side, see run_vqvae_step)
    "cfg4": (64, 64, 11000, 512, "VQ-VAE training step, 64 clips/GPU of 4 x 44000 samples, K=512 D=64 (BASELINE.json configs[3])"),
}
BETA = 0.25
JSON_OUT = sys.stdout


def load_traffic(workload: str):
    """DRAM bytes per tc_search_kernel launch from the committed ncu --set full capture of this workload (profiles/)."""
    p = os.path.join(ROOT, "profiles", "tc_search_traffic.json")
    if os.path.exists(p):
        d = json.load(open(p)).get(workload)
        if d:
            return d.get("dram_bytes_per_launch")
    return None


def fused_operands(W: int, B: int = 1 << 20, D: int = 256) -> bool:
    """Mirror of tc_can_fuse() in csrc/vqb_tc.cu: does the tensor-core kernel read the fp32 [B, D, W] latents itself?"""
    if D > 448 and (B * ((W + 127) // 128) < 4 or os.environ.get("VQB_TC_MODE") == "1"):
        return False
    return os.environ.get("VQB_TC_FUSE", "1") != "0" and W % 4 == 0 and (W % 128 == 0 or W >= 1024)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        note = None
        if not rows:   # timed region shorter than the 100 ms polling period (tiny workloads): one query right after it
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=20).stdout
                rows = [[c.strip() for c in l.split(",")] for l in out.splitlines()]
                rows = [r for r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
                note = "timed region shorter than the polling period: sampled once right after it"
            except (OSError, subprocess.TimeoutExpired):
                rows = []
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        res = {"sm_mhz": statistics.median(float(r[0]) for r in rows), "sm_max_mhz": float(rows[0][1]),
               "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows), "reasons": reasons}
        if note:
            res["note"] = note
        return res


def cpu_port_rate(K: int, D: int, rows_total: int, chunk: int, repeats: int = 1):
    """The reference's CPU algorithm (oracle/ref_port_torch.py: the reference's own torch ops, chunked) on all host
    cores: vectors/s on a bounded sample."""
    import torch
    from oracle.ref_port_torch import vq_forward_chunked
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(42)
    cb = torch.randn(K, D, generator=g)
    z = torch.randn(1, D, rows_total, generator=g)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        vq_forward_chunked(z, cb, BETA, chunk=chunk)
        best = min(best, time.perf_counter() - t0)
    return rows_total / best, best


def cpu_sample_size(K: int, D: int) -> int:
    # ~10-30 s of CPU work on the box's 16 cores (the port spends three sgemm-sized passes per chunk): 2^20 frames at
    # BASELINE config 3 take ~14 s; small workloads are run whole
    target_flops = 8.0e12
    n = int(target_flops / (2.0 * K * D))
    return max(4096, min(1 << 20, 1 << (n.bit_length() - 1)))


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) with all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, D, W, K, desc = WORKLOADS[args.workload]
    n = min(cpu_sample_size(K, D) // 4, B * W)
    chunk = min(n, 32768)
    times = []
    for i in range(args.warmup + args.steps):
        rate, dt = cpu_port_rate(K, D, n, chunk)
        if i >= args.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    value = n / dt
    cores = os.cpu_count() or 1
    sample = f"{n} of {B * W} frames per step in chunks of {chunk} (torch CPU threads = all {cores} cores)"
    line = {"metric": "latent vectors quantized/sec", "value": value, "unit": "vectors/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": desc, "K": K, "D": D, "frames_per_step": n, "note": "CPU oracle port of the reference quantiser"},
            "cpu_baseline": {"value": value, "unit": "vectors/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "vectors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=JSON_OUT, flush=True)


def run_vqvae_step(args):
    """--workload cfg4: encoder -> 1x1 conv -> fused quantiser -> decoder, stage-1 loss, backward, Adam (vqvae.py:59-66,
    81-86, 168-171) on synthetic Slakh-shaped batches, batch-sharded.  The quantiser exchanges its statistics through the
    library's NCCL communicator; the convolution gradients go through torch DDP like the reference's Lightning DDP.  `value`
    stays the BASELINE metric: latent frames quantised per second (B x 11000 per step and GPU)."""
    import torch
    import torch.distributed as dist
    import vq_b200
    from vq_b200 import _lib
    from vq_b200.distributed import StatsComm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = StatsComm()
    B, D, Wf, K, desc = WORKLOADS["cfg4"]
    T = 4 * Wf
    torch.manual_seed(42)                                        # same initial weights on every rank
    model = vq_b200.VQVAEStep(num_embedding=K, embedding_dim=D, precision=args.precision, stats_comm=comm).to(dev)
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = model.configure_optimizers()
    g = torch.Generator().manual_seed(42 + rank)
    host = [(torch.randn(B, 4, T, generator=g) * 0.1).pin_memory() for _ in range(2)]   # configs/data/default.yaml:5-9 shapes
    lib = _lib.lib()
    l1 = torch.nn.functional.l1_loss

    def step(instruments):
        mixed, target = model.make_batch(instruments)
        opt.zero_grad(set_to_none=True)
        output, emb, com, ppl = net(mixed)
        loss = emb + com
        for i in range(4):
            loss = loss + l1(output[:, i, :], target[:, i, :])
        loss.backward()
        opt.step()
        return loss, ppl

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    resident = [h.to(dev) for h in host]
    for i in range(args.warmup):
        step(resident[i % 2])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and not args.no_sampler:
        sampler.start()
    lib.vqb_debug_kernel_timing(1)
    lib.vqb_debug_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        loss, ppl = step(resident[i % 2])
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = int(lib.vqb_debug_launch_count(0))
    kt, kn = C.c_double(0), C.c_int(0)
    _lib.check("vqb_debug_kernel_time_ms", lib.vqb_debug_kernel_time_ms(C.byref(kt), C.byref(kn)))
    stages = {}
    for sid, sname in enumerate(("search", "prep", "fallback", "tail", "pack_stats")):
        st_ms, st_n = C.c_double(0), C.c_int(0)
        _lib.check("vqb_debug_stage_time_ms", lib.vqb_debug_stage_time_ms(sid, C.byref(st_ms), C.byref(st_n)))
        stages[sname] = st_ms.value / max(1, args.steps)
    lib.vqb_debug_kernel_timing(0)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    N = B * Wf
    value = N * world / (ms_per_step * 1e-3)

    # end to end: the step's waveforms start in pinned host memory, the loss is read back every step
    e2e = None
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        for i in range(args.e2e_steps):
            loss, ppl = step(host[i % 2].to(dev, non_blocking=True))
            loss_host = float(loss.item())
        torch.cuda.synchronize(dev)
        te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": N * world * args.e2e_steps / float(te.item()), "unit": "vectors/s", "steps": args.e2e_steps,
               "h2d_bytes_per_step": int(host[0].numel() * 4), "d2h_bytes_per_step": 4,
               "what": "pinned host waveforms -> H2D -> full training step -> loss.item()"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    roofline = None
    if kn.value > 0:
        k_ms = kt.value / kn.value
        achieved = 2.0 * K * D * N / (k_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "tc_search_kernel", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["bf16_tflops"], "peak_source": peaks["source"] + " bf16 burst (short kernel between cuDNN kernels)",
                    "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_per_step, "launches_timed": kn.value, "traffic": None,
                    "algorithmic_flops_per_launch": 2.0 * K * D * N,
                    "note": "K = 512, D = 64: accumulator read-out bound, see DESIGN.md section 7; the step is dominated by the cuDNN convolutions"}
    cpu = None
    if not args.no_cpu:
        rate, secs = cpu_port_rate(K, D, N, 32768)
        cpu = {"value": rate, "unit": "vectors/s", "cores": os.cpu_count(), "kind": "port", "seconds": secs,
               "sample": f"quantiser part only: {N} frames in chunks of 32768, torch-CPU port of the reference ops, all host cores"}
    line = {"metric": "latent vectors quantized/sec", "value": value, "unit": "vectors/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 shortlist + fp32 rescoring; convolutions: torch default (cuDNN, TF32 allowed)", "data": "synthetic",
            "config": {"workload": desc, "frames_per_gpu": N, "clips_per_gpu": B, "K": K, "D": D, "precision": args.precision,
                       "parallelism": f"dp{world}", "l2": "activations of one step (several GB) exceed the 126 MB L2; no explicit flush",
                       "step": "zero_grad + encoder/1x1 conv/quantiser/decoder forward + stage-1 loss + backward + Adam"},
            "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
            "quantiser_ms_per_step": sum(stages.values()), "stage_ms_per_step": stages,
            "clips_per_s": B * world / (ms_per_step * 1e-3),
            "losses": {"total": float(loss.item()), "perplexity": float(ppl.item())}}
    print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-resid", action="store_true", help="diagnostic: skip the per-code residual statistics")
    ap.add_argument("--no-q", action="store_true", help="diagnostic: index export only (no quantized output)")
    ap.add_argument("--no-sampler", action="store_true", help="diagnostic: do not poll nvidia-smi during the timed region")
    args = ap.parse_args()
    global JSON_OUT
    JSON_OUT = claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "cfg4":
        return run_vqvae_step(args)

    import torch
    import torch.distributed as dist
    import vq_b200
    from vq_b200 import _lib, functional as F
    from vq_b200.distributed import StatsComm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = StatsComm()

    B, D, W, K, desc = WORKLOADS[args.workload]
    N = B * W
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    z = torch.randn(B, D, W, device=dev, generator=g)
    gc = torch.Generator(device=dev).manual_seed(4242)
    codebook = torch.randn(K, D, device=dev, generator=gc)      # replicated: same seed on every rank
    stats = torch.empty(_lib.stats_len(K, D), device=dev)
    lib = _lib.lib()

    export_only = args.workload == "cfg5"

    def step():
        if export_only:     # Quantize.get_encodings_idx + the window preparation of AudioBert.forward (transform.py:15-16, bert.py:50-69)
            idx, _, st = F.vq_forward(z, codebook, precision=args.precision, want_q=False, want_resid=False, stats=stats)
            tokens, mask = F.window_indices(idx, B, window=512, pad_id=0)
            return idx, tokens, F.vq_finalize(st, K, D, BETA)
        idx, q, st = F.vq_forward(z, codebook, precision=args.precision, want_q=not args.no_q, want_resid=not args.no_resid, stats=stats)
        if comm is not None:
            comm.allreduce(st)
        losses = F.vq_finalize(st, K, D, BETA)
        return idx, q, losses

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        out = step()
    del out
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and not args.no_sampler:
        sampler.start()
    lib.vqb_debug_kernel_timing(1)
    lib.vqb_debug_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        idx, q, losses = step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = int(lib.vqb_debug_launch_count(0))
    kt, kn = C.c_double(0), C.c_int(0)
    _lib.check("vqb_debug_kernel_time_ms", lib.vqb_debug_kernel_time_ms(C.byref(kt), C.byref(kn)))
    stages = {}
    for sid, sname in enumerate(("search", "prep", "fallback", "tail", "pack_stats")):   # VQB_STAGE_* in include/vqb.h
        st_ms, st_n = C.c_double(0), C.c_int(0)
        _lib.check("vqb_debug_stage_time_ms", lib.vqb_debug_stage_time_ms(sid, C.byref(st_ms), C.byref(st_n)))
        stages[sname] = st_ms.value / max(1, args.steps)
    lib.vqb_debug_kernel_timing(0)
    clocks = sampler.stop() if rank == 0 else None
    counters = F.debug_counters(dev)
    loss_vals = losses.tolist()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = N * world / (ms_per_step * 1e-3)

    # ---- training step (forward + backward) as a second, explanatory number
    train = None
    if not args.no_train and not export_only:
        Gq = torch.randn(B, D, W, device=dev, generator=g) * 1e-3
        one = torch.ones((), device=dev)

        def train_step():
            idx_, q_, losses_ = step()
            return F.vq_backward(z, codebook, idx_, stats, Gq, one, one, BETA)
        for _ in range(2):
            train_step()
        barrier()
        ev0.record()
        n_train = max(2, args.steps // 2)
        for _ in range(n_train):
            dX, dE = train_step()
        ev1.record()
        barrier()
        tt = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        train = {"value": N * world / (float(tt.item()) / n_train * 1e-3), "unit": "vectors/s", "ms_per_step": float(tt.item()) / n_train,
                 "what": "forward + backward (dX, dE) incl. stats all-reduce"}
        del Gq, dX, dE
    del q, idx
    torch.cuda.empty_cache()

    # ---- end to end through the host-buffer C-ABI entry point (pinned host latents, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        z_host = torch.empty((B, D, W), dtype=torch.float32, pin_memory=True)
        z_host.copy_(z)
        cb_host = codebook.cpu().pin_memory()
        idx_host = torch.empty(N, dtype=torch.int64, pin_memory=True)
        stats_host = torch.empty(_lib.stats_len(K, D), dtype=torch.float32, pin_memory=True)
        del z
        torch.cuda.empty_cache()
        F.vq_forward_host(z_host, cb_host, precision=args.precision, want_resid=True, idx_out=idx_host, stats_out=stats_host)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            F.vq_forward_host(z_host, cb_host, precision=args.precision, want_resid=True, idx_out=idx_host, stats_out=stats_host)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        e2e = {"value": N * world * args.e2e_steps / dt, "unit": "vectors/s", "steps": args.e2e_steps,
               "h2d_bytes_per_step": int(z_host.numel() * 4 + cb_host.numel() * 4),
               "d2h_bytes_per_step": int(idx_host.numel() * 8 + stats_host.numel() * 4),
               "what": "vqb_forward_host: pinned host latents -> chunked H2D overlapped with compute -> D2H indices + stats"}
        _lib.check("vqb_host_release", lib.vqb_host_release())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    flops_per_launch = 2.0 * K * D * N                      # 2 K D flops per latent x N latents per launch
    roofline = None
    if kn.value > 0:
        k_ms = kt.value / kn.value
        achieved = flops_per_launch / (k_ms * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        roofline = {"bound": "tensor", "kernel": "tc_search_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "frac_of_burst_peak": achieved / peaks["bf16_tflops"], "peak_source": peaks["source"] +
                    " bf16 sustained (kernel timed inside a long step)", "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_per_step,
                    "launches_timed": kn.value, "traffic": load_traffic(args.workload),
                    "algorithmic_flops_per_launch": flops_per_launch,
                    # fused operand preparation reads the fp32 latents once (4 D bytes per frame); the unfused path reads a bf16 copy
                    "algorithmic_dram_bytes_per_launch": (4.0 if fused_operands(W, B, D) else 2.0) * D * N}
    # second roofline: the tail kernel (rescoring, gather, straight-through value, statistics) against the HBM copy bandwidth.
    # Algorithmic bytes per frame: read the latent (4 D) + write `quantized` (4 D) + write the int64 index (8).
    roofline_tail = None
    if stages.get("tail", 0.0) > 0.0:
        tail_bytes = ((8.0 if not (args.no_q or export_only) else 4.0) * D + 8.0) * N
        gbs = tail_bytes / (stages["tail"] * 1e-3) / 1e9
        roofline_tail = {"bound": "hbm", "kernel": "tail_tma_kernel", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": gbs / peaks["hbm_gbs"], "kernel_ms": stages["tail"], "kernel_share_of_step": stages["tail"] / ms_per_step,
                         "algorithmic_bytes_per_launch": tail_bytes, "traffic": load_traffic(args.workload + "_tail"),
                         "peak_source": peaks["source"] + " copy bandwidth"}
    cpu = None
    if not args.no_cpu:
        n_cpu = min(cpu_sample_size(K, D), N)
        chunk = min(n_cpu, 32768)
        rate, secs = cpu_port_rate(K, D, n_cpu, chunk)
        cpu = {"value": rate, "unit": "vectors/s", "cores": os.cpu_count(), "kind": "port", "seconds": secs,
               "sample": f"{n_cpu} of {N} frames in chunks of {chunk}, torch-CPU port of the reference ops, all host cores"}
    line = {"metric": "latent vectors quantized/sec", "value": value, "unit": "vectors/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 shortlist + fp32 rescoring" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": desc, "frames_per_gpu": N, "K": K, "D": D, "precision": args.precision, "parallelism": f"dp{world}",
                       "l2": (f"inputs ({N * D * 4 / 1e6:.0f} MB of latents per step) larger than the 126 MB L2; no explicit flush"
                              if N * D * 4 > 126e6 else
                              f"inputs ({N * D * 4 / 1e6:.1f} MB of latents per step) FIT in the 126 MB L2 and are not flushed: an L2-warm "
                              "number (parity-case workload, not the bench configuration)"),
                       "step": ("index export (indices + BERT windows + masks)" if export_only else
                                "training-mode forward (indices + quantized + stats + losses)")},
            "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roofline, "roofline_tail": roofline_tail, "cpu_baseline": cpu, "train_step": train,
            "stage_ms_per_step": stages,
            "shortlist": {"rescored_frames_per_step": counters["rescored"], "fallback_frames_per_step": counters["fallback"],
                          "mean_candidates": counters["shortlisted"] / max(1, N)},
            "losses": {"embedding": loss_vals[0], "commitment": loss_vals[1], "perplexity": loss_vals[2]}}
    print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
