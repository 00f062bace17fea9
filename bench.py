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


def cpu_reference_rate(K: int, D: int, rows_total: int, chunk: int, repeats: int = 1):
    """The reference's CPU quantiser on all host cores: vectors/s on a bounded sample.  Runs the UNMODIFIED reference module
    (vector_quantizer.py:6-54, loaded from /root/reference, baseline/_ref or oracle/_ref - see oracle/make_ref.py) forward on
    chunks of `chunk` frames under no_grad (un-chunked, BASELINE config 3 would need a 550 GB distance matrix); only if no
    reference source is on the machine does it fall back to the pinned torch port, and says so (`kind`)."""
    import torch
    from oracle import make_ref
    from oracle.ref_port_torch import vq_forward_chunked
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(42)
    cb = torch.randn(K, D, generator=g)
    z = torch.randn(1, D, rows_total, generator=g)
    VQ, where = make_ref.load_reference_class()
    if VQ is not None:
        vq = VQ(num_embedding=K, embedding_dim=D, commitment_cost=BETA)
        with torch.no_grad():
            vq.codebook.weight.copy_(cb)

        def run():
            with torch.no_grad():
                for s0 in range(0, rows_total, chunk):
                    vq(z[:, :, s0:s0 + chunk])
        kind = "reference"
    else:
        def run():
            vq_forward_chunked(z, cb, BETA, chunk=chunk)
        kind, where = "port", "oracle/ref_port_torch.py (no reference source found on this machine)"
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        run()
        best = min(best, time.perf_counter() - t0)
    return rows_total / best, best, kind, where


def cpu_sample_size(K: int, D: int) -> int:
    # ~10-30 s of CPU work on the box's host cores (the reference spends three sgemm-sized passes per chunk): 2^20 frames at
    # BASELINE config 3 take ~14 s; small workloads are run whole
    target_flops = 8.0e12
    n = int(target_flops / (2.0 * K * D))
    return max(4096, min(1 << 20, 1 << (n.bit_length() - 1)))


def run_reference(args):
    """--impl reference: the reference's own CPU quantiser with all host threads, a bounded sample of the workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, D, W, K, desc = WORKLOADS[args.workload]
    n = min(cpu_sample_size(K, D) // 4, B * W)
    chunk = min(n, 32768)
    times = []
    kind = where = None
    for i in range(args.warmup + args.steps):
        rate, dt, kind, where = cpu_reference_rate(K, D, n, chunk)
        if i >= args.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    value = n / dt
    cores = os.cpu_count() or 1
    sample = (f"{n} of {B * W} frames per step in chunks of {chunk} (torch CPU threads = all {cores} cores); "
              f"{'the unmodified reference module from ' + where if kind == 'reference' else where}")
    line = {"metric": "latent vectors quantized/sec", "value": value, "unit": "vectors/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"{desc}; CPU arm: a {n}-frame sample per step (rates are per vector)", "K": K, "D": D,
                       "frames_per_step": n, "note": "forward of the reference VectorQuantizer under no_grad, precision 'highest'"},
            "cpu_baseline": {"value": value, "unit": "vectors/s", "cores": cores, "kind": kind, "sample": sample},
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
    model = vq_b200.VQVAEStep(num_embedding=K, embedding_dim=D, precision=args.precision, stats_comm=comm,
                              stats_sync=args.stats_sync).to(dev)
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
    if comm is not None:
        comm.timings = []
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
    stages = stage_times(lib, args.steps)
    lib.vqb_debug_kernel_timing(0)
    if comm is not None and comm.timings:
        stages["allreduce"] = sum(a.elapsed_time(b) for a, b in comm.timings) / args.steps
        comm.timings = None
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
        rate, secs, kind, where = cpu_reference_rate(K, D, N, 32768)
        cpu = {"value": rate, "unit": "vectors/s", "cores": os.cpu_count(), "kind": kind, "seconds": secs,
               "sample": f"quantiser part only: {N} frames in chunks of 32768, reference VectorQuantizer.forward from {where}, all host cores"}
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


def gpu_numa_node(index: int):
    """NUMA node of the GPU's PCIe function and the CPUs of that node (None when the platform does not say)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(index)
        bdf = f"{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        cpus = open(f"/sys/devices/system/node/node{max(node, 0)}/cpulist").read().strip()
        return node, cpus
    except Exception:
        return None, None


def bind_to_numa(index: int):
    """Run this rank (and first-touch its pinned buffers) on the CPUs of its GPU's NUMA node, when there is more than one."""
    node, cpus = gpu_numa_node(index)
    try:
        n_nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
    except OSError:
        n_nodes = 1
    bound = False
    if node is not None and node >= 0 and n_nodes > 1 and cpus:
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        try:
            os.sched_setaffinity(0, ids & os.sched_getaffinity(0) or ids)
            bound = True
        except OSError:
            pass
    return {"gpu_numa_node": node, "numa_nodes_visible": n_nodes, "bound_to_node_cpus": bound}


def stage_times(lib, steps):
    out = {}
    for sid, sname in enumerate(("search", "prep", "fallback", "tail", "pack_stats")):   # VQB_STAGE_* in include/vqb.h
        st_ms, st_n = C.c_double(0), C.c_int(0)
        from vq_b200 import _lib
        _lib.check("vqb_debug_stage_time_ms", lib.vqb_debug_stage_time_ms(sid, C.byref(st_ms), C.byref(st_n)))
        out[sname] = st_ms.value / max(1, steps)
    return out


def parity_multi(dev, comm, rank, world, K, D, precision):
    """Driver-visible multi-GPU parity (VERDICT r01 Next #1e): every rank quantises its own small shard with the NCCL statistics
    exchange; rank 0 re-quantises the CONCATENATED batch alone and compares losses, perplexity, the codebook gradient and its own
    shard's indices."""
    import torch
    import torch.distributed as dist
    from vq_b200 import functional as F
    Bs, Ws = 2, 2048
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    z = torch.randn(Bs, D, Ws, device=dev, generator=g)
    cb = torch.randn(K, D, device=dev, generator=torch.Generator(device=dev).manual_seed(4242))
    one = torch.ones((), device=dev)
    idx, _, st = F.vq_forward(z, cb, precision=precision, want_q=False, want_resid=True)
    st_g, done = comm.allreduce_async(st)
    torch.cuda.current_stream(dev).wait_event(done)
    losses = F.vq_finalize(st_g, K, D, BETA)
    _, dE = F.vq_backward(z, cb, idx, st_g, None, one, one, BETA, need_dx=False)
    allz = [torch.empty_like(z) for _ in range(world)] if rank == 0 else None
    dist.gather(z, allz, dst=0)
    verdict = None
    if rank == 0:
        zc = torch.cat(allz, 0)
        idx1, _, st1 = F.vq_forward(zc, cb, precision=precision, want_q=False, want_resid=True)
        l1 = F.vq_finalize(st1, K, D, BETA)
        _, dE1 = F.vq_backward(zc, cb, idx1, st1, None, one, one, BETA, need_dx=False)
        rel = ((losses - l1).abs() / l1.abs()).max().item()
        de_err = (dE - dE1).abs().max().item() / max(dE1.abs().max().item(), 1e-30)
        same_idx = bool(torch.equal(idx, idx1[:Bs * Ws]))
        ok = rel <= 1e-5 and de_err <= 1e-4 and same_idx
        verdict = {"status": "ok" if ok else "MISMATCH", "frames": Bs * Ws * world, "max_rel_err_losses_perplexity": rel,
                   "max_err_dE_over_max_dE": de_err, "rank0_indices_equal": same_idx,
                   "tolerances": {"losses_perplexity_rtol": 1e-5, "dE": 1e-4}}
    dist.barrier()
    return verdict


def measure_device(args, dev, comm, rank, world, B, D, W, K, steps, warmup, export_only, with_train, sampler_index=None):
    """Device-resident measurement of one configuration: B batch items of W frames on THIS rank.  Returns a dict (rank-local
    times; the caller reduces over ranks)."""
    import torch
    import torch.distributed as dist
    from vq_b200 import _lib, functional as F
    from vq_b200.distributed import side_stream
    lib = _lib.lib()
    N = B * W
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    z = torch.randn(B, D, W, device=dev, generator=g)
    codebook = torch.randn(K, D, device=dev, generator=torch.Generator(device=dev).manual_seed(4242))   # replicated
    stats = torch.empty(_lib.stats_len(K, D), device=dev)
    overlap = comm is not None and args.stats_sync == "overlap"
    side = side_stream(dev) if overlap else None
    cur = torch.cuda.current_stream(dev)

    def step():
        if export_only:     # Quantize.get_encodings_idx + the window preparation of AudioBert.forward (transform.py:15-16, bert.py:50-69)
            idx, _, st = F.vq_forward(z, codebook, precision=args.precision, want_q=False, want_resid=False, stats=stats)
            tokens, mask = F.window_indices(idx, B, window=512, pad_id=0)
            return idx, tokens, F.vq_finalize(st, K, D, BETA), None
        idx, q, st = F.vq_forward(z, codebook, precision=args.precision, want_q=not args.no_q, want_resid=not args.no_resid, stats=stats)
        if comm is None:
            return idx, q, F.vq_finalize(st, K, D, BETA), None
        if not overlap:                       # round-1 behaviour: in place, on the quantiser's stream
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            comm.allreduce(st)
            t1.record()
            if comm.timings is not None:
                comm.timings.append((t0, t1))
            return idx, q, F.vq_finalize(st, K, D, BETA), None
        # the exchange (and the wait for the slowest rank) runs on a side stream next to the following kernels of this
        # stream; a trainer joins it in backward (quantizer.py, stats_sync="overlap"), this loop joins it at the end
        st_g, done = comm.allreduce_async(st)
        with torch.cuda.stream(side):
            losses = F.vq_finalize(st_g, K, D, BETA)
        return idx, q, losses, done

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(warmup):
        out = step()
    del out
    barrier()
    sampler = ClockSampler(sampler_index) if sampler_index is not None else None
    if sampler is not None:
        sampler.start()
    if comm is not None:
        comm.timings = []
    lib.vqb_debug_kernel_timing(1)
    lib.vqb_debug_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        idx, q, losses, done = step()
    if side is not None:
        cur.wait_stream(side)                 # every exchange and finalize of the timed steps is inside the timed region
    ev1.record()
    barrier()
    res = {"ms": ev0.elapsed_time(ev1), "launches": int(lib.vqb_debug_launch_count(0)), "N": N}
    kt, kn = C.c_double(0), C.c_int(0)
    _lib.check("vqb_debug_kernel_time_ms", lib.vqb_debug_kernel_time_ms(C.byref(kt), C.byref(kn)))
    res["kernel_ms_total"], res["kernel_launches"] = kt.value, kn.value
    res["stages"] = stage_times(lib, steps)
    lib.vqb_debug_kernel_timing(0)
    if comm is not None:
        ar = [a.elapsed_time(b) for a, b in comm.timings]
        comm.timings = None
        res["stages"]["allreduce"] = sum(ar) / max(1, steps)
        res["allreduce_on"] = "side stream, overlapped (not on the step's critical path)" if overlap else "quantiser stream, in place"
    res["clocks"] = sampler.stop() if sampler is not None else None
    res["counters"] = F.debug_counters(dev)
    res["losses"] = losses.tolist()

    # ---- training step (forward + backward) as a second, explanatory number
    res["train_ms"] = None
    if with_train and not export_only:
        Gq = torch.randn(B, D, W, device=dev, generator=g) * 1e-3
        one = torch.ones((), device=dev)

        def train_step():
            idx_, q_, losses_, done_ = step()
            st_ = stats
            if done_ is not None:             # join the exchange where backward needs the global statistics
                cur.wait_event(done_)
            return F.vq_backward(z, codebook, idx_, st_, Gq, one, one, BETA)
        for _ in range(2):
            train_step()
        barrier()
        ev0.record()
        n_train = max(2, steps // 2)
        for _ in range(n_train):
            dX, dE = train_step()
        if side is not None:
            cur.wait_stream(side)
        ev1.record()
        barrier()
        res["train_ms"] = ev0.elapsed_time(ev1) / n_train
        del Gq, dX, dE
    del q, idx
    res["z"], res["codebook"] = z, codebook
    return res


def reduce_ranks(dev, world, value):
    """(max, min) over ranks of a rank-local float."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value, -value], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0].item()), -float(t[1].item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "tf32"])
    ap.add_argument("--stats-sync", default="overlap", choices=["overlap", "forward"],
                    help="N > 1: statistics all-reduce on a side stream (default) or in place on the quantiser's stream")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the additional weak-scaling measurement")
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
    import vq_b200  # noqa: F401
    from vq_b200 import _lib, functional as F
    from vq_b200.distributed import StatsComm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_numa(local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = StatsComm()
    lib = _lib.lib()

    B_total, D, W, K, desc = WORKLOADS[args.workload]
    export_only = args.workload == "cfg5"
    # cfg3 / cfg2 / cfg1 are ONE batch sharded over the ranks (BASELINE.json configs[2]: "N=16M latents ... batch-sharded at
    # 1/2/4/8 B200" = what DDP does to one batch, configs/trainer/default.yaml:9-10) -> strong scaling.  cfg5 is quoted per GPU.
    strong = not export_only
    if strong and B_total % world:
        raise SystemExit(f"{args.workload}: batch of {B_total} items does not split evenly over {world} ranks")
    B = B_total // world if strong else B_total
    N = B * W
    N_job = N * world

    parity = parity_multi(dev, comm, rank, world, K, D, args.precision) if world > 1 else None

    m = measure_device(args, dev, comm, rank, world, B, D, W, K, args.steps, args.warmup, export_only, not args.no_train,
                       sampler_index=local if (rank == 0 and not args.no_sampler) else None)
    ms_max, ms_min = reduce_ranks(dev, world, m["ms"])
    ms_per_step = ms_max / args.steps
    value = N_job / (ms_per_step * 1e-3)
    train = None
    if m["train_ms"] is not None:
        tr_max, _ = reduce_ranks(dev, world, m["train_ms"])
        train = {"value": N_job / (tr_max * 1e-3), "unit": "vectors/s", "ms_per_step": tr_max,
                 "what": "forward + backward (dX, dE); N > 1: the statistics exchange is joined before backward"}
    z, codebook = m.pop("z"), m.pop("codebook")

    # ---- end to end through the host-buffer C-ABI entry point (pinned host latents, H2D + D2H inside the timed region)
    e2e = e2e_index_only = None
    if not args.no_e2e:
        z_host = torch.empty((B, D, W), dtype=torch.float32, pin_memory=True)      # allocated after the NUMA binding above
        z_host.copy_(z)
        cb_host = codebook.cpu().pin_memory()
        idx_host = torch.empty(N, dtype=torch.int64, pin_memory=True)
        stats_host = torch.empty(_lib.stats_len(K, D), dtype=torch.float32, pin_memory=True)
        del z
        torch.cuda.empty_cache()

        def e2e_leg(want_q):
            q_host = torch.empty((B, D, W), dtype=torch.float32, pin_memory=True) if want_q else None
            kw = dict(precision=args.precision, want_resid=True, idx_out=idx_host, stats_out=stats_host, want_q=want_q, q_out=q_host,
                      comm=comm)
            F.vq_forward_host(z_host, cb_host, **kw)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                F.vq_forward_host(z_host, cb_host, **kw)
            torch.cuda.synchronize(dev)
            dt_local = time.perf_counter() - t0
            dt, _ = reduce_ranks(dev, world, dt_local)
            h2d = int(z_host.numel() * 4 + cb_host.numel() * 4)
            d2h = int(idx_host.numel() * 8 + stats_host.numel() * 4 + (z_host.numel() * 4 if want_q else 0))
            rate_max, rate_min = reduce_ranks(dev, world, h2d * args.e2e_steps / dt_local / 1e9)
            return {"value": N_job * args.e2e_steps / dt, "unit": "vectors/s", "steps": args.e2e_steps, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "h2d_gbs_per_rank": {"max": rate_max, "min": rate_min},
                    "h2d_gbs_all_ranks": h2d * world * args.e2e_steps / dt / 1e9,
                    "stats_allreduce_included": world > 1, "numa": numa,
                    "what": ("vqb_forward_host: pinned host latents -> chunked H2D overlapped with compute -> D2H indices + stats"
                             + (" + straight-through output `quantized` (what a call of the reference returns, vector_quantizer.py:54)"
                                if want_q else "") + ("; statistics all-reduced over the ranks before the D2H" if world > 1 else ""))}
        def dma_ceiling(duplex):
            """What the host can feed: the same pinned buffers copied H2D (and, duplex, a same-sized buffer D2H at the same time)
            with NO kernel in between, all ranks at once - the ceiling the end-to-end number is measured against."""
            d_in = torch.empty((B, D, W), dtype=torch.float32, device=dev)
            s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            q_host = torch.empty((B, D, W), dtype=torch.float32, pin_memory=True) if duplex else None
            d_out = torch.empty((B, D, W), dtype=torch.float32, device=dev) if duplex else None
            best = float("inf")
            for _ in range(2):
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                with torch.cuda.stream(s_up):
                    d_in.copy_(z_host, non_blocking=True)
                if duplex:
                    with torch.cuda.stream(s_dn):
                        q_host.copy_(d_out, non_blocking=True)
                torch.cuda.synchronize(dev)
                best = min(best, time.perf_counter() - t0)
            lo, hi = reduce_ranks(dev, world, z_host.numel() * 4 / best / 1e9)[::-1]
            return {"gbs_per_rank_each_direction": {"min": lo, "max": hi}, "duplex": duplex,
                    "vectors_per_s_if_dma_bound": N_job / reduce_ranks(dev, world, best)[0]}

        if not export_only:
            e2e = e2e_leg(True)
            e2e["host_dma_ceiling"] = dma_ceiling(True)
            e2e["frac_of_dma_ceiling"] = e2e["value"] / e2e["host_dma_ceiling"]["vectors_per_s_if_dma_bound"]
        e2e_index_only = e2e_leg(False)
        e2e_index_only["host_dma_ceiling"] = dma_ceiling(False)
        e2e_index_only["frac_of_dma_ceiling"] = e2e_index_only["value"] / e2e_index_only["host_dma_ceiling"]["vectors_per_s_if_dma_bound"]
        if e2e is None:
            e2e = e2e_index_only
        if e2e.get("h2d_gbs_per_rank") and world > 1:
            e2e["limiter"] = ("host side, not the kernels: every rank streams its shard over its own PCIe Gen5 x16 link (~55 GB/s alone, "
                              "~46 GB/s each way full duplex); the per-rank rate falls as ranks are added because all GPUs hang off one "
                              "NUMA node and share its DRAM bandwidth - host_dma_ceiling is the same traffic with no kernel at all")
        _lib.check("vqb_host_release", lib.vqb_host_release())
        del z_host, idx_host
    else:
        del z

    # ---- N > 1: the weak-scaling companion (2^24 frames PER GPU, round 1's headline) as an extra object
    weak = None
    if world > 1 and strong and not args.no_weak:
        torch.cuda.empty_cache()
        mw = measure_device(args, dev, comm, rank, world, B_total, D, W, K, max(3, args.steps // 2), 3, False, False)
        mw.pop("z"), mw.pop("codebook")
        w_max, w_min = reduce_ranks(dev, world, mw["ms"])
        w_steps = max(3, args.steps // 2)
        weak = {"scaling": "weak", "frames_per_gpu": B_total * W, "ms_per_step": w_max / w_steps,
                "value": B_total * W * world / (w_max / w_steps * 1e-3), "unit": "vectors/s",
                "rank_ms_per_step": {"max": w_max / w_steps, "min": w_min / w_steps}, "stage_ms_per_step": mw["stages"], "steps": w_steps}
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            comm.close()
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    stages = m["stages"]
    flops_per_launch = 2.0 * K * D * N                      # 2 K D flops per latent x N latents per launch
    roofline = None
    if m["kernel_launches"] > 0 and args.precision != "fp32":
        k_ms = m["kernel_ms_total"] / m["kernel_launches"]
        achieved = flops_per_launch / (k_ms * 1e-3) / 1e12
        tf32 = args.precision == "tf32"
        peak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        peak_burst = peaks["bf16_tflops"]
        if tf32:                                            # tcgen05 kind::tf32 runs at half the bf16 rate (no tf32 entry in MEASURED_PEAKS.json)
            peak, peak_burst = peak / 2, peak_burst / 2
        roofline = {"bound": "tensor", "kernel": "tc_search_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "frac_of_burst_peak": achieved / peak_burst, "peak_source": peaks["source"] +
                    (" bf16 sustained (kernel timed inside a long step)" + (", halved for tf32" if tf32 else "")), "kernel_ms": k_ms,
                    "kernel_share_of_step": k_ms / ms_per_step, "launches_timed": m["kernel_launches"],
                    "traffic": load_traffic(args.workload) if N == WORKLOADS[args.workload][0] * W else None,
                    "traffic_source": "committed ncu --set full capture of this workload (profiles/tc_search_traffic.json), not measured in this run",
                    "algorithmic_flops_per_launch": flops_per_launch,
                    # fused operand preparation reads the fp32 latents once (4 D bytes per frame); the unfused path reads a bf16 copy
                    "algorithmic_dram_bytes_per_launch": (4.0 if fused_operands(W, B, D) else 2.0) * D * N}
    # second roofline: the tail kernel (rescoring, gather, straight-through value, statistics) against the HBM copy bandwidth.
    # Algorithmic bytes per frame: read the latent (4 D) + write `quantized` (4 D) + write the int64 index (8).
    roofline_tail = None
    if stages.get("tail", 0.0) > 0.0:
        tail_bytes = ((8.0 if not (args.no_q or export_only) else 4.0) * D + 8.0) * N
        gbs = tail_bytes / (stages["tail"] * 1e-3) / 1e9
        roofline_tail = {"bound": "hbm", "kernel": "tail3_kernel / tail2_kernel (stage: memsets + tail + residual fold)", "achieved": gbs, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "kernel_ms": stages["tail"],
                         "kernel_share_of_step": stages["tail"] / ms_per_step, "algorithmic_bytes_per_launch": tail_bytes,
                         "traffic": load_traffic(args.workload + "_tail") if N == WORKLOADS[args.workload][0] * W else None,
                         "traffic_source": "committed ncu --set full capture (profiles/tc_search_traffic.json), not measured in this run",
                         "peak_source": peaks["source"] + " copy bandwidth"}
    cpu = None
    if not args.no_cpu:
        n_cpu = min(cpu_sample_size(K, D), N_job)
        chunk = min(n_cpu, 32768)
        rate, secs, kind, where = cpu_reference_rate(K, D, n_cpu, chunk)
        cpu = {"value": rate, "unit": "vectors/s", "cores": os.cpu_count(), "kind": kind, "seconds": secs,
               "sample": f"{n_cpu} of {N_job} frames in chunks of {chunk}, " +
                         (f"the unmodified reference VectorQuantizer.forward (from {where})" if kind == "reference" else where) +
                         ", all host cores"}
    counters = m["counters"]
    line = {"metric": "latent vectors quantized/sec", "value": value, "unit": "vectors/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None,
            "dtype": {"bf16": "bf16 shortlist + fp32 rescoring", "tf32": "tf32 shortlist + fp32 rescoring", "fp32": "f32"}[args.precision],
            "data": "synthetic",
            "config": {"workload": desc, "frames_total": N_job, "frames_per_gpu": N, "K": K, "D": D, "precision": args.precision,
                       "parallelism": f"dp{world}" + (" (one batch sharded by batch item, codebook replicated)" if strong else ""),
                       "l2": (f"inputs ({N * D * 4 / 1e6:.0f} MB of latents per GPU and step) larger than the 126 MB L2; no explicit flush"
                              if N * D * 4 > 126e6 else
                              f"inputs ({N * D * 4 / 1e6:.1f} MB of latents per step) FIT in the 126 MB L2 and are not flushed: an L2-warm "
                              "number (parity-case workload, not the bench configuration)"),
                       "step": ("index export (indices + BERT windows + masks)" if export_only else
                                "training-mode forward (indices + quantized + stats + losses)" +
                                (f"; statistics all-reduce: {m.get('allreduce_on')}" if world > 1 else ""))},
            "clocks": m["clocks"], "gpu_launches": m["launches"], "e2e": e2e, "e2e_index_only": e2e_index_only, "roofline": roofline,
            "roofline_tail": roofline_tail, "cpu_baseline": cpu, "train_step": train,
            "stage_ms_per_step": stages, "rank_ms_per_step": {"max": ms_max / args.steps, "min": ms_min / args.steps},
            "weak": weak, "parity_multi": parity,
            "shortlist": {"rescored_frames_per_step": counters["rescored"], "fallback_frames_per_step": counters["fallback"],
                          "mean_candidates": counters["shortlisted"] / max(1, N)},
            "losses": {"embedding": m["losses"][0], "commitment": m["losses"][1], "perplexity": m["losses"][2]}}
    print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
