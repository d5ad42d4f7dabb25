#!/usr/bin/env python
"""Headline benchmark: CoarseTransformer fwd+bwd tokens/s at seq 2048 (BASELINE.json configs[2], "C3").

    python bench.py --gpus N --steps K --warmup W             # our CUDA path (one rank per GPU via torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU cores

One step = forward of CoarseTransformer(dim 1024, depth 6, heads 8, 4 hyper-connection streams, flash
path) on a [16, 2048]-token batch (372 semantic + 1674 coarse ids + 2 start tokens), the two cross
entropies of CoarseTransformerWrapper.forward (audiolm_pytorch.py:1826-1854), backward to every
parameter, and (N > 1) the flat-bucket gradient all-reduce.  Synthetic ids, random-init weights.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CFG = dict(num_semantic_tokens=500, codebook_size=1024, num_coarse_quantizers=3, dim=1024, depth=6, heads=8,
           flash_attn=True)
BATCH, N_SEM, N_COARSE = 16, 372, 1674  # 1 + 372 + 1 + 1674 = 2048 positions
SEQ = 1 + N_SEM + 1 + N_COARSE
METRIC = "CoarseTransformer tokens/sec fwd+bwd seq2048"


def synth_ids(batch, seed):
    g = torch.Generator().manual_seed(seed)
    sem = torch.randint(0, CFG["num_semantic_tokens"], (batch, N_SEM), generator=g)
    coarse = torch.randint(0, CFG["codebook_size"], (batch, N_COARSE), generator=g)
    return sem, coarse


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = sorted(int(float(s[0])) for s in self.samples if s[0].replace(".", "").isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [int(float(s[1])) for s in self.samples if s[1].replace(".", "").isdigit()]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(self.samples))


# ------------------------------------------------------------------------------------------------
# reference algorithm on the CPU (oracle port of the reference path; the real package cannot be
# installed offline and /root/reference does not exist on the GPU box)
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(batch):
    from oracle import transformer as ot
    from audiolm_pytorch_b200.audiolm import CoarseTransformer

    torch.manual_seed(1234)
    model = CoarseTransformer(**CFG)
    state = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    del model
    sem, coarse = synth_ids(batch, 0)
    coarse_labels = torch.cat((coarse, torch.full((batch, 1), CFG["codebook_size"])), dim=1)
    sem_labels = sem
    hk = dict(heads=CFG["heads"], depth=CFG["depth"], codebook_size=CFG["codebook_size"],
              num_coarse_quantizers=CFG["num_coarse_quantizers"])

    def step():
        for v in state.values():
            v.grad = None
        (sl, cl), _ = ot.coarse_forward(state, sem, coarse, **hk)
        loss = ot.coarse_wrapper_loss(sl, cl, sem_labels, coarse_labels)
        loss.backward()
        return float(loss.detach())

    return step


def run_cpu(steps, warmup, batch=1):
    # more than ~32 intra-op threads makes torch's CPU kernels slower at these sizes (measured: 128 threads ->
    # 113 s/step vs 11 s/step with 8), so the baseline uses min(cores, 32) threads and says so in `cores`
    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    step = cpu_step_fn(batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * SEQ * steps / dt, dt / steps * 1e3, torch.get_num_threads()


WORKLOAD = ("C3 CoarseTransformer d1024 L6 h8 4-stream hyper-connections, flash path, "
            "batch 16/GPU x seq 2048, fwd + CE + bwd")


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else this process (or NCCL) prints went to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tps, ms, cores = run_cpu(args.steps, args.warmup)
    sample = f"batch 1 x {SEQ} tokens per step, fp32, oracle port of the reference path (torch CPU, {cores} threads)"
    emit({
        "impl": "reference", "metric": METRIC, "value": tps, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + " (CPU arm: bounded sample, see cpu_baseline.sample)", "seq_len": SEQ,
                   "global_batch": BATCH * args.gpus, "parallelism": f"dp{args.gpus}"},
        "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------------
# second half of BASELINE.json's metric: SoundStream frames/s encode (config C1 shapes, batch of clips)
# ------------------------------------------------------------------------------------------------
def codec_encode_bench(dev, clips=32, iters=5):
    """encoder conv stack + 8-stage RVQ on `clips` x 2 s @ 24 kHz (48000 samples -> 150 frames each)."""
    from audiolm_pytorch_b200.soundstream import SoundStream

    torch.manual_seed(7)
    ss = SoundStream(codebook_size=1024, rq_num_quantizers=8, target_sample_hz=24000, use_local_attn=False)
    for rvq in ss.rq.rvqs:
        for layer in rvq.layers:
            layer._codebook.embed.normal_()
            layer._codebook.initted.fill_(True)
    ss = ss.to(dev).eval()
    wave = torch.randn(clips, 48000, device=dev)
    with torch.no_grad():
        for _ in range(2):
            ss(wave, return_encoded=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            ss(wave, return_encoded=True)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        from audiolm_pytorch_b200 import ops
        ops.profile_start()
        ss(wave, return_encoded=True)
        prof = ops.profile_stop()
    frames = clips * 150
    kern = {cls: {"ms_per_call": ms_, "launches": n_, "tflops_fp32": work / (ms_ * 1e-3) / 1e12 if ms_ else 0.0}
            for cls, (ms_, work, n_) in prof.items()}
    conv = kern.get("causal_conv1d", {})
    return {"metric": "SoundStream frames/sec encode", "value": frames / (ms * 1e-3), "unit": "frames/s",
            "ms_per_call": ms, "clips": clips, "samples_per_clip": 48000, "kernels": kern,
            # fp32 FMA peak of the part: SMs x 128 lanes x 2 FLOP x SM clock
            "conv_frac_of_fp32_fma_peak": conv.get("tflops_fp32", 0.0) / (148 * 128 * 2 * 1.965e9 / 1e12),
            "note": "fp32 CUDA-core convs + RVQ search (bit-exact code indices vs the fp32 oracle): FMA-bound, not "
                    "HBM-bound (133.8 MB algorithmic I/O per clip would take 21 us at the measured HBM peak)"}


# ------------------------------------------------------------------------------------------------
# our CUDA path
# ------------------------------------------------------------------------------------------------
def main_ours(args):
    import torch.distributed as dist

    from audiolm_pytorch_b200 import _lib, ops
    from audiolm_pytorch_b200.audiolm import CoarseTransformer
    from audiolm_pytorch_b200.heads import cross_entropy
    from audiolm_pytorch_b200.parallel import FlatGradBucket

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    torch.manual_seed(1234)
    model = CoarseTransformer(**CFG).to(dev).train()
    bucket = FlatGradBucket(model.parameters())
    overlap = world > 1 and os.environ.get("ALM_OVERLAP_ALLREDUCE") is not None
    if overlap:
        # optional: layer i's slice of the flat bucket is all-reduced (NCCL, its own stream) as soon as its gradients
        # are final.  Measured at N=2: 30.50 ms/step vs 30.25 ms with ONE all-reduce after the backward (the NCCL
        # kernels only get SMs between the persistent GEMMs and 7 small collectives cost more than one big one),
        # so the single all-reduce stays the default (profiles/r01_bench_n2_v3_overlap_ab.txt).
        ranges = [bucket.range_of(list(layer.parameters())) for layer in model.transformer.layers]
        model.transformer.grad_ready_hook = lambda i: bucket.reduce_range_async(*ranges[i])
    n_params = bucket.numel

    sem_h, coarse_h = synth_ids(BATCH, rank)
    sem_pin, coarse_pin = sem_h.pin_memory(), coarse_h.pin_memory()
    sem_d, coarse_d = sem_h.to(dev), coarse_h.to(dev)
    eos = torch.full((BATCH, 1), CFG["codebook_size"], device=dev)

    def step(sem, coarse):
        bucket.zero_()
        # a real training step sees new weights every iteration: rebuild the bf16 operand copies inside the
        # timed region (what bf16 autocast does at every Linear) instead of reusing last step's cache
        model.transformer.invalidate_weight_cache()
        model._heads.clear()
        coarse_labels = torch.cat((coarse, eos), dim=1)
        sl, cl = model(semantic_token_ids=sem, coarse_token_ids=coarse)
        ls = cross_entropy(sl, sem)
        lc = cross_entropy(cl, coarse_labels)
        n_s, n_c = sl.shape[1], cl.shape[1]
        loss = (ls * n_s + lc * n_c) / (n_s + n_c)
        loss.backward()
        bucket.finish()  # N > 1: the layers' slices were reduced under the backward; this sends the rest and scales
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        step(sem_d, coarse_d)

    # ---- device-resident timing ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    _lib.reset_launch_count()
    ms_total = timed(lambda: step(sem_d, coarse_d), args.steps)
    launches = _lib.launch_count() / args.steps
    clocks = sampler.stop() if sampler else None

    # ---- end to end: pinned host ids -> device, loss -> host, every step ----
    def e2e_step():
        s = sem_pin.to(dev, non_blocking=True)
        c = coarse_pin.to(dev, non_blocking=True)
        return step(s, c).item()

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)

    # ---- roofline of the dominant kernel class (tcgen05 GEMM), events on the launching stream ----
    barrier()
    ops.profile_start()
    for _ in range(2):
        step(sem_d, coarse_d)
    prof = ops.profile_stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    tokens = world * BATCH * SEQ
    ms_step = ms_total / args.steps
    g_ms, g_flops, g_n = prof.get("gemm_bf16_tcgen05", (0.0, 0.0, 0))
    achieved = g_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    kern = {}
    for cls, (ms_, work, n_) in prof.items():
        rate = work / (ms_ * 1e-3) if ms_ else 0.0
        kern[cls] = {"ms_per_step": ms_ / 2, "launches_per_step": n_ / 2}
        if ops.CLASS_UNIT.get(cls) == "byte":  # HBM-bound classes: algorithmic bytes / time vs the measured copy peak
            kern[cls].update(gbps=rate / 1e9, frac_of_hbm_peak=rate / 1e9 / pk["hbm"] if pk.get("hbm") else None)
        else:
            kern[cls]["tflops"] = rate / 1e12

    line = {
        "metric": METRIC, "value": tokens / (ms_step * 1e-3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD + ((" + grad all-reduce" + (" (overlapped with the backward)" if overlap else "")) if world > 1 else ""),
                   "global_batch": world * BATCH, "seq_len": SEQ, "parallelism": f"dp{world}", "params": n_params,
                   "l2": "working set (~9 GB of saved activations per step) far exceeds the 126 MB L2"},
        "e2e": {"value": tokens / (ms_e2e / args.steps * 1e-3), "unit": "tokens/s",
                "h2d_bytes_per_step": (sem_pin.numel() + coarse_pin.numel()) * 8, "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel (all fwd/dgrad/wgrad launches of a step)",
                     "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / pk["tf_sustained"] if pk["tf_sustained"] else None,
                     # DRAM bytes (read + write) of the 108 GEMM launches of one step from an ncu capture of this
                     # command (profiles/r01_ncu_gemm_dram_bytes_per_launch_one_step.csv): 13.60 GB + 3.52 GB
                     "traffic": 17.12e9, "traffic_unit": "bytes per step over all launches of the class",
                     "peak_source": pk["src"] + ", sustained figure (kernel timed inside a long step)",
                     "step_algorithmic_tflop": 388e6 * tokens / world / 1e12},
        "kernels": kern,
    }
    if world == 1:
        try:
            line["soundstream_encode"] = codec_encode_bench(dev)
        except Exception as e:  # pragma: no cover
            line["soundstream_encode"] = {"error": repr(e)}
    if not args.no_cpu:
        try:
            tps, ms_cpu, cores = run_cpu(steps=1, warmup=1)
            line["cpu_baseline"] = {"value": tps, "unit": "tokens/s", "cores": cores, "kind": "port",
                                    "sample": f"1 warm-up + 1 timed fwd+bwd of batch 1 x {SEQ} tokens, fp32 oracle port"}
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"error": repr(e)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the bounded CPU baseline leg")
    a = ap.parse_args()
    # NCCL / torch may print banners ("NCCL version ...") on fd 1: keep stdout for the JSON line only
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        main_reference(a)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (the hot path has no CPU fallback); "
                             "use --impl reference for the CPU arm")
        main_ours(a)
