#!/usr/bin/env python
"""bench.py — speaker-embeddings/sec of the ResCNN hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # our arm, one B200
    torchrun ... bench.py --gpus N --steps K --warmup W      # N replicas (utterance-sharded, no collective)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port) on host cores

A "step" is one forward of the hot path over one batch of 64 synthetic utterances (64 fbank x 160
frames -> 512-d), BASELINE.json configs[1].  One JSON line is printed by rank 0.
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

METRIC = "speaker-embeddings/sec (64-fbank x 160-frame -> 512-d)"
FLOP_PER_EMB = 2306670592
CONV_TC_FLOP_PER_EMB = 2296381440    # the 11 tensor-core convs: 8 x 3x3 (94,371,840 MAC) + 3 x 5x5 s2 (131,072,000 MAC)
L2_BYTES = 126 * 1024 * 1024


_REAL_STDOUT = None


def emit(line: dict):
    """Print the single JSON result line on the real stdout."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops_burst": d.get("bf16_tflops"), "tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.first = 0

    def mark(self):
        """Call at the start of the timed region: nvidia-smi needs up to a second to start, so the sampler is
        launched before warm-up and only the samples taken after this mark are reported."""
        t0 = time.time()
        while self.proc and not self.lines and time.time() - t0 < 3.0:
            time.sleep(0.01)
        self.first = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines[self.first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def make_model(dtype, device):
    """DeepSpeakerModel(512, 1211) with the reference's init and non-trivial BN statistics."""
    import torch

    from deepspeaker_pytorch_b200 import DeepSpeakerModel

    torch.manual_seed(0)
    m = DeepSpeakerModel(512, 1211, operand_dtype=dtype)
    g = torch.Generator().manual_seed(1)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data.uniform_(0.5, 1.5, generator=g)
            mod.bias.data.normal_(0, 0.1, generator=g)
            mod.running_mean.normal_(0, 0.1, generator=g)
            mod.running_var.uniform_(0.5, 1.5, generator=g)
    return m.to(device).eval()


def pick_cpu_threads(sd, T):
    """The reference runs on 'all the host threads it can use'; on many-core hosts oversubscribing the
    torch CPU kernels is slower than a subset, so the baseline gets the best of a few thread counts."""
    import torch

    from oracle import rescnn_oracle as O

    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    x = O.make_input(16, T, seed=0)
    best, best_rate = cands[-1], 0.0
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            O.forward(sd, x)
            t0 = time.perf_counter()
            O.forward(sd, x)
            rate = 16 / (time.perf_counter() - t0)
            if rate > best_rate:
                best, best_rate = c, rate
    return best


def cpu_forward_timer(sd, B, T, budget_s, threads):
    """Times the oracle's eval forward (restatement of /root/reference/model.py:185-218) on host cores."""
    import torch

    from oracle import rescnn_oracle as O

    torch.set_num_threads(threads)
    x = O.make_input(B, T, seed=0)
    with torch.no_grad():
        O.forward(sd, x)  # warm-up
        t0 = time.perf_counter()
        n = 0
        while True:
            O.forward(sd, x)
            n += 1
            el = time.perf_counter() - t0
            if el > budget_s or n >= 50:
                break
    return B * n / el, n, el


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  /root/reference does not exist
    on the GPU box, so this runs the oracle port (same PyTorch CPU kernels the reference dispatches to)."""
    if rank != 0:
        return
    import torch

    from oracle import rescnn_oracle as O

    sd = {k: v for k, v in make_model("fp16", "cpu").state_dict().items()}
    B, T = args.batch, args.frames
    threads = pick_cpu_threads(sd, T)
    torch.set_num_threads(threads)
    x = O.make_input(B, T, seed=0)
    with torch.no_grad():
        t0 = time.perf_counter()
        O.forward(sd, x)
        t1 = time.perf_counter() - t0
        # bound the whole run to ~2 minutes: shrink the per-step sample if needed
        total = (args.steps + args.warmup) * t1
        b = B if total <= 120 else max(1, int(B * 120 / total))
        xs = x[:b]
        for _ in range(args.warmup):
            O.forward(sd, xs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.forward(sd, xs)
        el = time.perf_counter() - t0
    val = b * args.steps / el
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "emb/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"batch-{B} embedding inference, synthetic 64x{T} fbank, eval mode (BASELINE configs[1])",
                   "batch": B, "frames": T},
        "cpu_baseline": {"value": val, "unit": "emb/s", "cores": threads, "host_cpus": os.cpu_count(),
                         "kind": "port",
                         "sample": f"{args.steps} steps x {b} utterances of the batch-{B} workload (oracle port of "
                                   f"model.py:185-218 on torch CPU fp32 kernels)"},
        "e2e": {"value": val, "unit": "emb/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_allpairs(args, rank):
    """BASELINE configs[3]: 1024-utterance all-pairs distance matrix + top-8 hard-negative select (single GPU,
    launch-latency bound: reported in microseconds).  No reference implementation exists."""
    if rank != 0:
        return
    import torch

    from deepspeaker_pytorch_b200 import allpairs_topk

    dev = torch.device("cuda", 0)
    N, D, k = 1024, 512, 8
    g = torch.Generator(device=dev).manual_seed(3)
    sets = []
    for _ in range(8):
        E = torch.randn(N, D, device=dev, generator=g)
        sets.append(10.0 * E / E.norm(dim=1, keepdim=True))
    labels = (torch.arange(N, device=dev) // 16).long()
    K, W = args.steps, args.warmup
    for i in range(W):
        allpairs_topk(sets[i % 8], labels, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        idx, val = allpairs_topk(sets[i % 8], labels, k)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / K * 1e3
    flop = 3.0 * N * N * D
    line = {"metric": "microseconds per 1024-utterance all-pairs distance + top-8 select", "value": us, "unit": "us",
            "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": us / 1e3, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "1024 x 512 embeddings (norm 10), 64 speakers x 16, k=8, different-speaker candidates "
                                   "(BASELINE configs[3])", "arithmetic": "fp32 direct differences (bit-exact indices vs the oracle)"},
            "gflops": flop / (us * 1e-6) / 1e9,
            "e2e": {"value": us, "unit": "us", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_train(args, rank, world, local_rank):
    """Triplet training step (restating train_triplet.py:215-224 with the drop-in classes): three train-mode
    forwards, TripletMarginLoss, backward, ONE gradient allreduce (N > 1), Adagrad step."""
    import torch
    import torch.distributed as dist

    from deepspeaker_pytorch_b200 import TripletMarginLoss
    from deepspeaker_pytorch_b200.parallel import GradBucket, broadcast_parameters, path_parameters

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = 128 if args.batch == 64 else args.batch
    T, K, W = args.frames, args.steps, args.warmup
    model = make_model(args.dtype, dev).train()
    broadcast_parameters(model)
    bucket = GradBucket(path_parameters(model))
    opt = torch.optim.Adagrad(path_parameters(model), lr=0.1, lr_decay=1e-4, weight_decay=0.0)   # train_triplet.py:70-77,378-382
    crit = TripletMarginLoss(0.1)
    g = torch.Generator(device=dev).manual_seed(rank)
    nset = 12
    xs = [tuple(torch.randn(B, 1, T, 64, device=dev, generator=g) for _ in range(3)) for _ in range(nset)]

    def step(xa, xp, xn):
        out_a, out_p, out_n = model(xa), model(xp), model(xn)
        loss = crit.forward(out_a, out_p, out_n)
        bucket.zero()
        loss.backward()
        bucket.allreduce_mean()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(W):
        step(*xs[i % nset])
    barrier()
    if rank == 0:
        sampler.mark()
    e0.record()
    for i in range(K):
        loss = step(*xs[i % nset])
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    # e2e: host inputs, loss read back
    xh = [tuple(torch.randn(B, 1, T, 64).pin_memory() for _ in range(3)) for _ in range(2)]
    xd = [tuple(torch.empty(B, 1, T, 64, device=dev) for _ in range(3)) for _ in range(2)]
    lh = torch.empty(1).pin_memory()
    barrier()
    e0.record()
    for i in range(K):
        for d, h_ in zip(xd[i % 2], xh[i % 2]):
            d.copy_(h_, non_blocking=True)
        lh.copy_(step(*xd[i % 2]).detach().reshape(1), non_blocking=True)
    e1.record()
    barrier()
    ms2 = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms2], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2 = t.item()
    if rank == 0:
        utt = 3 * B * world * K
        line = {
            "metric": "utterances/sec through the triplet training step (3 forwards + loss + backward + Adagrad)",
            "value": utt / (ms * 1e-3), "unit": "utt/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"triplet training step, batch {B} triplets per GPU (anchor/pos/neg), synthetic 64x{T} "
                                   f"fbank, branch A (train_triplet.py:215-224), Adagrad (BASELINE configs[2]/[4])",
                       "global_batch_triplets": B * world, "parallelism": f"dp{world}: one NCCL allreduce of 46.5 MB per step",
                       "l2": "three fresh 5 MB input batches per step; ~2 GB of saved activations per step exceed L2"},
            "clocks": clocks,
            "e2e": {"value": utt / (ms2 * 1e-3), "unit": "utt/s", "h2d_bytes_per_step": 3 * B * T * 64 * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms2 / K},
            "tflops_whole_step": 3 * B * 6911819776 / (ms / K * 1e-3) / 1e12,
            "last_loss": float(loss.item()),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=160)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=3, help="forwards in flight (compute streams) in the inference pipeline")
    ap.add_argument("--workload", default="infer", choices=["infer", "train", "allpairs"],
                    help="infer: batch-64 embedding inference (BASELINE configs[1], the headline metric); "
                         "train: triplet training step, batch-128 triplets per GPU (configs[2]/[4])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: libraries that write to fd 1 (NCCL prints its version banner there)
    # are routed to stderr until the result is printed
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "train":
        run_train(args, rank, world, local_rank)
        return
    if args.workload == "allpairs":
        run_allpairs(args, rank)
        return

    import torch
    import torch.distributed as dist

    from deepspeaker_pytorch_b200 import _lib as L

    assert torch.cuda.is_available(), "bench.py (our arm) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T, K, W = args.batch, args.frames, args.steps, args.warmup
    model = make_model(args.dtype, dev)
    in_bytes = B * T * 64 * 4
    nbuf = L2_BYTES // in_bytes + 8  # rotating inputs larger than L2
    g = torch.Generator(device=dev).manual_seed(rank)
    xs = [torch.randn(B, 1, T, 64, device=dev, generator=g) for _ in range(nbuf)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from deepspeaker_pytorch_b200 import EmbeddingPipeline

    pipe = EmbeddingPipeline(model, lanes=args.lanes)
    cur = torch.cuda.current_stream(dev)
    # ---- value: inputs resident in HBM; `lanes` forwards in flight through the public pipeline ----------------------
    with torch.no_grad():
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        # one-time setup outside the W warm-up steps: every lane builds its plan (first call) and captures its CUDA
        # graph (second call), so a short --steps run does not time graph instantiation
        for i in range(2 * args.lanes):
            pipe.embed_device(xs[i % nbuf])
        pipe.synchronize()
        for i in range(W):
            pipe.embed_device(xs[i % nbuf])
        pipe.synchronize()
        barrier()
        if rank == 0:
            sampler.mark()
        e0.record(cur)
        t_host = time.perf_counter()
        for i in range(K):
            pipe.embed_device(xs[i % nbuf])
        host_ms_value = (time.perf_counter() - t_host) * 1e3 / K   # host enqueue time per step (not a GPU time)
        for st in pipe.lanes:
            cur.wait_stream(st)
        e1.record(cur)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms = max_over_ranks(e0.elapsed_time(e1))
        value = world * B * K / (ms * 1e-3)

        # ---- e2e: host buffers through the public API, H2D + D2H inside the timed region -----------
        nhost = 8
        xh = [torch.randn(B, 1, T, 64).pin_memory() for _ in range(nhost)]
        oh = [torch.empty(B, 512).pin_memory() for _ in range(nhost)]
        for i in range(W):
            pipe.embed(xh[i % nhost], oh[i % nhost])
        pipe.synchronize()
        barrier()
        e0.record(pipe.h2d)
        t_host = time.perf_counter()
        for i in range(K):
            done = pipe.embed(xh[i % nhost], oh[i % nhost])
        host_ms_e2e = (time.perf_counter() - t_host) * 1e3 / K
        pipe.d2h.wait_event(done)
        e1.record(pipe.d2h)
        pipe.synchronize()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        e2e_value = world * B * K / (ms_e2e * 1e-3)

        # ---- roofline of the dominant kernel: per-launch CUDA-event times inside the forward ---------
        eng = model._engine
        import ctypes

        buf = (ctypes.c_float * 32)()
        n = ctypes.c_int32(0)
        nprof = 20

        def profile(level):
            L.check(eng.lib.dsk_set_profiling(eng.handle, level))
            acc = None
            for i in range(nprof):
                model(xs[i % nbuf])
                L.check(eng.lib.dsk_get_launch_times(eng.handle, buf, 32, ctypes.byref(n)))
                v = [buf[j] for j in range(n.value)]
                acc = v if acc is None else [a + b for a, b in zip(acc, v)]
            L.check(eng.lib.dsk_set_profiling(eng.handle, 0))
            return [a / nprof for a in acc]

        # level 2: events only at the section boundaries conv1 | 11 tensor-core convs | tail, so the conv launches run
        # back to back as in production; level 1: an event after every launch (adds ~5 us of event latency to each)
        sec_ms = profile(2)
        per_launch_ms = profile(1)
    conv_ms = sec_ms[1]
    step_ms_prof = sum(sec_ms)
    peaks = load_peaks()
    achieved = B * CONV_TC_FLOP_PER_EMB / (conv_ms * 1e-3) / 1e12
    peak = peaks["tflops_sustained"] if (ms > 2000 and peaks["tflops_sustained"]) else peaks["tflops_burst"]
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of the same 11 launches at batch 64, from the committed
        # `ncu --set full` capture profiles/r01_final2_conv_ncu_full.md (ncu flushes caches between kernels)
        "traffic": 194709248 if (B == 64 and T == 160) else None,
        "kernel": "conv3x3_halo_kernel: the 11 tensor-core conv launches of a step (8 x 3x3 s1 + 3 x parity-planar 5x5 s2), "
                  "timed back to back between two CUDA events on the forward's stream",
        "flop_per_launch_set": B * CONV_TC_FLOP_PER_EMB, "launch_set_ms": conv_ms,
        "share_of_step": conv_ms / step_ms_prof, "section_ms": {"conv1": sec_ms[0], "tensor_core_convs": sec_ms[1], "tail": sec_ms[2]},
        "per_launch_ms_event_bracketed": [round(x, 5) for x in per_launch_ms],
        "peak_source": peaks["source"] + (" sustained" if peak == peaks["tflops_sustained"] else " burst"),
    }

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        threads = pick_cpu_threads(sd, T)
        v, n_it, el = cpu_forward_timer(sd, B, T, budget_s=12.0, threads=threads)
        cpu_baseline = {"value": v, "unit": "emb/s", "cores": threads, "host_cpus": os.cpu_count(),
                        "kind": "port",
                        "sample": f"{n_it} forwards of the same batch-{B} workload in {el:.1f} s (oracle port of "
                                  f"/root/reference/model.py:185-218, torch CPU fp32)"}

    line = {
        "metric": METRIC, "value": value, "unit": "emb/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"batch-{B} embedding inference, synthetic 64x{T} fbank, eval-mode BN, "
                               f"DeepSpeakerModel(512,1211) random init (BASELINE configs[1])",
                   "batch_per_gpu": B, "frames": T, "forwards_in_flight": args.lanes, "parallelism": f"replicas x{world} (utterance-sharded, no collective)",
                   "operands": f"{args.dtype} tensor-core operands, fp32 accumulate/BN/fc/norm",
                   "l2": f"inputs rotate over {nbuf} buffers = {nbuf * in_bytes >> 20} MiB > 126 MiB L2; "
                         f"activations ({B * 1843200 >> 20} MiB/step) are rewritten every step"},
        "clocks": clocks,
        "host_enqueue_ms_per_step": host_ms_value,
        "e2e": {"value": e2e_value, "unit": "emb/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": B * 512 * 4,
                "ms_per_step": ms_e2e / K, "host_enqueue_ms_per_step": host_ms_e2e, "api": "EmbeddingPipeline.embed(pinned host batch) -> pinned host embeddings: H2D, "
                                                  f"the engine forward ({args.lanes} lanes, one CUDA graph per forward) and D2H on their own streams"},
        "gpu_launches": 15 * K,
        "roofline": roofline,
        "tflops_whole_step": B * FLOP_PER_EMB / (ms / K * 1e-3) / 1e12,
    }
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
