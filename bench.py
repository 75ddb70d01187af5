#!/usr/bin/env python
"""bench.py -- embedding-optimisation images/s of the StableKeypoints Stage-1 hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--tokens 77] [--precision reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one optimizer step of BASELINE cfg2 on each rank: ONE synthetic 512^2 image through
run_and_find_attn twice (original + affine-warped; VAE encode, noised frozen SD1.5 UNet forward with the four
up-block cross-attention captures, fused capture+collect), token selection (Gaussian-KL top-25, furthest-point 10),
sharpening + equivariance losses, backward into the [1,N,768] embedding, gradient all-reduce over ranks, Adam.
Weak scaling: every rank processes its own image each step (the reference's 1 image/GPU, optimize.py:333).

Prints ONE JSON line (rank 0).  `value` = images/s with the images already resident in HBM; `e2e` = the same loop fed
from pinned host memory with the loss read back every step.  `--impl reference` times the CPU oracle port of the same
iteration (oracle/ = restated fp32 UNet + restated reference hook/collect/loss code) on this box's host cores.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "embedding-opt images/sec @512^2 SD1.5 UNet"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tokens", type=int, default=77, help="N learned tokens (north_star 77; notebook 100; CLI default 500)")
    ap.add_argument("--res", type=int, default=128, help="feature_upsample_res R")
    ap.add_argument("--precision", default=os.environ.get("SKP_PRECISION", "fp32"), choices=["fp32", "reference", "tf32"],
                    help="numerics of the few torch ops left in the trunk (fp32 = no TF32 anywhere: the parity-tested setting)")
    ap.add_argument("--early-exit", action="store_true", help="stop the forward after the 4th captured layer (outputs identical)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--trunk", default=os.environ.get("SKP_TRUNK", "tc"), choices=["tc", "torch"],
                    help="tc: trunk convs/linears on the tcgen05 split-bf16 GEMM (fp32-grade accuracy); torch: cuDNN/cuBLAS")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="drive the step from Python instead of one CUDA graph")
    ap.add_argument("--no-vae", action="store_true", help="feed latents directly (VAE encoder is a 'next' row)")
    return ap.parse_args()


def stage1_args(tokens):
    return argparse.Namespace(layers=[0, 1, 2, 3], noise_level=-1, device="cuda", top_k=10, furthest_point_num_samples=25,
                              sigma=2.0, num_subjects=1, top_k_strategy="gaussian", equivariance_attn_loss_weight=1000.0,
                              sharpening_loss_weight=100.0, num_tokens=tokens, lr=5e-3)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU oracle arm
def oracle_iteration_factory(tokens, res):
    from oracle import hotpath as hp
    from oracle import sd15
    torch.set_num_threads(os.cpu_count() or 1)
    pipe = sd15.make_pipeline(seed=0, attn_gain=4.0)
    ldm, controllers, _ = hp.load_oracle_ldm(pipe, res)
    g = torch.Generator().manual_seed(2)
    context = torch.randn(1, tokens, 768, generator=g).requires_grad_(True)
    image = hp.synthetic_image(seed=1, size=512)
    m, v = torch.zeros_like(context), torch.zeros_like(context)
    state = {"step": 0}

    def one_iteration():
        theta = hp.sample_affine_params(1)
        hp.stage1_iteration(ldm, controllers, image, context, theta, top_k=10, num_candidates=25, sigma=2.0)
        state["step"] += 1
        with torch.no_grad():
            hp.adam_step(context, context.grad, m, v, state["step"])
        context.grad = None

    return one_iteration


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    it = oracle_iteration_factory(a.tokens, a.res)
    for _ in range(a.warmup):
        it()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        it()
    dt = time.perf_counter() - t0
    val = a.steps / dt
    cores = torch.get_num_threads()
    sample = f"{a.steps} Stage-1 iterations (1 image each: 2 captured fwd incl. VAE + selection + losses + backward + Adam), N={a.tokens}, R={a.res}, fp32"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(a, cpu=True),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(a, cpu=False):
    return {"workload": "cfg2: CelebA-wild-like synthetic 512x512 embedding optimisation, K=10 of N tokens, batch=1 per GPU",
            "tokens": a.tokens, "feature_upsample_res": a.res, "top_k": 10, "candidates": 25,
            "precision": "fp32" if cpu else a.precision, "trunk": "cpu" if cpu else a.trunk, "early_exit": bool(a.early_exit) and not cpu,
            "cuda_graph": (not cpu) and bool(getattr(a, "graph", False)),
            "streams": "cpu" if cpu else ("3 inside the step graph: the two captured forwards (+ their backwards) overlap, and the VAE "
                                           "encodes of the NEXT image are prefetched during this step (one-step software pipeline; every "
                                           "timed step still runs 2 VAE encodes + 2 UNet forwards + backward + Adam)"
                                           if os.environ.get("SKP_VAE_PREFETCH", "1") != "0" and not getattr(a, "no_vae", False) else
                                           "2 inside the step graph (the two captured forwards overlap)"),
            "vae_encode": "included" if (cpu or not a.no_vae) else "skipped (latents fed directly)",
            "l2": "no flush needed: 3.4 GB of fp32 UNet weights are streamed every forward (>> 126 MB L2)",
            "weights": "random-init SD1.5-shaped (no checkpoints offline)"}


# --------------------------------------------------------------------------------------------- B200 arm
ALGO = {
    # entry point -> (bound, function(meta) -> algorithmic bytes or flops per launch) for the roofline of the top kernel
}


def run_b200_arm(a):
    import torch.distributed as dist
    from stablekeypoints_b200 import _lib, optimize, optimize_token
    from stablekeypoints_b200.invertable_transform import RandomAffineWithInverse
    from stablekeypoints_b200.optimize import SyntheticKeypointDataset

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ldm, controllers, _ = optimize_token.load_ldm(f"cuda:{local}", "synthetic:0", feature_upsample_res=a.res, attn_gain=4.0,
                                                  precision=a.precision, trunk=a.trunk)
    ldm.unet.early_exit = a.early_exit
    args = stage1_args(a.tokens)
    g = torch.Generator().manual_seed(2)
    context = torch.randn(1, a.tokens, 768, generator=g).to(dev).requires_grad_(True)
    opt = optimize.EmbeddingOptimizer(context, lr=args.lr, capturable=True)
    tr = RandomAffineWithInverse(degrees=15, scale=(0.8, 1.0), translate=(0.25, 0.25))
    ds = SyntheticKeypointDataset(length=8, seed=1 + rank)
    host_imgs = [ds[i]["img"][None].contiguous().pin_memory() for i in range(4)]
    dev_imgs = [h.to(dev) for h in host_imgs]
    if a.no_vae:
        from stablekeypoints_b200 import ptp_utils
        dev_imgs = [ptp_utils.image2latent(ldm, x, "cuda") for x in dev_imgs]
    torch.manual_seed(1000 + rank)

    def eager_step(img):
        out = optimize.stage1_iteration(ldm, controllers, img, context, tr, args, accum=1)
        opt.step()
        opt.zero_grad()
        return out

    step = eager_step

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graph = g2 = ee_step = None
    launches_per_step = None
    if a.graph:
        # the whole optimizer step (2 captured forwards + selection + losses + backward + all-reduce + Adam) as ONE
        # CUDA graph; per-step inputs (image, theta) go through static buffers
        graph = optimize.Stage1Graph(ldm, controllers, context, opt, args, image_shape=tuple(dev_imgs[0].shape))
        graph.set_inputs(dev_imgs[0], tr.sample_theta(1))
        l0 = _lib.launch_count()
        graph.capture()
        launches_per_step = (_lib.launch_count() - l0) // (graph._warmup + 1)
        graph.set_inputs(dev_imgs[0], tr.sample_theta(1))
        graph.prime()                                   # VAE prefetch pipeline: encode the first image ahead of step 0

        def step(img):  # noqa: F811
            graph.set_inputs(img, tr.sample_theta(1))   # pinned host -> static device buffer (async H2D) or D2D
            return graph.replay()

    def timed(n, feed_host):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        s.record()
        for i in range(n):
            if feed_host:
                out = step(host_imgs[i % len(host_imgs)])      # H2D of the pinned image inside the step
                float(out["loss"])                             # D2H read of the step's result
            else:
                step(dev_imgs[i % len(dev_imgs)])
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        nl = launches_per_step * n if launches_per_step is not None else _lib.launch_count() - l0
        return float(ms.item()), nl

    for i in range(a.warmup):
        step(dev_imgs[i % len(dev_imgs)])
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, launches = timed(a.steps, feed_host=False)
    clk = clocks.stop() if rank == 0 else None
    value = world * a.steps / (ms / 1e3)
    ms_e2e, _ = timed(a.steps, feed_host=not a.no_vae)
    e2e = world * a.steps / (ms_e2e / 1e3)

    # ---- per-kernel shares of one step (our C-ABI launches bracketed by CUDA events) + attn-store kernel roofline
    extra = {}
    # the profiled step contains the gradient all-reduce: EVERY rank must run it (only rank 0 reports)
    _lib.start_profile()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    eager_step(dev_imgs[0])
    e.record()
    prof = _lib.stop_profile()
    step_ms = s.elapsed_time(e)
    if rank == 0:
        shares = {k: {"calls": len(v), "ms": round(sum(v), 4)} for k, v in sorted(prof.items(), key=lambda kv: -sum(kv[1]))}
        extra["skp_kernel_ms_in_one_profiled_step"] = shares
        extra["profiled_eager_step_ms"] = round(step_ms, 3)
        host_only = ("skp_gemm_nt_tc_plan", "skp_self_attn_dp", "skp_gemm_tc_force_bn")
        extra["skp_kernel_ms_total_eager"] = round(sum(sum(v) for k, v in prof.items() if k not in host_only), 3)
        extra["profile_note"] = ("per-entry-point CUDA-event brackets of ONE eager (un-graphed, CPU-launch-bound) step: use the "
                                 "shares, not the absolute ms; the timed region above replays the whole step as one CUDA graph")
        extra["roofline"] = attn_store_roofline(a, dev)
        extra["roofline_gemm"] = gemm_roofline(dev)
        if a.early_exit is False and world == 1:
            ldm.unet.early_exit = True
            ee_step = eager_step
            if a.graph:
                g2 = optimize.Stage1Graph(ldm, controllers, context, opt, args, image_shape=tuple(dev_imgs[0].shape))
                g2.set_inputs(dev_imgs[0], tr.sample_theta(1))
                g2.capture()
                g2.set_inputs(dev_imgs[0], tr.sample_theta(1))
                g2.prime()

                def ee_step(img):
                    g2.set_inputs(img, tr.sample_theta(1))
                    return g2.replay()
            for i in range(2):
                ee_step(dev_imgs[i % len(dev_imgs)])
            ms_ee, _ = timed_single(ee_step, dev_imgs, a.steps)
            extra["early_exit_images_per_s_1gpu"] = round(a.steps / (ms_ee / 1e3), 3)
            ldm.unet.early_exit = False
    if world > 1:
        dist.barrier()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(a), "clocks": clk,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 0 if a.no_vae else 3 * 512 * 512 * 4 * world,
                    "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": launches,
        }
        line.update(extra)
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a)
        print(json.dumps(line), flush=True)
    if world > 1:
        # the step graphs hold a captured NCCL all-reduce: release them BEFORE the communicator goes away (destroying the
        # process group underneath a live graph is what used to hang), then tear down normally
        dist.barrier()
        torch.cuda.synchronize()
        graph = g2 = step = ee_step = None     # noqa: F841
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def timed_single(step, imgs, n):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(n):
        step(imgs[i % len(imgs)])
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e), 0


# dram__bytes_read.sum + dram__bytes_write.sum of ONE capture_store_row_kernel launch (N=77, R=128, s=16) from the
# `ncu --set full` capture summarised in profiles/r01_attn_store_row.md
ATTN_STORE_DRAM_TRAFFIC = 1372928
ATTN_STORE_TRAFFIC_NOTE = ("inside the kernel the 40.4 MB store is absorbed by the 126 MB L2 (write-back happens after the kernel): "
                           "DRAM traffic during the launch is 0.66 MB read + 0.71 MB written")


def attn_store_roofline(a, dev):
    """The attn-store kernel (skp_capture_store_fwd) on the C=1280 captured layer shape: HBM-bound, algorithmic bytes =
    the probability store heads*R^2*N*4 (+ the low-res logits read), timed live with CUDA events, L2 flushed between."""
    from stablekeypoints_b200 import ops
    peaks = {}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    peak, src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")
    heads, s, n, r = 8, 16, a.tokens, a.res
    logits = torch.randn(heads, s * s, n, device=dev) * 3
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    times = []
    for i in range(8):
        flush.fill_(float(i))
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        ops.capture_store(logits, r)
        en.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(st.elapsed_time(en))
    ms = sum(times) / len(times)
    algo = heads * r * r * n * 4 + heads * s * s * n * 4
    ach = algo / (ms * 1e-3) / 1e9
    # BASELINE cfg5 (SDXL-shaped capture: 20 heads, 32x32 layer, R=256): a 404 MB store that does not fit the 126 MB L2
    lg5 = torch.randn(20, 32 * 32, n, device=dev) * 3
    t5 = []
    for i in range(6):
        flush.fill_(float(i))
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        ops.capture_store(lg5, 256)
        en.record()
        torch.cuda.synchronize()
        if i >= 2:
            t5.append(st.elapsed_time(en))
    ms5 = sum(t5) / len(t5)
    algo5 = 20 * 256 * 256 * n * 4 + lg5.numel() * 4
    cfg5 = {"shape": "h=20, s=32 -> R=256, N=%d (BASELINE cfg5, SDXL-shaped)" % n, "achieved": round(algo5 / (ms5 * 1e-3) / 1e9, 1),
            "frac": round(algo5 / (ms5 * 1e-3) / 1e9 / peak, 4), "algorithmic_bytes": algo5, "ms_per_launch": round(ms5, 5)}
    del lg5
    return {"kernel": "skp_capture_store_fwd (attn-store, C=1280 layer: h=8, s=16 -> R=%d, N=%d)" % (r, n), "bound": "hbm", "cfg5": cfg5,
            "achieved": round(ach, 1), "peak": peak, "peak_source": src, "unit": "GB/s", "frac": round(ach / peak, 4),
            "traffic": ATTN_STORE_DRAM_TRAFFIC if (n, r) == (77, 128) else None, "traffic_note": ATTN_STORE_TRAFFIC_NOTE,
            "algorithmic_bytes": algo, "ms_per_launch": round(ms, 5), "l2": "flushed (512 MiB fill) between launches"}


def gemm_roofline(dev):
    """The kernel with the largest share of the step: the tcgen05 split-bf16 GEMM (every frozen conv / linear of UNet + VAE).
    Tensor-bound; algorithmic FLOPs = 2*M*N*K (each is ISSUED three times -- hi.hi, hi.lo, lo.hi -- for fp32-grade accuracy,
    so the algorithmic ceiling is peak/3).  Shape: the 128^2 x 512 -> 512 VAE 3x3 convolution as an implicit GEMM."""
    from stablekeypoints_b200 import ops
    peaks = {}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    peak, src = (peaks.get("bf16_tflops"), "measured burst (MEASURED_PEAKS.json)") if peaks.get("bf16_tflops") else (1590.0, "fallback (B200_PROFILING.md)")
    h = w = 128
    cin = cout = 512
    x = torch.randn(h * w, cin, device=dev)
    fcw = ops.FrozenConv3x3(torch.randn(cout, cin, 3, 3, device=dev) / (9 * cin) ** 0.5, need_dgrad=False)
    hi, lo = ops.split_bf16(x)
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    times = []
    for i in range(8):
        flush.fill_(float(i))
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        ops.conv3x3_implicit(hi, lo, h, w, fcw.fwd_split, cout)
        en.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(st.elapsed_time(en))
    ms = sum(times) / len(times)
    flops = 2.0 * h * w * cout * 9 * cin
    ach = flops / (ms * 1e-3) / 1e12
    return {"kernel": "skp_conv3x3_tc / gemm_nt_tc_kernel (tcgen05 split-bf16 implicit GEMM, 128x128x512 -> 512, 3x3)", "bound": "tensor",
            "achieved": round(ach, 1), "achieved_issued": round(3 * ach, 1), "peak": peak, "peak_source": src, "unit": "TFLOP/s",
            "frac": round(ach / peak, 4), "frac_issued": round(3 * ach / peak, 4), "traffic": None, "algorithmic_flops": flops,
            "ms_per_launch": round(ms, 5), "l2": "flushed (512 MiB fill) between launches"}


def cpu_baseline(a):
    it = oracle_iteration_factory(a.tokens, a.res)
    it()  # warm-up (allocator, threads)
    t0 = time.perf_counter()
    it()
    dt = time.perf_counter() - t0
    return {"value": 1.0 / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"1 Stage-1 iteration after 1 warm-up (1 image: 2 captured fwd incl. VAE + selection + losses + backward + Adam), N={a.tokens}, R={a.res}, fp32 CPU oracle"}


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)


if __name__ == "__main__":
    main()
