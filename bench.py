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
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="skip the 1-GPU context lines (full forward, N=100/500, batch_size=4 accumulation, torch eager)")
    return ap.parse_args()


def stage1_args(tokens, top_k=10, candidates=25):
    return argparse.Namespace(layers=[0, 1, 2, 3], noise_level=-1, device="cuda", top_k=top_k, furthest_point_num_samples=candidates,
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
        "impl_detail": {"precision": "fp32", "trunk": "cpu oracle port (oracle/sd15.py + oracle/hotpath.py)", "cuda_graph": False},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(a, cpu=False):
    """The WORKLOAD only (identical for both arms; implementation descriptors live in `impl_detail`)."""
    return {"workload": "cfg2: CelebA-wild-like synthetic 512x512 embedding optimisation, K=10 of N tokens, batch=1 per GPU",
            "tokens": a.tokens, "feature_upsample_res": a.res, "top_k": 10, "candidates": 25,
            "vae_encode": "included" if (cpu or not a.no_vae) else "skipped (latents fed directly)",
            "l2": "no flush needed: 3.4 GB of fp32 UNet weights are streamed every forward (>> 126 MB L2)",
            "weights": "random-init SD1.5-shaped (no checkpoints offline)"}


def impl_detail(a):
    from stablekeypoints_b200 import ptp_utils
    return {"precision": a.precision, "trunk": a.trunk, "cuda_graph": bool(a.graph),
            "early_exit": bool(ptp_utils.EARLY_EXIT or a.early_exit),
            "early_exit_note": "run_and_find_attn stops the UNet after the 4th captured layer: the reference discards pred_noise "
                               "(ptp_utils.py:246), outputs are identical (tests/test_gpu_pipeline.py::test_tiny_early_exit_same_maps); "
                               "the full-forward rate is reported as full_forward_images_per_s_1gpu",
            "streams": ("3 inside the step graph: the two captured forwards (+ their backwards) overlap, and the VAE encodes of the NEXT "
                        "image are prefetched during this step (one-step software pipeline; every timed step still runs 2 VAE encodes + "
                        "2 UNet forwards + backward + Adam)" if os.environ.get("SKP_VAE_PREFETCH", "1") != "0" and not a.no_vae else
                        "2 inside the step graph (the two captured forwards overlap)")}


# --------------------------------------------------------------------------------------------- B200 arm
class Stage1Runner:
    """One rank's Stage-1 loop for the bench: graph (default) or eager steps over device / pinned-host images."""

    def __init__(self, ldm, controllers, tokens, dev, rank, graph=True, accum=1, no_vae=False, image_size=512, top_k=10,
                 candidates=25):
        from stablekeypoints_b200 import optimize, ptp_utils
        from stablekeypoints_b200.invertable_transform import RandomAffineWithInverse
        from stablekeypoints_b200.optimize import SyntheticKeypointDataset
        self.ldm, self.controllers, self.accum = ldm, controllers, accum
        self.args = stage1_args(tokens, top_k, candidates)
        g = torch.Generator().manual_seed(2)
        self.context = torch.randn(1, tokens, ldm.unet.cfg.cross_attention_dim, generator=g).to(dev).requires_grad_(True)
        self.opt = optimize.EmbeddingOptimizer(self.context, lr=self.args.lr, capturable=True)
        self.tr = RandomAffineWithInverse(degrees=15, scale=(0.8, 1.0), translate=(0.25, 0.25))
        ds = SyntheticKeypointDataset(length=8, size=image_size, seed=1 + rank)
        self.host_imgs = [ds[i]["img"][None].contiguous().pin_memory() for i in range(4)]
        self.dev_imgs = [h.to(dev) for h in self.host_imgs]
        if no_vae:
            self.dev_imgs = [ptp_utils.image2latent(ldm, x, "cuda") for x in self.dev_imgs]
        self.optimize = optimize
        self.graph = None
        self.launches_per_iter = None
        if graph:
            from stablekeypoints_b200 import _lib
            self.graph = optimize.Stage1Graph(ldm, controllers, self.context, self.opt, self.args, image_shape=tuple(self.dev_imgs[0].shape),
                                              accum=accum)
            self.graph.set_inputs(self.dev_imgs[0], self.tr.sample_theta(1))
            l0 = _lib.launch_count()
            self.graph.capture()
            self.launches_per_iter = (_lib.launch_count() - l0) // (self.graph._warmup + 1)
            self.graph.set_inputs(self.dev_imgs[0], self.tr.sample_theta(1))
            self.graph.prime()                              # VAE prefetch pipeline: encode the first image ahead of step 0
        self._it = 0

    def eager_iteration(self, img):
        out = self.optimize.stage1_iteration(self.ldm, self.controllers, img, self.context, self.tr, self.args, accum=self.accum)
        self._it += 1
        if self._it % self.accum == 0:
            self.opt.step()
            self.opt.zero_grad()
            self.ldm.unet.invalidate_context_cache()
        return out

    def iteration(self, img):
        """One image: set_inputs (pinned host -> static device buffer is an async H2D copy) + graph replay, or the eager step."""
        if self.graph is None:
            return self.eager_iteration(img)
        self.graph.set_inputs(img, self.tr.sample_theta(1))
        return self.graph.replay()

    def close(self):
        self.graph = None


def run_b200_arm(a):
    import torch.distributed as dist
    from stablekeypoints_b200 import _lib, optimize_token, ptp_utils

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
    torch.manual_seed(1000 + rank)
    run = Stage1Runner(ldm, controllers, a.tokens, dev, rank, graph=a.graph, no_vae=a.no_vae)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(r, n, feed_host, sync_ranks=True):
        """n images through r (device-resident or pinned-host fed with the loss read back); CUDA events, max over ranks."""
        barrier() if sync_ranks else torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        s.record()
        for i in range(n):
            if feed_host:
                out = r.iteration(r.host_imgs[i % len(r.host_imgs)])   # H2D of the pinned image inside the step
                float(out["loss"])                                     # D2H read of the step's result
            else:
                r.iteration(r.dev_imgs[i % len(r.dev_imgs)])
        e.record()
        barrier() if sync_ranks else torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1 and sync_ranks:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        nl = r.launches_per_iter * n if r.launches_per_iter is not None else _lib.launch_count() - l0
        return float(ms.item()), nl

    for i in range(a.warmup):
        run.iteration(run.dev_imgs[i % len(run.dev_imgs)])
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, launches = timed(run, a.steps, feed_host=False)
    clk = clocks.stop() if rank == 0 else None
    value = world * a.steps / (ms / 1e3)
    ms_e2e, _ = timed(run, a.steps, feed_host=not a.no_vae)
    e2e = world * a.steps / (ms_e2e / 1e3)

    # ---- per-kernel shares of one step (our C-ABI launches bracketed by CUDA events) + rooflines + context lines
    extra = {}
    # the profiled step contains the gradient all-reduce: EVERY rank must run it (only rank 0 reports)
    _lib.start_profile()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    run.eager_iteration(run.dev_imgs[0])
    e.record()
    prof = _lib.stop_profile()
    step_ms = s.elapsed_time(e)
    if rank == 0:
        shares = {k: {"calls": len(v), "ms": round(sum(v), 4)} for k, v in sorted(prof.items(), key=lambda kv: -sum(kv[1]))}
        extra["skp_kernel_ms_in_one_profiled_step"] = shares
        extra["profiled_eager_step_ms"] = round(step_ms, 3)
        host_only = ("skp_gemm_nt_tc_plan", "skp_self_attn_dp", "skp_gemm_tc_force_bn", "skp_capture_tc_ok", "skp_capture_tc_workspace")
        extra["skp_kernel_ms_total_eager"] = round(sum(sum(v) for k, v in prof.items() if k not in host_only), 3)
        extra["profile_note"] = ("per-entry-point CUDA-event brackets of ONE eager (un-graphed, CPU-launch-bound) step: use the "
                                 "shares, not the absolute ms; the timed region above replays the whole step as one CUDA graph")
        extra["roofline"], extra["roofline_l2_resident"] = attn_store_rooflines(a, dev)
        extra["roofline_gemm"] = gemm_roofline(dev)
    if world == 1 and a.extras:
        # context lines (1 GPU only, never the headline): same loop at other settings; each builds its own step graph
        def rate(tokens=a.tokens, accum=1, graph=a.graph, n=a.steps, env=None, ldm_=None, ctl_=None):
            r = Stage1Runner(ldm_ or ldm, ctl_ or controllers, tokens, dev, rank, graph=graph, accum=accum, no_vae=a.no_vae)
            for i in range(2 * accum):
                r.iteration(r.dev_imgs[i % len(r.dev_imgs)])
            ms_, _ = timed(r, n * accum, feed_host=False, sync_ranks=False)
            r.close()
            del r
            torch.cuda.empty_cache()
            return round(n * accum / (ms_ / 1e3), 3)

        if not a.early_exit and ptp_utils.EARLY_EXIT:
            ptp_utils.EARLY_EXIT = False
            extra["full_forward_images_per_s_1gpu"] = rate()
            ptp_utils.EARLY_EXIT = True
        if a.tokens == 77:
            extra["tokens_100_images_per_s_1gpu"] = rate(tokens=100)
            extra["tokens_500_images_per_s_1gpu"] = rate(tokens=500, n=max(3, a.steps // 2))
        extra["batch4_accum_images_per_s_1gpu"] = rate(accum=4, n=max(2, a.steps // 2))
        extra["batch4_accum_note"] = ("reference CLI default batch_size=4 on one GPU: B//G = 4 accumulated iterations per optimizer step "
                                      "(optimize.py:339,420-425) through the iteration / update CUDA graphs; images/s, to compare with `value`")
        # BASELINE cfg5: the same loop at SDXL-base shapes (1024^2 image, 2048-wide context, captured layers 32x32 / 20 heads x 64,
        # R = 256, K = 16); parity at this shape: tests/test_gpu_pipeline.py::test_cfg5_sdxl_shaped_stage1_vs_oracle
        try:
            extra["cfg5_sdxl_shaped_1gpu"] = cfg5_context(a, dev, rank, timed)
        except Exception as ex:       # context only: never fail the bench line on it
            extra["cfg5_sdxl_shaped_1gpu"] = {"error": repr(ex)[:300]}
        # the honest same-box library bar: the SAME loop with the trunk on cuDNN / cuBLAS (torch eager, no graph), TF32 as the
        # reference's torch defaults allow and strict fp32, each with its parity error against this build's maps
        try:
            extra["torch_eager_b200"] = torch_eager_context(a, dev, rank, ldm, controllers, rate)
        except Exception as ex:       # context only: never fail the bench line on it
            extra["torch_eager_b200"] = {"error": repr(ex)[:300]}
    if world > 1:
        dist.barrier()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(a), "impl_detail": impl_detail(a), "clocks": clk,
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
        run.close()
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def cfg5_context(a, dev, rank, timed):
    """Stage-1 images/s on BASELINE cfg5's shapes: a 3-level UNet of the SDXL-base widths (320, 640, 1280; attention at
    64^2 and 32^2 only; one transformer block per attention module) with a 2048-wide context and 20 heads (head dim 64 at the
    captured 32x32 / C=1280 layers), 1024^2 synthetic images through the SD VAE encoder, R = 256, K = 16 of 77 tokens."""
    from stablekeypoints_b200 import optimize_token
    from stablekeypoints_b200.sd15_engine import UNetConfig, VAEConfig
    ucfg = UNetConfig(block_out_channels=(320, 640, 1280), cross_attention_dim=2048, heads=20, down_has_attn=(False, True, True))
    ldm5, ctl5, _ = optimize_token.load_ldm(str(dev), "synthetic:5", feature_upsample_res=256, attn_gain=4.0, precision=a.precision,
                                            trunk=a.trunk, unet_config=ucfg, vae_config=VAEConfig())
    r = Stage1Runner(ldm5, ctl5, 77, dev, rank, graph=a.graph, image_size=1024, top_k=16, candidates=32)
    for i in range(2):
        r.iteration(r.dev_imgs[i % len(r.dev_imgs)])
    n = max(3, a.steps // 2)
    ms_, _ = timed(r, n, feed_host=False, sync_ranks=False)
    ms_e, _ = timed(r, n, feed_host=True, sync_ranks=False)
    out = {"images_per_s": round(n / (ms_ / 1e3), 3), "ms_per_step": round(ms_ / n, 3), "e2e_images_per_s": round(n / (ms_e / 1e3), 3),
           "config": {"workload": "cfg5: SDXL-shaped 1024x1024 cross-attention capture path, K=16 of 77 tokens, batch=1", "tokens": 77,
                      "context_dim": 2048, "feature_upsample_res": 256, "top_k": 16, "candidates": 32, "captured_layers": "3 x (32x32, C=1280, 20 heads x 64)",
                      "unet": "3 levels (320, 640, 1280), one transformer block per attention module", "vae_encode": "included (1024^2)"},
           "launches_per_step": r.launches_per_iter}
    r.close()
    del r, ldm5, ctl5
    torch.cuda.empty_cache()
    optimize_token.set_precision(a.precision)
    return out


def torch_eager_context(a, dev, rank, ldm, controllers, rate):
    """torch-eager-on-B200 context: library trunk (cuDNN convs, cuBLAS linears, SDPA attention) under the same surface."""
    from stablekeypoints_b200 import optimize_token, ptp_utils
    out = {}
    g = torch.Generator().manual_seed(5)
    img = torch.rand(1, 3, 512, 512, generator=g).to(dev)
    ctx = torch.randn(1, a.tokens, 768, generator=g).to(dev)
    noise = torch.randn(1, 4, 64, 64, generator=g).to(dev)
    with torch.no_grad():
        ref_maps = ptp_utils.run_and_find_attn(ldm, img, ctx, layers=[0, 1, 2, 3], upsample_res=-1, controllers=controllers, noise=noise)[0]
    for prec in ("reference", "fp32"):
        ldm_t, ctl_t, _ = optimize_token.load_ldm(str(dev), "synthetic:0", feature_upsample_res=a.res, attn_gain=4.0, precision=prec,
                                                  trunk="torch")
        with torch.no_grad():
            m = ptp_utils.run_and_find_attn(ldm_t, img, ctx, layers=[0, 1, 2, 3], upsample_res=-1, controllers=ctl_t, noise=noise)[0]
        err = float((m - ref_maps).abs().max() / ref_maps.abs().max())
        out[prec] = {"images_per_s": rate(graph=False, n=3, ldm_=ldm_t, ctl_=ctl_t), "maps_rel_diff_vs_this_build": err,
                     "what": "cuDNN TF32 convolutions + fp32 cuBLAS (the reference's torch defaults)" if prec == "reference"
                             else "strict fp32 cuDNN / cuBLAS"}
        del ldm_t, ctl_t
        torch.cuda.empty_cache()
    optimize_token.set_precision(a.precision)
    out["note"] = "torch eager (no CUDA graph), same surface and loop; context only -- the tolerance of the path is 1e-3"
    return out


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def _ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` summaries
    (profiles/r02_attn_store_traffic.json, written by scripts/ncu_traffic.py from the .ncu-rep of the same command)."""
    p = os.path.join(ROOT, "profiles", "r02_attn_store_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p)).get(key)
    return (d["dram_bytes"], d["source"]) if d else (None, None)


def _time_capture_store(ops, lib, logits, res, flush, tc, reps=5, warm=3):
    lib().skp_capture_tc(2 if tc else 0)
    ts = []
    for i in range(reps + warm):
        flush.fill_(float(i))
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        ops.capture_store(logits, res)
        en.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(st.elapsed_time(en))
    lib().skp_capture_tc(1)
    return sum(ts) / len(ts)


def attn_store_rooflines(a, dev):
    """The attn-store kernel (the kernel BASELINE.json's metric names), HBM-bound: algorithmic bytes = the probability store
    heads*R^2*N*4 + the low-res logits read, timed live with CUDA events on the launching stream, L2 flushed (512 MiB fill)
    between launches.  Both implementations are timed (tcgen05 formulation skp_capture_tc.cu, SIMT register kernel
    skp_capture_store.cu); the line reports the one the library's policy picks for the shape.
      roofline             : BASELINE cfg5's SDXL-shaped layer (20 heads, 32x32 -> R=256): a 404 MB store that does NOT fit
                             the 126 MB L2, i.e. a real HBM stream;
      roofline_l2_resident : the SD1.5 C=1280 captured layer (8 heads, 16x16 -> R=128, N tokens): the 40 MB store is absorbed
                             by L2 inside the launch, the kernel is instruction / latency bound -- GB/s there is an L2 write
                             rate, reported for continuity with round 1, not an HBM fraction."""
    from stablekeypoints_b200 import ops
    from stablekeypoints_b200._lib import lib
    peaks = _peaks()
    peak, src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    n = a.tokens

    def one(heads, s, r, key, label, note):
        logits = torch.randn(heads, s * s, n, device=dev) * 3
        algo = heads * r * r * n * 4 + logits.numel() * 4
        ms_tc = _time_capture_store(ops, lib, logits, r, flush, True) if n <= 128 else None
        ms_simt = _time_capture_store(ops, lib, logits, r, flush, False)
        picked_tc = bool(lib().skp_capture_tc_ok((__import__("ctypes").c_int * 1)(s), 1, n, r, 1))
        ms = ms_tc if (picked_tc and ms_tc is not None) else ms_simt
        ach = algo / (ms * 1e-3) / 1e9
        traffic, tsrc = _ncu_traffic(key + ("_tc" if picked_tc else "_reg")) if n == 77 else (None, None)
        out = {"kernel": "skp_capture_store%s_fwd (attn-store, %s)" % ("_tc" if picked_tc else "", label), "bound": "hbm",
               "achieved": round(ach, 1), "peak": peak, "peak_source": src, "unit": "GB/s", "frac": round(ach / peak, 4),
               "traffic": traffic, "traffic_source": tsrc, "algorithmic_bytes": algo, "ms_per_launch": round(ms, 5),
               "implementation": "tcgen05 (skp_capture_tc.cu)" if picked_tc else "SIMT register kernel (skp_capture_store.cu)",
               "ms_tcgen05": None if ms_tc is None else round(ms_tc, 5), "ms_simt": round(ms_simt, 5),
               "l2": "flushed (512 MiB fill) between launches", "note": note}
        del logits
        return out

    big = one(20, 32, 256, "cfg5", "h=20, s=32 -> R=256, N=%d: BASELINE cfg5, SDXL-shaped" % n,
              "404 MB store > 126 MB L2: an HBM stream")
    small = one(8, 16, a.res, "sd15", "SD1.5 C=1280 layer: h=8, s=16 -> R=%d, N=%d" % (a.res, n),
                "L2-resident: the store is absorbed by L2 during the launch; the kernel is instruction/latency bound, so this "
                "figure is an L2 write rate, not an HBM fraction")
    return big, small


def gemm_roofline(dev):
    """The kernel with the largest share of the step: the tcgen05 split-bf16 GEMM (every frozen conv / linear of UNet + VAE).
    Tensor-bound; algorithmic FLOPs = 2*M*N*K (each is ISSUED three times -- hi.hi, hi.lo, lo.hi -- for fp32-grade accuracy,
    so the algorithmic ceiling is peak/3).  Shape: the 128^2 x 512 -> 512 VAE 3x3 convolution as an implicit GEMM."""
    from stablekeypoints_b200 import ops
    peaks = _peaks()
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
