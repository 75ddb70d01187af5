"""Mirror of the live part of the reference's ``unsupervised_keypoints/optimize.py`` (collect_maps :27-79, losses
:157-206, optimize_embedding :269-452) on the B200 kernels, one process per GPU."""
from __future__ import annotations

import os
import time
from typing import Optional

import torch

from . import ops, ptp_utils
from .invertable_transform import RandomAffineWithInverse, invert_theta


def collect_maps(controller, from_where=["up_cross"], upsample_res=512, layers=[0, 1, 2, 3], indices=None):
    """optimize.py:27-79: mean over the selected stored layers and the B*h axis, optional token gather and bilinear
    resize, returns [N or K, R', R'] and RESETS the controller (side effect at :77).  `from_where` is accepted and
    ignored exactly like the reference."""
    stored = [m for i, m in enumerate(controller.step_store["attn"]) if i in layers]
    if not stored:
        raise RuntimeError("collect_maps: the controller holds no stored attention maps for layers %s" % (layers,))
    if indices is not None:
        indices = torch.as_tensor(indices)
    r = int(stored[0].shape[1] ** 0.5)
    n_tok = stored[0].shape[2] if indices is None else indices.numel()
    # the reference's guard compares sqrt(#tokens) with upsample_res (optimize.py:63)
    resize = upsample_res != -1 and n_tok ** 0.5 != upsample_res
    out = ops.collect_maps_op(stored, upsample_res if resize else -1, indices)
    controller.reset()
    return out


def equivariance_loss(embeddings_initial, embeddings_transformed, transform, index):
    """optimize.py:157-163: MSE(initial, transform.inverse(transformed)[index]); transformed is [G,K,R,R]."""
    theta_inv = invert_theta(transform.last_params["theta"])[int(index)]
    k = embeddings_initial.shape[0]
    sel = torch.arange(k, device=embeddings_initial.device)
    return ops.equivariance_loss_op(embeddings_initial, embeddings_transformed[int(index)], sel, theta_inv)


def sharpening_loss(attn_map, sigma=1.0, temperature=1e1, device="cuda", num_subjects=1):
    """optimize.py:166-179: MSE(map, Gaussian centred on the map's own arg-max)."""
    sel = torch.arange(attn_map.shape[0], device=attn_map.device)
    return ops.sharpen_loss_op(attn_map, sel, sigma, num_subjects)


def find_gaussian_loss_at_point(attn_map, pos, sigma=1.0, temperature=1e-1, device="cuda", indices=None, num_subjects=1):
    """optimize.py:182-206 with explicit positions: pos [num, T, 2] in [0,1] -> nearest pixel-centre peaks."""
    _, h, w = attn_map.shape
    cell = (pos.to(attn_map.device) * torch.tensor([h, w], device=attn_map.device)).floor().long()
    peaks = (cell[..., 0].clamp(0, h - 1) * w + cell[..., 1].clamp(0, w - 1)).contiguous()
    sel = torch.arange(attn_map.shape[0], device=attn_map.device) if indices is None else torch.as_tensor(indices)
    if indices is not None:
        peaks = peaks[:, sel.to(peaks.device)].contiguous()
    return ops._SharpenLoss.apply(attn_map, sel.to(attn_map.device, torch.int64).contiguous(), peaks, sigma)


# ----------------------------------------------------------------------------- one Stage-1 iteration
def stage1_losses(attn_map, attn_map_t, theta, *, top_k=10, num_candidates=25, sigma=2.0, num_subjects=1,
                  top_k_strategy="gaussian", forced_indices=None, theta_inv=None):
    """optimize.py:380-401 for this rank's image: candidates by Gaussian-KL on the ORIGINAL maps, furthest-point
    sampling on the TRANSFORMED maps' arg-maxes, then both losses on the selected tokens -- all on device, the
    token indices never visit the host."""
    if forced_indices is not None:
        idx = torch.as_tensor(forced_indices, device=attn_map.device)
    else:
        if top_k_strategy == "entropy":
            cand = ptp_utils.entropy_sort(attn_map, num_candidates)
        elif top_k_strategy == "gaussian":
            cand = ptp_utils.find_top_k_gaussian(attn_map, num_candidates, sigma=sigma, num_subjects=num_subjects)
        elif top_k_strategy == "consistent":
            cand = torch.arange(num_candidates, device=attn_map.device)
        else:
            raise NotImplementedError
        idx = ptp_utils.furthest_point_sampling(attn_map_t, top_k, cand)
    sharp = ops.sharpen_loss_op(attn_map, idx, sigma, num_subjects)
    if theta_inv is None:                       # host path; a device theta_inv keeps the step CUDA-graph capturable
        theta_inv = invert_theta(theta)
    equiv = ops.equivariance_loss_op(attn_map, attn_map_t, idx, theta_inv[0])
    return idx, sharp, equiv


def stage1_iteration(ldm, controllers, image, context, transform: RandomAffineWithInverse, args, *, accum: int = 1,
                     theta=None, theta_inv=None, noise_a=None, noise_b=None, forced_indices=None, from_where=None,
                     side_stream: Optional["torch.cuda.Stream"] = None, latents=None):
    """optimize.py:341-422 for one rank: two captured forwards, selection, loss = w_e*equiv + w_s*sharp, / accum,
    backward into ``context`` (its .grad accumulates).

    side_stream: the two captured forwards (original / warped image) only share the K|V projection of the embedding, and
    at one image per rank most of their kernels are far too small to fill 148 SMs.  With a side stream the warped
    image's forward (and, through autograd's stream tracking, its backward) runs concurrently with the original's.
    latents = (latent, latent_of_warped_image): VAE encodes done ahead of time (Stage1Graph's prefetch stage); `image`
    is ignored, `theta` / `theta_inv` must be the warp those latents were made with."""
    kw = dict(layers=args.layers, noise_level=args.noise_level, from_where=from_where, upsample_res=-1,
              device=args.device, controllers=controllers)
    dev = ldm.unet.device
    if latents is not None:
        image, transformed_pre = latents       # 4-channel tensors pass through image2latent (ptp_utils.py:293-298 contract)
        transform.last_params = {"theta": theta}
    else:
        transformed_pre = None
        image = image.to(dev, non_blocking=True) if isinstance(image, torch.Tensor) else image
    if side_stream is None or not isinstance(image, torch.Tensor):
        attn_maps = ptp_utils.run_and_find_attn(ldm, image, context, noise=noise_a, **kw)
        transformed_img = transform(image, theta=theta) if transformed_pre is None else transformed_pre
        attn_maps_t = ptp_utils.run_and_find_attn(ldm, transformed_img, context, noise=noise_b, **kw)
    else:
        main = torch.cuda.current_stream()
        transformed_img = transform(image, theta=theta) if transformed_pre is None else transformed_pre
        ldm.unet.project_context(context)            # shared K|V projection: once, before the fork (both forwards hit the cache)
        side_stream.wait_stream(main)
        attn_maps = ptp_utils.run_and_find_attn(ldm, image, context, noise=noise_a, **kw)
        with torch.cuda.stream(side_stream):
            attn_maps_t = ptp_utils.run_and_find_attn(ldm, transformed_img, context, noise=noise_b, **kw)
        main.wait_stream(side_stream)
        for t in attn_maps_t:                        # produced on the side stream, consumed (and freed) on the main one
            t.record_stream(main)
    idx, sharp, equiv = stage1_losses(attn_maps[0], attn_maps_t[0], transform.last_params["theta"], top_k=args.top_k,
                                      num_candidates=args.furthest_point_num_samples, sigma=args.sigma,
                                      num_subjects=args.num_subjects, top_k_strategy=args.top_k_strategy,
                                      forced_indices=forced_indices, theta_inv=theta_inv)
    loss = equiv * args.equivariance_attn_loss_weight + sharp * args.sharpening_loss_weight
    (loss / accum).backward()
    return {"loss": loss.detach(), "sharp": sharp.detach(), "equiv": equiv.detach(), "indices": idx,
            "maps": attn_maps[0].detach(), "maps_t": attn_maps_t[0].detach()}


def allreduce_sum_(grad: torch.Tensor, group=None) -> int:
    """The only collective of the path: sum of d(context) over ranks (NCCL on GPUs, gloo in the CPU tests).
    Returns the world size so the caller can fold the 1/world mean into the optimizer kernel."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    ws = dist.get_world_size(group)
    if ws > 1:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
    return ws


def rank_shard(n_items: int, rank: int, world: int, epoch: int = 0, seed: int = 0):
    """Indices of the dataset this rank visits in `epoch` (same partition DistributedSampler(shuffle, drop_last) makes)."""
    g = torch.Generator().manual_seed(seed + epoch)
    perm = torch.randperm(n_items, generator=g).tolist()
    per = n_items // world
    return perm[: per * world][rank::world]


class EmbeddingOptimizer:
    """Adam on the [1,N,D] embedding (optimize.py:320,424-425) with the data-parallel gradient mean fused in:
    one NCCL all-reduce(sum) of the N*D fp32 gradient per optimizer step, then skp_adam_step with grad_scale =
    1/world_size on every rank (replicated state, identical updates)."""

    def __init__(self, context: torch.Tensor, lr: float = 5e-3, betas=(0.9, 0.999), eps: float = 1e-8, group=None,
                 capturable: bool = False):
        self.context, self.lr, self.betas, self.eps, self.group = context, lr, betas, eps, group
        self.exp_avg = torch.zeros_like(context)
        self.exp_avg_sq = torch.zeros_like(context)
        self.steps = 0
        # capturable: the step count lives on the device so the update can be replayed from a CUDA graph
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=context.device) if capturable else None

    def step(self):
        g = self.context.grad
        ws = allreduce_sum_(g, self.group)
        self.steps += 1
        with torch.no_grad():
            ops.adam_step_(self.context, g, self.exp_avg, self.exp_avg_sq,
                           self.step_dev if self.step_dev is not None else self.steps, self.lr, self.betas[0],
                           self.betas[1], self.eps, 1.0 / ws)
        # the kernel wrote through the raw pointer: tell autograd / the engine's K|V cache the tensor changed
        torch.autograd.graph.increment_version(self.context)

    def zero_grad(self, set_to_none: bool = True):
        """set_to_none=False zeroes the gradient buffer in place: the buffer keeps its address, which is what lets separate
        CUDA graphs (iterations / optimizer update) accumulate into and read from it (B//G > 1, optimize.py:420-425)."""
        if set_to_none or self.context.grad is None:
            self.context.grad = None
        else:
            self.context.grad.zero_()


class Stage1Graph:
    """One whole Stage-1 optimizer step (optimize.py:341-425 for this rank: two captured forwards, selection, losses,
    backward, gradient all-reduce, Adam) captured ONCE into a CUDA graph and replayed per image.

    The step is ~10^4 small launches at batch 1, i.e. CPU-launch-bound when driven from Python; the graph removes that.
    Everything data-dependent stays on the device (token selection included), the per-step inputs are written into
    static buffers (`image`, `theta`, `theta_inv`) before each replay, noise comes from torch's graph-safe Philox state."""

    def __init__(self, ldm, controllers, context, optimizer: "EmbeddingOptimizer", args, image_shape=(1, 3, 512, 512),
                 accum: int = 1, warmup: int = 3, from_where=None):
        assert optimizer.step_dev is not None, "Stage1Graph needs EmbeddingOptimizer(capturable=True)"
        if args.top_k_strategy in ("gaussian", "entropy", "consistent") and args.furthest_point_num_samples < args.top_k:
            raise ValueError("Stage1Graph: furthest_point_num_samples < top_k makes the number of selected tokens data-dependent "
                             "(ptp_utils.py:139-159); run the eager loop (cuda_graph=False)")
        # accum == 1: ONE graph holds the whole optimizer step.  accum > 1 (the reference's B//G gradient accumulation,
        # optimize.py:339,420-425): an ITERATION graph (two captured forwards + selection + losses + backward, the gradient
        # added into a persistent buffer) replayed `accum` times, then an UPDATE graph (all-reduce + Adam + zeroing).
        self.accum = int(accum)
        self._iters_since_update = 0
        self.update_graph = None
        dev = ldm.unet.device
        self.ldm, self.controllers, self.context, self.optimizer, self.args = ldm, controllers, context, optimizer, args
        self.image = torch.zeros(image_shape, device=dev)
        self.theta = torch.zeros(image_shape[0], 2, 3, device=dev)
        self.theta_inv = torch.zeros(image_shape[0], 2, 3, device=dev)
        self.transform = RandomAffineWithInverse()
        self.from_where = from_where
        self.out = None
        self.graph = None
        self._warmup = warmup
        # second stream inside the graph: the two forwards / backwards of a step overlap (SKP_TWO_STREAM=0 disables)
        # the two UNet streams run at high priority, the VAE prefetch stream at the default (lower) one, so the small
        # latency-bound UNet kernels get SMs ahead of the large VAE convolutions (stream capture records the priority per
        # kernel node): 44.7 -> 46.8 images/s on B200.  SKP_STREAM_PRIORITY=0 captures everything at the default priority.
        self._prio = -1 if os.environ.get("SKP_STREAM_PRIORITY", "1") != "0" else 0
        self.side = torch.cuda.Stream(device=dev, priority=self._prio) if os.environ.get("SKP_TWO_STREAM", "1") != "0" else None
        # VAE prefetch (SKP_VAE_PREFETCH=0 disables): the two VAE encodes of an image do not depend on the embedding, so the
        # graph of step i encodes the image of step i+1 on a third stream while the UNet work of step i -- mostly kernels
        # too small to fill the GPU -- runs on the other two.  One-step software pipeline: replay() returns the losses of
        # the image given to the PREVIOUS set_inputs(); prime() fills the pipeline, flush() drains it.
        self.prefetch = (os.environ.get("SKP_VAE_PREFETCH", "1") != "0" and len(image_shape) == 4 and image_shape[1] == 3
                         and getattr(ldm, "vae", None) is not None)
        if self.prefetch:
            lat_shape = (2, 4, image_shape[2] // 8, image_shape[3] // 8)
            self.vae_stream = torch.cuda.Stream(device=dev)
            self.lat_next = torch.zeros(lat_shape, device=dev)          # [latent(image), latent(warped image)] of the NEXT step
            self.lat_cur = torch.zeros(lat_shape, device=dev)
            self.theta_next, self.theta_inv_next = torch.zeros_like(self.theta), torch.zeros_like(self.theta_inv)
            self.theta_cur, self.theta_inv_cur = torch.zeros_like(self.theta), torch.zeros_like(self.theta_inv)

    @property
    def latency(self) -> int:
        """Steps between set_inputs(image) and the replay() that returns that image's losses."""
        return 1 if self.prefetch else 0

    def _encode_next(self):
        """Prefetch stage: VAE-encode the input image and its warp by the input theta into the pipeline registers
        (no grad; ptp_utils.py:289-304, invertable_transform.py:38-70)."""
        with torch.no_grad():
            self.theta_next.copy_(self.theta)
            self.theta_inv_next.copy_(self.theta_inv)
            warped = self.transform(self.image, theta=self.theta)
            self.lat_next[0:1].copy_(ptp_utils.image2latent(self.ldm, self.image, self.args.device))
            self.lat_next[1:2].copy_(ptp_utils.image2latent(self.ldm, warped, self.args.device))

    def prime(self):
        """Fill the pipeline with the image / theta last given to set_inputs (eager, outside the graph)."""
        if self.prefetch:
            self._encode_next()
        return self

    def _update(self):
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=self.accum == 1)
        self.ldm.unet.invalidate_context_cache()

    def _step(self):
        out = self._iteration()
        self._update()
        return out

    def _iteration(self):
        if not self.prefetch:
            out = stage1_iteration(self.ldm, self.controllers, self.image, self.context, self.transform, self.args,
                                   accum=self.accum, theta=self.theta, theta_inv=self.theta_inv, from_where=self.from_where,
                                   side_stream=self.side)
        else:
            main = torch.cuda.current_stream()
            with torch.no_grad():          # roll the pipeline registers: what was prefetched becomes this step's input
                self.lat_cur.copy_(self.lat_next)
                self.theta_cur.copy_(self.theta_next)
                self.theta_inv_cur.copy_(self.theta_inv_next)
            self.vae_stream.wait_stream(main)
            with torch.cuda.stream(self.vae_stream):
                self._encode_next()
            out = stage1_iteration(self.ldm, self.controllers, None, self.context, self.transform, self.args,
                                   accum=self.accum, theta=self.theta_cur, theta_inv=self.theta_inv_cur,
                                   from_where=self.from_where, side_stream=self.side,
                                   latents=(self.lat_cur[0:1], self.lat_cur[1:2]))
            main.wait_stream(self.vae_stream)
        return out

    def set_inputs(self, image: torch.Tensor, theta: torch.Tensor):
        """image: [1,3,H,W] (host pinned or device); theta: [1,2,3] host tensor (inverse computed here in fp64)."""
        self.image.copy_(image, non_blocking=True)
        self.theta.copy_(theta.to(torch.float32), non_blocking=True)
        self.theta_inv.copy_(invert_theta(theta), non_blocking=True)

    def capture(self):
        self.optimizer.zero_grad()
        opt = self.optimizer
        saved = [t.clone() for t in (self.context.detach(), opt.exp_avg, opt.exp_avg_sq, opt.step_dev)]
        saved_steps = opt.steps
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        if self.accum > 1:                     # persistent gradient buffer shared by the two graphs
            self.context.grad = torch.zeros_like(self.context)
        with torch.cuda.stream(side):
            for _ in range(self._warmup):      # allocator warm-up, time-constant caches
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():                  # the warm-up steps must not count as training steps
            for dst, src in zip((self.context, opt.exp_avg, opt.exp_avg_sq, opt.step_dev), saved):
                dst.copy_(src)
        opt.steps = saved_steps
        torch.autograd.graph.increment_version(self.context)
        self.ldm.unet.invalidate_context_cache()
        self.graph = torch.cuda.CUDAGraph()
        cap = dict(stream=torch.cuda.Stream(device=self.image.device, priority=self._prio)) if self._prio != 0 else {}
        if self.accum == 1:
            with torch.cuda.graph(self.graph, **cap):
                self.out = self._step()
        else:
            with torch.cuda.graph(self.graph, **cap):
                self.out = self._iteration()
            self.update_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.update_graph, pool=self.graph.pool()):
                self._update()
        self._iters_since_update = 0
        return self

    def replay(self):
        """Replays one iteration (accum == 1: the whole optimizer step; accum > 1: the update graph follows every
        `accum`-th iteration).  With the VAE prefetch the returned losses belong to the image of the previous set_inputs()."""
        self.graph.replay()
        if self.update_graph is not None:
            self._iters_since_update += 1
            if self._iters_since_update == self.accum:
                self.update_graph.replay()
                self._iters_since_update = 0
                torch.autograd.graph.increment_version(self.context)
        return self.out


class SyntheticKeypointDataset(torch.utils.data.Dataset):
    """{"img": [3,S,S] float in [0,1]} of seeded Gaussian blobs + low-frequency noise: the output contract of the
    reference datasets (datasets/celeba.py:94-114) for the synthetic 512^2 configs of BASELINE.json."""

    def __init__(self, length: int = 64, size: int = 512, seed: int = 1, blobs: int = 12):
        self.length, self.size, self.seed, self.blobs = length, size, seed, blobs

    def __len__(self):
        return self.length

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 100003 + i)
        s = self.size
        ys = ((torch.arange(s).float() + 0.5) / s).reshape(1, s, 1)
        xs = ((torch.arange(s).float() + 0.5) / s).reshape(1, 1, s)
        img = torch.zeros(3, s, s)
        for _ in range(self.blobs):
            cy, cx = torch.rand(2, generator=g).tolist()
            sd = 0.03 + 0.08 * torch.rand(1, generator=g).item()
            img += torch.rand(3, 1, 1, generator=g) * torch.exp(-((ys - cy) ** 2 + (xs - cx) ** 2) / (2 * sd * sd))
        low = torch.nn.functional.interpolate(torch.rand(1, 3, 8, 8, generator=g), size=(s, s), mode="bilinear",
                                              align_corners=False)[0]
        return {"img": (0.7 * img + 0.3 * low).clamp(0, 1), "kpts": torch.zeros(1, 2), "visibility": torch.ones(1)}


def _make_dataset(args, stage: int = 1):
    """optimize.py:277-303 (stage 1) / keypoint_regressor.py:25-50 (stage 2: the same table without CelebA's max_len).
    The reference's dataset classes are host-side file readers that are out of scope here; they are imported from the
    user's ``datasets`` package when present, `args.dataset` (an object) wins, and "synthetic" builds the seeded
    synthetic set."""
    if getattr(args, "dataset", None) is not None:
        return args.dataset
    name = args.dataset_name
    if name == "synthetic":
        ml = getattr(args, "max_len", -1)
        return SyntheticKeypointDataset(length=ml if ml is not None and ml > 0 else 64, size=getattr(args, "synthetic_size", 512))
    import importlib
    ml = dict(max_len=args.max_len) if stage == 1 else {}
    table = {"celeba_aligned": ("datasets.celeba", "CelebA", dict(split="train", dataset_loc=args.dataset_loc, **ml)),
             "celeba_wild": ("datasets.celeba", "CelebA", dict(split="train", dataset_loc=args.dataset_loc, align=False, **ml)),
             "cub_aligned": ("datasets.cub", "TrainSet", dict(data_root=args.dataset_loc, image_size=512)),
             "cub_001": ("datasets.cub_parts", "CUBDataset", dict(dataset_root=args.dataset_loc, split="train", single_class=1)),
             "cub_002": ("datasets.cub_parts", "CUBDataset", dict(dataset_root=args.dataset_loc, split="train", single_class=2)),
             "cub_003": ("datasets.cub_parts", "CUBDataset", dict(dataset_root=args.dataset_loc, split="train", single_class=3)),
             "cub_all": ("datasets.cub_parts", "CUBDataset", dict(dataset_root=args.dataset_loc, split="train")),
             "taichi": ("datasets.taichi", "TrainSet", dict(data_root=args.dataset_loc, image_size=512)),
             "human3.6m": ("datasets.human36m", "TrainSet", dict(data_root=args.dataset_loc, validation=getattr(args, "validation", False))),
             "unaligned_human3.6m": ("datasets.unaligned_human36m", "TrainSet", dict(data_root=args.dataset_loc, image_size=512)),
             "deepfashion": ("datasets.deepfashion", "TrainSet", dict(data_root=args.dataset_loc, image_size=512)),
             "custom": ("datasets.custom_images", "CustomDataset", dict(data_root=args.dataset_loc, image_size=512))}
    if name not in table:
        raise NotImplementedError
    mod, cls, kw = table[name]
    return getattr(importlib.import_module(mod), cls)(**kw)


def optimize_embedding(ldm, args, controllers, num_gpus, context=None,
                       from_where=["down_cross", "mid_cross", "up_cross"]):
    """optimize.py:269-452.  Same loop, same flags (`args` is the reference's argparse Namespace), same printed
    scalars.  Data parallelism: this process owns one GPU (num_gpus == 1 here); under torch.distributed every rank
    draws its own shard of the shuffled dataset and the embedding gradient is all-reduced once per optimizer step,
    which reproduces the reference's mean over devices (optimize.py:405-406) and its B//G accumulation (:420-425)."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    dataset = _make_dataset(args)
    transform = RandomAffineWithInverse(degrees=args.augment_degrees, scale=args.augment_scale,
                                        translate=args.augment_translate)
    dev = ldm.unet.device
    if context is None:
        context = ptp_utils.init_random_noise(dev, num_words=args.num_tokens)
        if world > 1:
            dist.broadcast(context, src=0)
    context = context.to(dev).detach().clone().contiguous()
    context.requires_grad = True
    accum = max(1, args.batch_size // (num_gpus * world))
    # every iteration replays a 3-stream CUDA graph (Stage1Graph); with gradient accumulation (B//G > 1, the reference CLI
    # default batch_size=4) the Adam update is a second small graph replayed every `accum` iterations
    use_graph = (getattr(args, "cuda_graph", os.environ.get("SKP_LOOP_GRAPH", "1") != "0")
                 and args.furthest_point_num_samples >= args.top_k)
    trace = getattr(args, "trace", None)       # optional list: per-iteration outputs are appended (tests / debugging)
    optimizer = EmbeddingOptimizer(context, lr=args.lr, capturable=bool(use_graph))
    start = it_start = time.time()
    running = {"equiv": 0.0, "sharp": 0.0, "total": 0.0}
    sampler = None
    if world > 1:
        sampler = torch.utils.data.distributed.DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=True,
                                                                  drop_last=True)
    loader = torch.utils.data.DataLoader(dataset, batch_size=num_gpus, shuffle=sampler is None, sampler=sampler,
                                         drop_last=True, pin_memory=True)
    it = iter(loader)
    epoch = 0

    def next_batch():
        nonlocal it, epoch
        try:
            return next(it)
        except StopIteration:
            epoch += 1
            if sampler is not None:            # DataLoader(shuffle=True) reshuffles every pass (optimize.py:341-345)
                sampler.set_epoch(epoch)
            it = iter(loader)
            return next(it)

    graph = None
    n_iters = int(int(args.num_steps) * accum)
    if use_graph and n_iters > 0:
        first = next_batch()
        graph = Stage1Graph(ldm, controllers, context, optimizer, args, image_shape=tuple(first["img"].shape),
                            accum=accum, from_where=from_where)
        graph.transform = transform
        graph.set_inputs(first["img"], transform.sample_theta(first["img"].shape[0]))
        graph.capture()            # warm-up steps are rolled back (optimizer state and embedding restored)
        graph.prime()              # fill the VAE prefetch stage with the first image
    for iteration in range(n_iters):
        if graph is not None:
            if graph.latency == 1:
                # VAE prefetch: feed the NEXT image (its encodes overlap this step); replay() trains on the current one
                if iteration + 1 < n_iters:
                    nxt = next_batch()
                    graph.set_inputs(nxt["img"], transform.sample_theta(nxt["img"].shape[0]))
            else:
                cur = first if iteration == 0 else next_batch()
                if iteration > 0:
                    graph.set_inputs(cur["img"], transform.sample_theta(cur["img"].shape[0]))
            out = graph.replay()
            out = {k: v.clone() for k, v in out.items() if k in (("loss", "sharp", "equiv", "indices") if trace is not None
                                                                 else ("loss", "sharp", "equiv"))}
        else:
            batch = next_batch()
            out = stage1_iteration(ldm, controllers, batch["img"], context, transform, args, accum=accum,
                                   from_where=from_where)
        if trace is not None:
            trace.append({k: v.detach().clone() for k, v in out.items() if k in ("loss", "sharp", "equiv", "indices")})
        running["equiv"] += out["equiv"] / accum * args.equivariance_attn_loss_weight
        running["sharp"] += out["sharp"] / accum * args.sharpening_loss_weight
        running["total"] += out["loss"] / accum
        if (iteration + 1) % accum == 0:
            if graph is None:              # the graph holds the all-reduce + Adam update itself
                optimizer.step()
                optimizer.zero_grad()
            ldm.unet.invalidate_context_cache()
            if rank == 0:
                msg = {"loss": float(running["total"]), "running_equivariance_attn_loss": float(running["equiv"]),
                       "running_sharpening_loss": float(running["sharp"]), "iteration time": time.time() - it_start}
                if getattr(args, "wandb", False):
                    import wandb
                    wandb.log(msg)
                else:
                    print(f"loss: {float(out['loss'] / accum)}, _loss_equivariance_attn: {msg['running_equivariance_attn_loss']} "
                          f"sharpening_loss: {msg['running_sharpening_loss']}, running_total_loss: {msg['loss']}, "
                          f"iteration time: {msg['iteration time']}")
            running = {"equiv": 0.0, "sharp": 0.0, "total": 0.0}
            it_start = time.time()
    if rank == 0:
        print(f"optimization took {time.time() - start} seconds")
    return context.detach()
