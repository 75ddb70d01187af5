"""Mirror of Stage 2 of the reference's ``unsupervised_keypoints/keypoint_regressor.py`` (find_best_indices :16-108;
SURVEY "next" row f3).  The regressors of Stage 4 are tiny NumPy pinv fits and stay with the reference."""
from __future__ import annotations

import torch

from . import ptp_utils
from .optimize import _make_dataset


def vote_top_k(indices_list: torch.Tensor, top_k: int) -> torch.Tensor:
    """keypoint_regressor.py:101-106: the top_k most frequently selected token ids (torch.unique order breaks ties)."""
    indices, counts = torch.unique(indices_list, return_counts=True)
    return indices[counts.argsort(descending=True)][:top_k]


@torch.no_grad()
def find_best_indices(ldm, context, args, controllers, num_gpus, from_where=["down_cross", "mid_cross", "up_cross"]):
    """keypoint_regressor.py:16-108: `num_indices` no-grad captured forwards, per image the Gaussian-KL candidates and a
    furthest-point sample measured on the SAME maps (quirk: Stage 1 measures on the transformed maps), then a vote."""
    dataset = _make_dataset(args, stage=2)
    loader = torch.utils.data.DataLoader(dataset, batch_size=num_gpus, shuffle=True, drop_last=True)
    it = iter(loader)
    picked = []
    for _ in range(args.num_indices // num_gpus):
        try:
            batch = next(it)
        except StopIteration:
            it = iter(loader)
            batch = next(it)
        maps = ptp_utils.run_and_find_attn(ldm, batch["img"], context, layers=args.layers, noise_level=args.noise_level,
                                           from_where=from_where, upsample_res=args.feature_upsample_res,
                                           controllers=controllers, device=args.device)
        for m in maps:
            if args.top_k_strategy == "entropy":
                cand = ptp_utils.entropy_sort(m, args.furthest_point_num_samples)
            elif args.top_k_strategy == "gaussian":
                cand = ptp_utils.find_top_k_gaussian(m, args.furthest_point_num_samples, sigma=args.sigma,
                                                     num_subjects=args.num_subjects)
            elif args.top_k_strategy == "consistent":
                cand = torch.arange(args.furthest_point_num_samples, device=m.device)
            else:
                raise NotImplementedError
            picked.append(ptp_utils.furthest_point_sampling(m, args.top_k, cand))
    return vote_top_k(torch.cat(picked).cpu(), args.top_k)
