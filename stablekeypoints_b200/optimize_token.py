"""Mirror of the live part of the reference's ``unsupervised_keypoints/optimize_token.py`` (:24-78, :203-241)."""
from __future__ import annotations

import os

import torch

from . import ptp_utils
from .sd15_engine import (DDIMSchedule, Pipeline, UNetConfig, UNetEngine, VAEConfig, VAEEncoderEngine,
                          synthetic_state_dict, unet_param_shapes, vae_encoder_param_shapes)


def set_precision(mode: str) -> None:
    """Numerics of the torch part of the trunk.  "fp32": strict fp32 everywhere.  "tf32": TF32 allowed for cuDNN
    convs AND cuBLAS matmuls.  "reference": what torch-1.13 defaults gave the reference (TF32 convs, fp32 matmuls)."""
    if mode not in ("fp32", "tf32", "reference"):
        raise ValueError(mode)
    torch.backends.cudnn.allow_tf32 = mode in ("tf32", "reference")
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    # fixed shapes, one image per rank: let cuDNN time its algorithms once (its fp32 heuristics pick FFT convolutions
    # that launch thousands of complex GEMVs per forward)
    torch.backends.cudnn.benchmark = os.environ.get("SKP_CUDNN_BENCHMARK", "1") == "1"


# AutoencoderKL mid-block attention: diffusers 0.8.0 (the reference's pin, requirements.yaml:174) names the projections
# query / key / value / proj_attn (nn.Linear [C, C]); checkpoints re-saved by diffusers >= 0.18 call them
# to_q / to_k / to_v / to_out.0, and CompVis-converted ones keep them as 1x1 convolutions [C, C, 1, 1].
_VAE_ATTN_RENAMES = ((".to_q.", ".query."), (".to_k.", ".key."), (".to_v.", ".value."), (".to_out.0.", ".proj_attn."))


def _normalize_vae_keys(sd):
    """Bring a VAE state dict to the diffusers-0.8.0 key names / shapes the engine's parameter table uses."""
    out = {}
    for k, v in sd.items():
        if ".attentions." in k:
            for new, old in _VAE_ATTN_RENAMES:
                k = k.replace(new, old)
            if k.endswith(".weight") and v.dim() == 4 and v.shape[-2:] == (1, 1) and any(n in k for n in (".query.", ".key.", ".value.", ".proj_attn.")):
                v = v.reshape(v.shape[0], v.shape[1])
        out[k] = v
    return out


def _load_safetensors_dir(path: str, sub: str):
    from safetensors.torch import load_file
    sd = None
    for name in ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.fp16.safetensors"):
        f = os.path.join(path, sub, name)
        if os.path.exists(f):
            sd = load_file(f)
            break
    if sd is None:
        f = os.path.join(path, sub, "diffusion_pytorch_model.bin")
        if os.path.exists(f):
            sd = torch.load(f, map_location="cpu")
    if sd is None:
        raise FileNotFoundError(f"no diffusers weights under {os.path.join(path, sub)}")
    return _normalize_vae_keys(sd) if sub == "vae" else sd


def load_ldm(device, type="CompVis/stable-diffusion-v1-4", feature_upsample_res=256, my_token=None, *,
             unet_state_dict=None, vae_state_dict=None, unet_config: UNetConfig = None, vae_config: VAEConfig = None,
             seed: int = 0, attn_gain: float = 1.0, precision: str = None, trunk: str = None):
    """optimize_token.py:24-78 -> (ldm, controllers, effective_num_gpus).

    One process drives ONE GPU (the multi-GPU layout is one process per GPU + NCCL, not nn.DataParallel), so
    ``controllers`` has exactly one AttentionStore keyed by this process's device and effective_num_gpus == 1.
    ``type`` is a local diffusers-format directory (unet/, vae/ safetensors), "synthetic[:seed]" for seeded random
    weights of the SD1.x shapes (there is no hub access here) or "synthetic-small[:seed]" for a 1/10-width model of the
    same topology; explicit state dicts override all of them."""
    if str(device) == "cpu":
        raise RuntimeError("stablekeypoints_b200 has no CPU path (the reference's CPU path is timed from oracle/)")
    dev = torch.device(device if str(device) != "cuda" else f"cuda:{torch.cuda.current_device()}")
    if precision is not None:
        set_precision(precision)
    if unet_state_dict is None and unet_config is None and type.startswith("synthetic-small"):
        # same topology at 1/10 of the widths (768-wide context like SD1.x, so the reference's init_random_noise fits):
        # lets the reference's own main.py run end to end in seconds where no checkpoint is available
        unet_config = UNetConfig(block_out_channels=(32, 64, 128, 128), heads=4, norm_num_groups=8)
        vae_config = vae_config or VAEConfig(block_out_channels=(8, 16, 32, 32), norm_num_groups=4)
        attn_gain = 6.0 if attn_gain == 1.0 else attn_gain      # peaky (trained-looking) maps instead of near-uniform ones
    ucfg, vcfg = unet_config or UNetConfig(), vae_config or VAEConfig()
    if unet_state_dict is None:
        if type.startswith("synthetic"):
            if ":" in type:
                seed = int(type.split(":")[1])
            unet_state_dict = synthetic_state_dict(unet_param_shapes(ucfg), dev, seed, attn_gain)
            vae_state_dict = synthetic_state_dict(vae_encoder_param_shapes(vcfg), dev, seed + 1)
        elif os.path.isdir(type):
            unet_state_dict = _load_safetensors_dir(type, "unet")
            vae_state_dict = _load_safetensors_dir(type, "vae")
        else:
            raise FileNotFoundError(f"model '{type}' is not a local directory and there is no network: pass a "
                                    "diffusers-format directory or 'synthetic[:seed]'")
    trunk = trunk or os.environ.get("SKP_TRUNK", "tc")
    unet = UNetEngine(unet_state_dict, ucfg, dev, trunk=trunk)
    vae = VAEEncoderEngine(vae_state_dict, vcfg, dev, trunk=trunk)
    ldm = Pipeline(unet, vae, DDIMSchedule(dev))
    controllers = {dev: ptp_utils.AttentionStore()}
    ptp_utils.register_attention_control(unet, controllers[dev], feature_upsample_res=feature_upsample_res)
    return ldm, controllers, 1


def gaussian_circle(pos, size=64, sigma=16, device="cuda"):
    """optimize_token.py:203-223 (API parity; the loss kernels evaluate this target on the fly)."""
    _pos = (pos * size).unsqueeze(1).unsqueeze(1)
    ar = torch.arange(size, device=pos.device).float() + 0.5
    rows, cols = ar.reshape(1, size, 1), ar.reshape(1, 1, size)
    d2 = (cols - _pos[..., 1]) ** 2 + (rows - _pos[..., 0]) ** 2
    return torch.exp(-1 * d2 / (2.0 * sigma ** 2.0))


def gaussian_circles(pos, size=64, sigma=16, device="cuda"):
    """optimize_token.py:225-241: pos [num_points, batch, 2] -> mean over points."""
    return torch.stack([gaussian_circle(pos[i], size=size, sigma=sigma, device=device) for i in range(pos.shape[0])]).mean(0)
