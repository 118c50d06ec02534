"""BASELINE.json configs[4]: CLIP-guided VelocityDiffusion sampling, 50 steps, with the guidance loss as the
measured inner loop.

The UNet is the CALLER of the hot path, not part of it: `VelocityUNet` below is a plain-PyTorch (cuDNN, bf16 autocast)
stand-in with the shape of the reference's cc12m_1 model (perceptor/models/velocity_diffusion/cc12m_1.py:108-302:
base width 128, channel multipliers 1-2-2-4-4-8-8, four modulated residual conv blocks per side of each level, eight
in the innermost one, single-group GroupNorm, self-attention with 64-wide heads from 16x16 down, ViT-B/16-sized CLIP
conditioning), random init -- there is no network for checkpoints.  Everything between the UNet and the next UNet
call is this repo's native path: Predictions.denoised_images -> CLIP guidance loss forward + backward ->
Predictions.guided -> Predictions.step.  Per step the three segments are timed with CUDA events.

    python tools/guided_sampling_bench.py [--model ViT-L-14] [--cutouts 64] [--steps 50]
"""
import argparse
import json
import math
import os
import sys

import torch
from torch import nn
from torch.nn import functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceptor_b200 import losses, velocity_diffusion as vd  # noqa: E402
from perceptor_b200.vit import SHAPES  # noqa: E402

WIDTHS = [128, 256, 256, 512, 512, 1024, 1024]
ATTN_FROM = 4  # levels 4.. (16x16 and coarser) interleave self-attention
COND = 1024


class Fourier(nn.Module):
    def __init__(self, n_out, std=1.0):
        super().__init__()
        self.register_buffer("freq", torch.randn(n_out // 2, 1) * std)

    def forward(self, t):
        f = 2 * math.pi * t[:, None] @ self.freq.T
        return torch.cat([f.cos(), f.sin()], dim=-1)


class ModConvBlock(nn.Module):
    """conv3x3 -> norm -> (1 + scale) x + shift from the conditioning vector -> relu, twice, plus a skip path"""

    def __init__(self, c_in, c_mid, c_out, last=False):
        super().__init__()
        self.conv1, self.conv2 = nn.Conv2d(c_in, c_mid, 3, padding=1), nn.Conv2d(c_mid, c_out, 3, padding=1)
        self.mod1 = nn.Linear(COND, 2 * c_mid, bias=False)
        self.mod2 = None if last else nn.Linear(COND, 2 * c_out, bias=False)
        self.skip = None if c_in == c_out else nn.Conv2d(c_in, c_out, 1, bias=False)

    @staticmethod
    def _modulate(h, lin, cond):
        scale, shift = lin(cond).chunk(2, dim=-1)
        return F.relu(F.group_norm(h, 1) * (scale[..., None, None] + 1) + shift[..., None, None])

    def forward(self, x, cond):
        h = self._modulate(self.conv1(x), self.mod1, cond)
        h = self.conv2(h)
        if self.mod2 is not None:
            h = self._modulate(h, self.mod2, cond)
        return h + (x if self.skip is None else self.skip(x))


class Attention2d(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.heads = c // 64
        self.norm = nn.GroupNorm(1, c)
        self.qkv, self.out = nn.Conv2d(c, 3 * c, 1), nn.Conv2d(c, c, 1)

    def forward(self, x):
        n, c, h, w = x.shape
        q, k, v = self.qkv(self.norm(x)).view(n, 3, self.heads, c // self.heads, h * w).transpose(3, 4).unbind(1)
        y = F.scaled_dot_product_attention(q, k, v).transpose(2, 3).reshape(n, c, h, w)
        return x + self.out(y)


class Level(nn.Module):
    def __init__(self, i, c_in):
        super().__init__()
        c = WIDTHS[i]
        innermost = i == len(WIDTHS) - 1
        c_prev = WIDTHS[i - 1] if i > 0 else 3
        attn = i >= ATTN_FROM

        def stack(chans):
            mods = []
            for j in range(len(chans) - 1):
                last = i == 0 and chans[j + 1] == 3
                mods.append(ModConvBlock(chans[j], c, chans[j + 1], last))
                if attn:
                    mods.append(Attention2d(chans[j + 1]))
            return nn.ModuleList(mods)

        if innermost:
            self.down, self.inner, self.up = stack([c_in] + [c] * 7 + [c_prev]), None, None
        else:
            self.down = stack([c_in] + [c] * 4)
            self.inner = Level(i + 1, c)
            self.up = stack([2 * c, c, c, c, c_prev])

    @staticmethod
    def _run(mods, x, cond):
        for m in mods:
            x = m(x, cond) if isinstance(m, ModConvBlock) else m(x)
        return x

    def forward(self, x, cond):
        x = self._run(self.down, x, cond)
        if self.inner is None:
            return x
        y = F.interpolate(self.inner(F.avg_pool2d(x, 2), cond), scale_factor=2, mode="bilinear", align_corners=False)
        return self._run(self.up, torch.cat([y, x], dim=1), cond)


class VelocityUNet(nn.Module):
    """cc12m_1-shaped velocity model: (x in [-1,1] NCHW 256x256, t [N], clip_embed [N,512]) -> velocities."""

    def __init__(self):
        super().__init__()
        self.t_map, self.t_plane = Fourier(128), Fourier(16)
        self.mapping = nn.Sequential(nn.Linear(512 + 128, COND), nn.ReLU(), nn.Linear(COND, COND), nn.ReLU(),
                                     nn.Linear(COND, COND), nn.ReLU(), nn.Linear(COND, COND))
        self.net = Level(0, 3 + 16)

    def forward(self, x, t, clip_embed):
        clip_embed = F.normalize(clip_embed, dim=-1) * clip_embed.shape[-1] ** 0.5
        cond = self.mapping(torch.cat([clip_embed, self.t_map(t)], dim=1))
        planes = self.t_plane(t)[..., None, None].expand(-1, -1, x.shape[2], x.shape[3])
        return self.net(torch.cat([x, planes], dim=1), cond)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="ViT-L-14")
    ap.add_argument("--cutouts", type=int, default=64)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--channels-last", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    fmt = torch.channels_last if args.channels_last else torch.contiguous_format
    unet = VelocityUNet().to(dev).eval().requires_grad_(False).to(memory_format=fmt)
    n_params = sum(p.numel() for p in unet.parameters())
    shape = SHAPES[args.model]
    clip_loss = losses.CLIP(args.model, n_cutouts=args.cutouts, min_size=args.size // 4, max_size=args.size, seed=0)
    g = torch.Generator().manual_seed(0)
    clip_loss.add_encodings_(torch.randn(2, shape.embed, generator=g))
    clip_loss.model.engine().prebuild_tables(args.size // 4, args.size)
    cond = torch.randn(1, 512, generator=g).to(dev)
    schedule = vd.schedule_ts(n_steps=args.steps)

    def velocities(diffused, t):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            x = vd.encode(diffused).contiguous(memory_format=fmt)
            return unet(x, t.to(dev).expand(diffused.shape[0]), cond).float().contiguous()

    def run(record):
        diffused = vd.decode(torch.randn(1, 3, args.size, args.size, generator=g)).to(dev)
        ev = []
        loss = None
        for from_t, to_t in schedule:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record()
            pred = vd.Predictions(diffused, from_t[None], velocities(diffused, from_t[None]))
            e[1].record()
            with torch.enable_grad():
                v = pred.velocities.requires_grad_()
                den = pred.replace(velocities=v).denoised_images
                e[2].record()
                loss = clip_loss(den)
                (grad,) = torch.autograd.grad(loss, v)
            e[3].record()
            diffused = pred.guided(-grad).step(to_t[None])
            ev.append(e)
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        torch.cuda.synchronize()
        if not record:
            return None
        unet_ms = sum(e[0].elapsed_time(e[1]) for e in ev)
        guide_ms = sum(e[2].elapsed_time(e[3]) for e in ev)
        total_ms = ev[0][0].elapsed_time(end)
        return unet_ms, guide_ms, total_ms, float(loss.detach())

    run(False)
    unet_ms, guide_ms, total_ms, loss = run(True)
    n = len(schedule)
    print(json.dumps({
        "workload": "guided_sampling_cc12m1_shaped_unet", "sampling_steps": n, "image": f"{args.size}x{args.size}",
        "guidance_model": args.model, "cutouts_per_step": args.cutouts, "unet_params_m": round(n_params / 1e6, 1),
        "total_ms": total_ms, "ms_per_sampling_step": total_ms / n, "unet_ms_per_step": unet_ms / n,
        "guidance_ms_per_step": guide_ms / n, "glue_and_host_ms_per_step": (total_ms - unet_ms - guide_ms) / n,
        "guidance_cutouts_per_s_in_loop": args.cutouts * n / (guide_ms * 1e-3),
        "guidance_share_of_loop": guide_ms / total_ms, "final_loss": loss,
        "data": "synthetic (random-init UNet and CLIP weights, random start noise)"}))


if __name__ == "__main__":
    main()
