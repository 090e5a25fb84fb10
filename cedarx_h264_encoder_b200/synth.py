"""Deterministic synthetic moving-pattern clips (BASELINE.json configs): integer only, stateless per
pixel, so the same frames come out of this torch implementation (CPU or CUDA tensors) and of the
C generator the golden model uses (oracle/h264_golden.c gm_synth_frame; equality is a test).

Textured background translating (+3, +2) px/frame, a 128x128 foreground block moving (-5, +4)
px/frame with wrap-around, +-2 hash noise, slowly drifting chroma ramps."""
import torch


def _tri256(v):
    v = v & 255
    return torch.where(v < 128, v, 255 - v)


def synth_frame(width, height, t, fmt=0, device="cpu"):
    """Returns (luma [h, w] uint8, chroma [h/2 or h, w] uint8 interleaved CbCr) for frame index t."""
    y = torch.arange(height, device=device, dtype=torch.int64).view(-1, 1)
    x = torch.arange(width, device=device, dtype=torch.int64).view(1, -1)
    xs, ys = x + 3 * t, y + 2 * t
    v = 40 + (_tri256(xs * 2 + ys) >> 1) + ((((xs >> 3) ^ (ys >> 3)) & 7) * 6) + (((xs * ys) >> 6) & 7)
    fx, fy = (100 - 5 * t) % width, (60 + 4 * t) % height
    rx, ry = (x - fx) % width, (y - fy) % height
    fg = 200 - (_tri256(rx * 4 + ry * 2) >> 1) + ((rx ^ ry) & 15)
    v = torch.where((rx < 128) & (ry < 128), fg, v)
    m = 0xFFFFFFFF
    h = (x * 0x9E3779B1 + y * 0x85EBCA77 + t * 0xC2B2AE3D) & m
    h = h ^ (h >> 15)
    h = (h * 0x2C1B3C6D) & m
    h = h ^ (h >> 12)
    v = v + ((h >> 8) % 5) - 2
    luma = v.clamp(0, 255).to(torch.uint8)
    crows = height if fmt == 1 else height // 2
    r = torch.arange(crows, device=device, dtype=torch.int64).view(-1, 1)
    cy = (r >> 1) if fmt == 1 else r
    cx = torch.arange(width // 2, device=device, dtype=torch.int64).view(1, -1)
    u = (128 + ((cx + t) & 63) - 32).expand(crows, -1)
    vv = (128 + ((cy + 2 * t) & 63) - 32).expand(-1, width // 2)
    chroma = torch.stack([u, vv], dim=2).reshape(crows, width).to(torch.uint8)
    return luma, chroma


def synth_clip(width, height, frame_indices, fmt=0, device="cpu", out=None):
    """Packed frames [n, w*h + chroma bytes] uint8 exactly as the CLI reads them (luma then chroma)."""
    n = len(frame_indices)
    fb = width * height * (2 if fmt == 1 else 3) // (1 if fmt == 1 else 2)
    if out is None:
        out = torch.empty((n, fb), dtype=torch.uint8, device=device)
    for i, t in enumerate(frame_indices):
        luma, chroma = synth_frame(width, height, int(t), fmt, device)
        out[i, :width * height] = luma.reshape(-1).to(out.device)
        out[i, width * height:] = chroma.reshape(-1).to(out.device)
    return out
