"""GPU (torch) renderer for the synthetic streams of oracle/synth.py -- bench infrastructure only.
Same ray-cast box-room model as Stream.render, evaluated in float32 on the device so that
thousands of 640x480 frames can be produced in seconds.  (The parity tests use the numpy renderer.)"""
import numpy as np
import torch


class TorchRenderer:
    def __init__(self, stream, device):
        self.s = stream
        self.dev = device
        self.tex = torch.from_numpy(stream.textures).to(device)            # (6,T,T) f32
        fx, fy, cx, cy = stream.K
        xs = (torch.arange(stream.w, device=device, dtype=torch.float32) - cx) / fx
        ys = (torch.arange(stream.h, device=device, dtype=torch.float32) - cy) / fy
        Y, X = torch.meshgrid(ys, xs, indexing='ij')
        self.rays = torch.stack([X, Y, torch.ones_like(X)], dim=-1)        # (h,w,3) camera-frame directions
        self.room = torch.tensor(stream.ROOM, device=device, dtype=torch.float32)
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(stream.seed * 100003 + 17)

    @torch.no_grad()
    def render(self, t):
        s = self.s
        R_wc = torch.tensor(s.rotation(t) @ s.R_bc, device=self.dev, dtype=torch.float32)
        c = torch.tensor(s.position(t) + s.rotation(t) @ s.p_bc, device=self.dev, dtype=torch.float32)
        d = self.rays @ R_wc.T
        tpos = (self.room - c) / d
        tneg = (-self.room - c) / d
        tt = torch.where(d > 0, tpos, tneg)
        tmin, axis = torch.min(tt, dim=-1)
        hit = c + d * tmin[..., None]
        sign = torch.gather(d, -1, axis[..., None])[..., 0] > 0
        wall = axis * 2 + sign.long()
        a1, a2 = (axis + 1) % 3, (axis + 2) % 3
        u = torch.gather(hit, -1, a1[..., None])[..., 0]
        v = torch.gather(hit, -1, a2[..., None])[..., 0]
        e1, e2 = self.room[a1], self.room[a2]
        T = s.TEX
        tu = ((u + e1) / (2 * e1) * (T - 1)).clamp(0, T - 1.001)
        tv = ((v + e2) / (2 * e2) * (T - 1)).clamp(0, T - 1.001)
        iu, iv = tu.floor().long(), tv.floor().long()
        fu, fv = tu - iu, tv - iv
        tex = self.tex
        val = (tex[wall, iv, iu] * (1 - fu) * (1 - fv) + tex[wall, iv, iu + 1] * fu * (1 - fv) +
               tex[wall, iv + 1, iu] * (1 - fu) * fv + tex[wall, iv + 1, iu + 1] * fu * fv)
        val = val + torch.randn(val.shape, device=self.dev, generator=self.gen) * 1.5
        return val.round().clamp(0, 255).to(torch.uint8)

    def render_sequence(self, n_frames, out=None):
        out = out if out is not None else torch.empty((n_frames, self.s.h, self.s.w), dtype=torch.uint8, device=self.dev)
        for k in range(n_frames):
            out[k] = self.render(self.s.frame_time(k))
        return out
