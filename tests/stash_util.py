"""Test-side decoder of the tensor-core forward's activation stash (layout of csrc/tc_common.cuh::TcStash)."""
import torch


def layout(n_layers, H, n_tiles):
    H2 = H // 2
    tH, tH2 = (H // 64) * 16384, ((H2 + 63) // 64) * 16384
    yH, yH2 = 128 * H * 2, 128 * H2 * 2
    off, out = 0, {}

    def take(name, per_tile):
        nonlocal off
        out[name] = off
        off = (off + per_tile * n_tiles + 1023) & ~1023

    for l in range(n_layers):
        take(f"a{l}", tH); take(f"y{l}", yH)
    take("feat", tH)
    for k in ("r1", "s1", "s2", "s3", "b1"):
        take(k, tH2)
    for k in ("r1y", "s1y", "s2y", "s3y", "b1y"):
        take(k, yH2)
    take("e", 16384)
    out["total"] = off
    return out


def unpack_atoms(buf, off, n_tiles, F):
    """atoms [gt][fg][pg][8][64] fp16 with 16-byte chunk XOR swizzle -> (n_tiles*128, F) float."""
    fgs = (F + 63) // 64
    raw = buf[off:off + n_tiles * fgs * 16384].view(torch.float16).view(n_tiles, fgs, 16, 8, 8, 8)   # [.., pg, p8, chunk, 8]
    p8 = torch.arange(8, device=buf.device).view(8, 1)
    ch = torch.arange(8, device=buf.device).view(1, 8)
    src = (ch ^ p8)                                           # logical chunk c of row p8 lives at physical chunk c ^ p8
    out = torch.gather(raw, 4, src.view(1, 1, 1, 8, 8, 1).expand(n_tiles, fgs, 16, 8, 8, 8))
    out = out.permute(0, 2, 3, 1, 4, 5).reshape(n_tiles * 128, fgs * 64)
    return out[:, :F].float()


def unpack_yb(buf, off, n_tiles, F):
    """yb [gt][F/8][128][8] fp16 activation derivatives w0 cos(w0 y) -> (n_tiles*128, F) float."""
    raw = buf[off:off + n_tiles * 128 * F * 2].view(torch.float16).view(n_tiles, F // 8, 128, 8)
    return raw.permute(0, 2, 1, 3).reshape(n_tiles * 128, F).float()
