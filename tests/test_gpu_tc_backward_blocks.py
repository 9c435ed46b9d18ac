"""GPU unit tests of the tensor-core backward building blocks (through debug entry points of the C ABI)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("P,Fa,Fb,ks", [(128, 128, 64, 1), (1000, 256, 320, 1), (4096, 512, 576, 3), (777, 128, 256, 2)])
def test_weight_gradient_gemm_matches_torch(P, Fa, Fb, ks):
    """dW = Ya^T Xb with K = points on MN-major UMMA operands in the point-atom layout.
    Inputs are rounded to fp16 by the packer; products accumulate in fp32: tolerance 1e-4 of max |ref|."""
    from satnerf_b200 import capi_dev
    g = torch.Generator().manual_seed(P + Fa)
    xa = (torch.randn(P, Fa, generator=g) * 0.5).cuda()
    xb = torch.randn(P, Fb, generator=g).cuda()
    got = capi_dev.debug_dw_gemm(xa, xb, ks)
    ref = xa.half().double().T @ xb.half().double()
    err = float((got.double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-4, err


def test_forward_stash_matches_oracle_activations():
    """Training-mode tensor-core forward: stashed activations (atoms) and activation derivatives (yb: cos(y), 30 cos(30 y) for
    trunk layer 0) against the CPU oracle's fp32 activations.  fp16 storage + fp16-operand GEMMs: 1e-2 absolute on the
    activations (|a| <= 1), 2e-2 of the derivative's scale."""
    import satnerf_b200 as sb
    from satnerf_b200 import capi
    from gpu_util import make_args
    from oracle import render_oracle as orc
    from stash_util import layout, unpack_atoms, unpack_yb
    import torch.nn.functional as F
    H, S, R = 128, 64, 6
    args = make_args(fc_units=H, n_samples=S, precision="tc")
    torch.manual_seed(11)
    field = sb.load_model(args)
    emb = torch.nn.Embedding(30, 4)
    p = {k: v.detach().clone() for k, v in field.state_dict().items()}
    rays, ts = orc.synthetic_sat_rays(R, seed=12)
    z = orc.stratified_depths(rays[:, 6:7], rays[:, 7:8], S, torch.rand(R, S, generator=torch.Generator().manual_seed(13)))
    xyz = (rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]).reshape(-1, 3)
    sun = rays[:, 8:11].repeat_interleave(S, 0)
    te = emb.weight.detach()[ts].repeat_interleave(S, 0)
    acts, pres = [], []
    h = xyz
    for i in range(8):
        if i == 4:
            h = torch.cat([xyz, h], -1)
        y = F.linear(h, p[f"fc_net.{2 * i}.weight"], p[f"fc_net.{2 * i}.bias"]) * (30.0 if i == 0 else 1.0)
        h = torch.sin(y); acts.append(h); pres.append(y)
    feat = F.linear(h, p["feats_from_xyz.weight"], p["feats_from_xyz.bias"])
    field = field.cuda()
    pd = capi.PassDesc(R, S, 11, 0, capi.FP16_TC, 0.0, 0, 0, 0.0)
    nbytes = capi.render_stash_bytes(field.desc, pd)
    stash = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    outs = {k: torch.empty(s, device="cuda") for k, s in dict(rgb=(R, 3), depth=(R,), weights=(R, S), transparency=(R, S), albedo=(R, S, 3),
                                                              sun=(R, S, 1), sky=(R, S, 3), beta=(R, S, 1), sigma=(R, S)).items()}
    capi.render_forward(field.desc, pd, dict(params=field.flat_params(), rays=rays.cuda(), z_vals=z.cuda(), t_emb=emb.weight.detach()[ts].cuda(),
                                             stash=stash, **outs))
    torch.cuda.synchronize()
    n_tiles = (R + 1) // 2                     # S=64: groups of 2 rays = one 128-point tile
    lay = layout(8, H, n_tiles)
    assert lay["total"] + 1024 == nbytes
    P = R * S
    for l in range(8):
        a = unpack_atoms(stash, lay[f"a{l}"], n_tiles, H)[:P].cpu()
        assert (a - acts[l]).abs().max() < 1e-2, (l, float((a - acts[l]).abs().max()))
        w0 = 30.0 if l == 0 else 1.0
        c = unpack_yb(stash, lay[f"y{l}"], n_tiles, H)[:P].cpu()
        assert (c - w0 * torch.cos(pres[l])).abs().max() < 2e-2 * w0, l
    f16 = unpack_atoms(stash, lay["feat"], n_tiles, H)[:P].cpu()
    assert (f16 - feat).abs().max() < 1e-2 * max(1.0, float(feat.abs().max()))
    e = unpack_atoms(stash, lay["e"], n_tiles, 64)[:P].cpu()
    assert (e[:, 0:3] - xyz).abs().max() < 1e-3 and (e[:, 3:6] - sun).abs().max() < 1e-3 and (e[:, 6:10] - te).abs().max() < 3e-3
    assert (e[:, 10] == 1).all() and (e[:, 11:] == 0).all()
    # head first-layer activations
    s1 = torch.sin(F.linear(torch.cat([feat, sun], -1), p["sun_v_net.0.weight"], p["sun_v_net.0.bias"]))
    s2 = torch.sin(F.linear(s1, p["sun_v_net.2.weight"], p["sun_v_net.2.bias"]))
    s3 = torch.sin(F.linear(s2, p["sun_v_net.4.weight"], p["sun_v_net.4.bias"]))
    r1 = torch.sin(F.linear(feat, p["rgb_from_xyzdir.0.weight"], p["rgb_from_xyzdir.0.bias"]))
    b1 = torch.sin(F.linear(torch.cat([feat, te], -1), p["beta_from_xyz.0.weight"], p["beta_from_xyz.0.bias"]))
    for name, ref in (("s1", s1), ("s2", s2), ("s3", s3), ("r1", r1), ("b1", b1)):
        got = unpack_atoms(stash, lay[name], n_tiles, H // 2)[:P].cpu()
        assert (got - ref).abs().max() < 1.5e-2, (name, float((got - ref).abs().max()))
