"""CPU tests of the host logic: the C-ABI library loads and exports every declared symbol, the flat
parameter layout matches the modules, parameter names/shapes/init match the reference, and the product
refuses to run without CUDA (no fallback)."""
import argparse
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _args(**kw):
    base = dict(model="sat-nerf", fc_layers=8, fc_units=64, t_embbeding_tau=4, t_embbeding_vocab=30, n_samples=8,
                n_importance=0, noise_std=0.0, sc_lambda=0.0, chunk=5120)
    base.update(kw)
    return argparse.Namespace(**base)


def test_library_exports_every_declared_symbol():
    from satnerf_b200 import capi
    header = open(os.path.join(ROOT, "include", "satnerf_b200.h")).read()
    declared = set(re.findall(r"SNB_API\s+[\w\s\*]+?\b(snb_\w+)\s*\(", header))
    assert len(declared) >= 13
    lib = capi.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(capi.exported_symbols())
    assert lib.snb_abi_version() == 5


def test_dev_library_exports_its_header():
    """The developer entry points live in a separate library (libsatnerf_b200_dev.so), not in the product ABI."""
    from satnerf_b200 import capi, capi_dev
    header = open(os.path.join(ROOT, "include", "satnerf_b200_dev.h")).read()
    declared = set(re.findall(r"SNB_API\s+[\w\s\*]+?\b(snb_\w+)\s*\(", header))
    assert declared == set(capi_dev.exported_symbols()) and len(declared) == 6
    lib = capi_dev.lib()
    for name in declared:
        assert hasattr(lib, name), name
    prod = capi.lib()
    for name in declared:
        assert not hasattr(prod, name), f"{name} leaked into the product library"


@pytest.mark.parametrize("model,h", [("sat-nerf", 64), ("sat-nerf", 512), ("s-nerf", 256), ("nerf", 256)])
def test_flat_layout_matches_module(model, h):
    import satnerf_b200 as sb
    from satnerf_b200 import capi
    m = sb.load_model(_args(model=model, fc_units=h))
    lay = capi.param_layout(m.desc)
    named = list(m.named_parameters())
    assert len(lay) * 2 == len(named)
    off = 0
    for i, (w, b, n_out, n_in) in enumerate(lay):
        assert named[2 * i][0].endswith(".weight") and tuple(named[2 * i][1].shape) == (n_out, n_in)
        assert w == off; off += n_out * n_in
        assert b == off; off += n_out
    assert off == capi.param_count(m.desc) == sum(p.numel() for p in m.parameters())
    flat = m.flat_params()
    assert flat.numel() == off and named[0][1].data_ptr() == flat.data_ptr()
    # in-place updates through the parameters are visible in the flat buffer (what Adam does)
    with torch.no_grad():
        named[3][1].add_(1.0)
    w3, b3, no3, ni3 = lay[1]
    assert torch.equal(flat[b3:b3 + no3], named[3][1].detach())
    g = m.flat_grads()
    assert all(p.grad is not None and p.grad.data_ptr() >= g.data_ptr() for p in m.parameters())


def test_param_count_sat_nerf_512():
    import satnerf_b200 as sb
    from satnerf_b200 import capi
    assert capi.param_count(sb.load_model(_args(fc_units=512)).desc) == 2635785


@pytest.mark.skipif(not os.path.exists("/root/reference/models"), reason="reference not present")
@pytest.mark.parametrize("model", ["sat-nerf", "s-nerf", "nerf"])
def test_names_shapes_and_init_match_reference(model):
    import importlib
    import sys
    import satnerf_b200 as sb
    sys.path.insert(0, "/root/reference")
    try:
        ref_models = importlib.import_module("models")
    finally:
        sys.path.remove("/root/reference")
    a = _args(model=model, fc_units=128)
    torch.manual_seed(3); ours = sb.load_model(a)
    torch.manual_seed(3); theirs = ref_models.load_model(a)
    sd1, sd2 = ours.state_dict(), theirs.state_dict()
    assert list(sd1) == list(sd2)
    for k in sd1:
        assert torch.equal(sd1[k], sd2[k]), k
    assert ours.number_of_outputs == theirs.number_of_outputs
    ours.load_state_dict(sd2)          # released checkpoints load by name


def test_no_cpu_fallback():
    import satnerf_b200 as sb
    a = _args()
    ms = {"coarse": sb.load_model(a), "t": torch.nn.Embedding(30, 4)}
    rays = torch.rand(4, 11)
    with pytest.raises(RuntimeError, match="CUDA"):
        sb.render_rays(ms, a, rays, torch.zeros(4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        ms["coarse"](torch.rand(4, 3), input_sun_dir=torch.rand(4, 3), input_t=torch.rand(4, 4))


def test_bad_descriptors_are_rejected():
    from satnerf_b200 import capi
    with pytest.raises(ValueError):
        capi.field_desc("foo", 8, 64, [4])
    d = capi.field_desc("sat-nerf", 8, 63, [4], 4)
    with pytest.raises(RuntimeError, match="fc_units"):
        capi.param_count(d)


def test_device_ray_sampler_has_dataloader_semantics():
    """satnerf_b200.data.DeviceRaySampler == DataLoader(dataset, shuffle=True, batch_size=B) of main.py:96-110: every ray once per
    epoch, batches of B with a short last one, the keys and dtypes of SatelliteDataset.__getitem__ (datasets/satellite.py:347-350)."""
    import torch
    from satnerf_b200.data import DeviceRaySampler, combined_loader
    N, B = 1000, 256
    rays = torch.arange(N, dtype=torch.float32)[:, None].repeat(1, 11)
    data = {"rays": rays, "rgbs": torch.rand(N, 3), "ts": torch.randint(0, 17, (N, 1)).float()}
    sm = DeviceRaySampler(data, B, device="cpu", generator=torch.Generator().manual_seed(3))
    assert len(sm) == 4
    seen = []
    for b in sm:
        assert set(b) == {"rays", "rgbs", "ts"} and b["ts"].dtype == torch.int64 and b["rays"].shape[1] == 11
        assert torch.equal(b["rgbs"], data["rgbs"][b["rays"][:, 0].long()])        # rows stay aligned across the tensors
        seen.append(b["rays"][:, 0].long())
    assert [len(s) for s in seen] == [256, 256, 256, 232]
    assert torch.equal(torch.sort(torch.cat(seen)).values, torch.arange(N))         # a permutation: every ray exactly once
    assert not torch.equal(torch.cat(seen), torch.arange(N))
    second = torch.cat([b["rays"][:, 0].long() for b in sm])
    assert not torch.equal(second, torch.cat(seen))                                # reshuffled every epoch
    # rank shards of one global batch are disjoint and cover it
    parts = [next(iter(DeviceRaySampler(data, B, device="cpu", generator=torch.Generator().manual_seed(5), rank=r, world=2))) for r in (0, 1)]
    whole = next(iter(DeviceRaySampler(data, B, device="cpu", generator=torch.Generator().manual_seed(5))))
    assert torch.equal(torch.cat([p["rays"] for p in parts]), whole["rays"])
    # dict of loaders: one epoch = the longest loader, the shorter one restarts
    depth = DeviceRaySampler({"rays": rays[:300], "depths": torch.rand(300, 2), "ts": data["ts"][:300]}, B, device="cpu")
    batches = list(combined_loader({"color": sm, "depth": depth}))
    assert len(batches) == 4 and all(set(b) == {"color", "depth"} for b in batches)
    assert "depths" in batches[0]["depth"] and batches[2]["depth"]["rays"].shape[0] == 256


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The drop-in boundary is a C ABI: every ctypes.Structure of satnerf_b200/capi.py must have the size and the field offsets a C
    compiler gives the struct of include/satnerf_b200.h it mirrors (a probe program is compiled with gcc against the header)."""
    import ctypes as C
    import shutil
    import subprocess
    from satnerf_b200 import capi
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = {"snb_field_desc": capi.FieldDesc, "snb_pass_desc": capi.PassDesc, "snb_loss_desc": capi.LossDesc, "snb_render_io": capi.RenderIO,
             "snb_render_grads": capi.RenderGrads, "snb_rpc_model": capi.RpcModel, "snb_sharded_buffer": capi.ShardedBuffer}
    header = open(os.path.join(ROOT, "include", "satnerf_b200.h")).read()
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "satnerf_b200.h"', "int main(void) {"]
    want = {}
    for cname, cls in pairs.items():
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), header, re.S)
        assert body, cname
        body_txt = re.sub(r"/\*.*?\*/", "", body.group(1), flags=re.S)
        fields = []
        for decl in body_txt.split(";"):                          # `type a, b[20], *c`
            decl = decl.strip()
            if not decl:
                continue
            names = [re.sub(r"\[.*?\]", "", n).strip().lstrip("*").strip() for n in re.sub(r"^.*?([\w\*\[\]]+(?:\s*,\s*[\w\*\[\]]+)*)$", r"\1", decl).split(",")]
            fields += [n for n in names if n]
        py_fields = [f[0].rstrip("_") for f in cls._fields_]
        assert [f.rstrip("_") for f in fields] == py_fields, (cname, fields, py_fields)
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for f, pf in zip(fields, cls._fields_):
            lines.append(f'  printf("{cname} {pf[0]} %zu\\n", offsetof({cname}, {f}));')
        want[cname] = cls
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, stdout=subprocess.PIPE, text=True).stdout
    for ln in out.splitlines():
        cname, what, val = ln.split()
        cls = want[cname]
        if what == "size":
            assert C.sizeof(cls) == int(val), (cname, C.sizeof(cls), val)
        else:
            assert getattr(cls, what).offset == int(val), (cname, what, getattr(cls, what).offset, val)


def test_x3_sine_constants_give_fp32_level_accuracy():
    """The epilogue sine of the hi+lo tensor-core path (x3_sin, csrc/simt_field.cu) restated in numpy float32 with the constants
    READ FROM THE SOURCE: magic-number rint(y / pi), three-term Cody-Waite reduction, odd degree-11 polynomial, sign from the
    parity of k.  Max abs error against float64 < 2e-7 for |y| <= 8000 (libm's float sine: 0.7e-7), and the three reduction
    constants must sum to pi."""
    src = open(os.path.join(ROOT, "satnerf_b200", "csrc", "simt_field.cu")).read()
    body = src[src.index("__device__ __forceinline__ float x3_sin(float y)"):]
    body = body[:body.index("template <int ACT> __device__ __forceinline__ float x3_act")]
    num = r"(-?\d+\.\d*(?:e-?\d+)?)f"
    inv_pi, magic = (np.float32(x) for x in re.search(r"fmaf\(y, %s, %s\)" % (num, num), body).groups())
    pis = [np.float32(x) for x in re.findall(r"r = fmaf\(k, %s," % num, body)]
    q = re.search(r"fmaf\(%s, r2, %s\)" % (num, num), body).groups() + tuple(re.findall(r"q = fmaf\(q, r2, %s\)" % num, body))
    coef = [np.float32(x) for x in q]                      # c11, c9, c7, c5, c3
    assert len(pis) == 3 and len(coef) == 5 and magic == np.float32(12582912.0)
    assert abs(-sum(float(p) for p in pis) - np.pi) < 1e-13

    def fma(a, b, c):
        return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)

    rng = np.random.default_rng(3)
    for lim in (4.0, 60.0, 8000.0):
        y = rng.uniform(-lim, lim, 400000).astype(np.float32)
        t = fma(y, inv_pi, np.full_like(y, magic))
        k = (t - magic).astype(np.float32)
        r = y
        for p in pis:
            r = fma(k, p, r)
        r2 = (r * r).astype(np.float32)
        acc = (np.full_like(y, coef[0]).astype(np.float64) * r2 + np.float64(coef[1])).astype(np.float32)
        for c in coef[2:]:
            acc = (acc.astype(np.float64) * r2 + np.float64(c)).astype(np.float32)
        s = ((r * r2).astype(np.float32).astype(np.float64) * acc + r).astype(np.float32)
        s = np.where((t.view(np.int32) & 1).astype(bool), -s, s)
        err = float(np.abs(s.astype(np.float64) - np.sin(y.astype(np.float64))).max())
        assert err < 2e-7, (lim, err)
