"""satnerf_b200 — B200-native (sm_100a) implementation of Sat-NeRF's volumetric-rendering hot path.

Drop-in surface (same names and signatures as the reference, centreborelli/satnerf):
    from satnerf_b200.rendering import render_rays, sample_pdf, batched_inference
    from satnerf_b200.models import load_model, SatNeRF, ShadowNeRF, NeRF
    from satnerf_b200.geo import get_rays, SatelliteGeometry          (datasets/satellite.py ray generation / DSM helpers)
    from satnerf_b200.hostio import HostPipeline                      (host-buffer front end: copy-out overlapped with the next batch)
"""
from .models import NeRF, SatNeRF, ShadowNeRF, load_model  # noqa: F401
from .rendering import batched_inference, inference, render_rays, sample_pdf  # noqa: F401

__all__ = ["render_rays", "sample_pdf", "inference", "batched_inference", "load_model", "SatNeRF", "ShadowNeRF", "NeRF"]
