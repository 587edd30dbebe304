"""ctypes mirror of include/picsp_b200_host.h: the C++ host driver (ini reader, loader, whole run)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .lib import CRunConfig, PicspError, load_library

_dp = C.POINTER(C.c_double)


def parse_ini(path, banner=False):
    """picsp_host_parse_ini -> dict of the reference's normalised globals (main.cpp:252-294)."""
    L = load_library()
    cfg = CRunConfig()
    rc = L.picsp_host_parse_ini(os.fsencode(path), C.byref(cfg), 1 if banner else 0)
    if rc != 0:
        raise PicspError(rc, "input parameters are incompatible or the file cannot be parsed")
    return cfg


def config_dict(cfg):
    return {n: getattr(cfg, n) for n, _ in CRunConfig._fields_}


def load_species(cfg, seed=0):
    """Both species through one loader (ions first, as main.cpp:437-438) -> [(x,y,vx,vy), (x,y,vx,vy)]."""
    L = load_library()
    ld = L.picsp_host_loader_create(seed)
    out = []
    try:
        for s, n in ((0, cfg.nParticlesI), (1, cfg.nParticlesE)):
            arrs = [np.empty(n) for _ in range(4)]
            rc = L.picsp_host_loader_fill(ld, C.byref(cfg), s, *(a.ctypes.data_as(_dp) for a in arrs))
            if rc != 0:
                raise PicspError(rc, "loader failed")
            out.append(tuple(arrs))
    finally:
        L.picsp_host_loader_destroy(ld)
    return out


def run(ini_path, out_path=None, max_steps=-1, quiet=True, device=0):
    rc = load_library().picsp_host_run(os.fsencode(ini_path), os.fsencode(out_path) if out_path else None,
                                       max_steps, 1 if quiet else 0, device)
    if rc != 0:
        raise PicspError(rc, "picsp_host_run failed (see stderr)")
