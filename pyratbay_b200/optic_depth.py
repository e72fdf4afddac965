"""Optical depth from the extinction coefficient on the GPU, mirroring
pyratbay/opacity/optic_depth.py:16-146 (next-tier row)."""
import ctypes

import numpy as np

from . import _lib
from ._lib import check

transmission_rt = ['transit']
eclipse_rt = ['eclipse', 'eclipse_two_stream']
emission_rt = ['emission', 'emission_two_stream', 'f_lambda']   # code_constants.py:84-99


def transit_path(radius, nskip=0):
    """Distances travelled through each shell by rays grazing the atmosphere at the layer
    radii (atmosphere/atmosphere.py:737-799).  Returns a list of 1D arrays."""
    rad = np.asarray(radius, np.double)[nskip:]
    nlayers = len(rad)
    path = [np.empty(0, np.double) for _ in range(nskip)]
    for r in range(nlayers):
        raypath = np.empty(r, np.double)
        for i in range(r):
            raypath[i] = (np.sqrt(rad[i]**2 - rad[r]**2) - np.sqrt(rad[i+1]**2 - rad[r]**2))
        path.append(raypath)
    return path


def _depth(transit, ec, geometry, maxdepth, itop, ibottom, device):
    lib = _lib.load()
    _lib.require_device()
    nlayers, nwave = ec.shape
    depth = np.zeros((nlayers, nwave), np.double)
    ideep = np.zeros(nwave, np.int32)
    ec = np.ascontiguousarray(ec, np.double)
    geometry = np.ascontiguousarray(geometry, np.double)
    check(lib.pb200_optical_depth(
        ctypes.c_int(device), ctypes.c_int(int(transit)),
        depth.ctypes.data_as(ctypes.c_void_p), ideep.ctypes.data_as(ctypes.c_void_p),
        ec.ctypes.data_as(ctypes.c_void_p), geometry.ctypes.data_as(ctypes.c_void_p),
        ctypes.c_double(maxdepth), ctypes.c_int(int(itop)), ctypes.c_int(int(ibottom)),
        ctypes.c_int(nlayers), ctypes.c_int(nwave)))
    return depth, ideep


def _path_matrix(raypath, nlayers):
    mat = np.zeros((nlayers, nlayers), np.double)
    for r, row in enumerate(raypath):
        mat[r, :len(row)] = row
    return mat


def optical_depth(rt_path, extinction, radius=None, itop=0, ibottom=None, maxdepth=np.inf,
                  extinction_cloudy=None, raypath=None, device=0):
    """Same signature and return tuple as the reference's optical_depth:
    (raypath, depth, ideep, depth_clear, ideep_clear)."""
    is_patchy = extinction_cloudy is not None
    extinction = np.asarray(extinction, np.double)
    nlayers, nwave = extinction.shape
    if rt_path in transmission_rt:
        is_transit = True
    elif rt_path in emission_rt + eclipse_rt:
        is_transit = False
    else:
        raise ValueError('Invalid radiative-transfer path')
    if ibottom is None:
        ibottom = nlayers
    if radius is None and raypath is None:
        raise ValueError(
            'Need to provide either radius or raypath to compute the '
            'path for the optical-depth calculation')
    if raypath is None and is_transit:
        raypath = transit_path(radius, itop)
    if raypath is None and not is_transit:
        raypath = -np.ediff1d(np.asarray(radius, np.double))     # -cu.ediff(radius)

    ec = np.copy(extinction)
    depth_clear = ideep_clear = None
    if is_patchy:
        ec_clear = np.copy(extinction)
        ec[itop:] += np.asarray(extinction_cloudy)[itop:]

    if is_transit:
        paths = _path_matrix(raypath, nlayers)
        depth, ideep = _depth(True, ec, paths, maxdepth, itop, ibottom, device)
        ideep = ideep.astype(np.intc)
        if is_patchy:
            depth_clear, ideep_clear = _depth(True, ec_clear, paths, maxdepth, itop, nlayers,
                                              device)
            ideep_clear = ideep_clear.astype(np.intc)
    else:
        if 'two_stream' in rt_path:
            maxdepth = np.inf
        depth, ideep = _depth(False, ec, raypath, maxdepth, itop, ibottom, device)
        ideep = ideep.astype(int)
        if is_patchy:
            depth_clear, ideep_clear = _depth(False, ec_clear, raypath, maxdepth, itop, nlayers,
                                              device)
            ideep_clear = ideep_clear.astype(int)
    return raypath, depth, ideep, depth_clear, ideep_clear
