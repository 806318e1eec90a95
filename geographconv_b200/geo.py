"""Geolocation error of the predictions on the GPU.

``geo_eval`` has the signature and return value of gcnmain.geo_eval (gcnmain.py:43-63): mean / median error distance
in km and Acc@161 of the predicted class medians against the users' true locations.  The per-user Python loop and the
``haversine`` package call are replaced by one kernel of libgcnb200.so (csrc/geo.cu, float64); the three statistics are
NumPy reductions over the returned distances, exactly as in the reference.

``GraphConv.predict_classes`` + ``geo_eval_device`` keep the predictions on the GPU between the forward pass and the
distance kernel: nothing of size N x C (the probability matrix ``predict`` returns) crosses PCIe.
There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import logging

import numpy as np
import torch

from .layers import get_dev

AVG_EARTH_RADIUS_KM = 6371.0088  # haversine >= 1.0 (the reference lists the package unpinned)


def _f64(d, arr):
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    t = torch.empty(max(arr.size, 1), dtype=torch.float64, device=d.dev)
    if arr.size:
        d.ctx.call("gcnb_h2d", C.c_void_p(t.data_ptr()), C.c_void_p(arr.ctypes.data), arr.nbytes)
    return t, arr


def class_tables(classLatMedian, classLonMedian):
    """{'0': lat, ...} dictionaries (dump.pkl fields 11 and 12, gcnmain.py:153) -> dense float64 arrays by class id."""
    n = 1 + max(int(k) for k in classLatMedian)
    lat = np.full(n, np.nan)
    lon = np.full(n, np.nan)
    for k, v in classLatMedian.items():
        lat[int(k)] = v
    for k, v in classLonMedian.items():
        lon[int(k)] = v
    return lat, lon


def distances_device(d, preds_dev, n, class_lat, class_lon, lat_true, lon_true):
    """Device int64 predictions -> host float64 distances (km)."""
    keep = []
    t_clat, a = _f64(d, class_lat); keep.append(a)
    t_clon, a = _f64(d, class_lon); keep.append(a)
    t_lat, a = _f64(d, lat_true); keep.append(a)
    t_lon, a = _f64(d, lon_true); keep.append(a)
    dist = torch.empty(max(n, 1), dtype=torch.float64, device=d.dev)
    bad = torch.zeros(1, dtype=torch.int32, device=d.dev)
    d.fence()
    p = lambda t: C.c_void_p(t.data_ptr())
    d.ctx.call("gcnb_geo_distance_f64", p(preds_dev), int(n), p(t_clat), p(t_clon), len(class_lat), p(t_lat), p(t_lon),
               AVG_EARTH_RADIUS_KM, p(dist), p(bad))
    host = np.empty(n, dtype=np.float64)
    flag = np.zeros(1, dtype=np.int32)
    if n:
        d.ctx.call("gcnb_d2h", C.c_void_p(host.ctypes.data), p(dist), host.nbytes)
        d.ctx.call("gcnb_d2h", C.c_void_p(flag.ctypes.data), p(bad), 4)
    d.ctx.sync()
    if flag[0]:
        raise KeyError("a predicted class id has no median location")  # the reference's dict lookup raises KeyError
    return host


def geo_eval(y_true, y_pred, U_eval, classLatMedian, classLonMedian, userLocation, device=None, preds_device=None):
    """Drop-in for gcnmain.geo_eval (gcnmain.py:43-63).  ``preds_device``: optional device int64 tensor holding
    ``y_pred`` already (from ``GraphConv.predict_classes``), which skips the upload."""
    assert len(y_pred) == len(U_eval), "#preds: %d, #users: %d" % (len(y_pred), len(U_eval))
    d = get_dev(device)
    n = len(U_eval)
    loc = np.array([userLocation[u].split(',')[:2] for u in U_eval], dtype=np.float64).reshape(n, 2)
    class_lat, class_lon = class_tables(classLatMedian, classLonMedian)
    y_pred = np.ascontiguousarray(y_pred, dtype=np.int64)
    if preds_device is None:
        preds_device = torch.empty(max(n, 1), dtype=torch.int64, device=d.dev)
        if n:
            d.ctx.call("gcnb_h2d", C.c_void_p(preds_device.data_ptr()), C.c_void_p(y_pred.ctypes.data), y_pred.nbytes)
    dist = distances_device(d, preds_device, n, class_lat, class_lon, loc[:, 0], loc[:, 1])
    distances = dist.tolist()
    acc_at_161 = 100 * int(np.count_nonzero(dist < 161)) / float(len(distances))
    logging.info("Mean: " + str(int(np.mean(distances))) + " Median: " + str(int(np.median(distances))) + " Acc@161: "
                 + str(int(acc_at_161)))
    latlon_true = loc.tolist()
    latlon_pred = np.stack([class_lat[y_pred], class_lon[y_pred]], axis=1).tolist() if n else []
    return np.mean(distances), np.median(distances), acc_at_161, distances, latlon_true, latlon_pred
