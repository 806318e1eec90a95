"""CPU restatement of the consumer right after ``predict`` -- TEST INFRASTRUCTURE, never on the product path.

``geo_eval`` follows gcnmain.py:43-63 line by line.  The distance comes from the third-party ``haversine`` package
(gcnmain.py:20; listed unpinned in requirements.txt:12, not vendored), restated here from its published formula
(haversine >= 1.0: mean earth radius 6371.0088 km).  Pinned by the package's documented known answer
Lyon (45.7597, 4.8422) - Paris (48.8567, 2.3508) = 392.2172595594006 km (tests/test_oracle.py).
"""
import math

import numpy as np

AVG_EARTH_RADIUS_KM = 6371.0088


def haversine(point1, point2):
    lat1, lng1 = point1
    lat2, lng2 = point2
    lat1, lng1, lat2, lng2 = map(math.radians, (lat1, lng1, lat2, lng2))
    lat = lat2 - lat1
    lng = lng2 - lng1
    d = math.sin(lat * 0.5) ** 2 + math.cos(lat1) * math.cos(lat2) * math.sin(lng * 0.5) ** 2
    return 2 * AVG_EARTH_RADIUS_KM * math.asin(math.sqrt(d))


def geo_eval(y_true, y_pred, U_eval, classLatMedian, classLonMedian, userLocation):
    """gcnmain.py:43-63: (mean, median, acc@161, distances, latlon_true, latlon_pred)."""
    assert len(y_pred) == len(U_eval), "#preds: %d, #users: %d" % (len(y_pred), len(U_eval))
    distances, latlon_pred, latlon_true = [], [], []
    for i in range(0, len(y_pred)):
        user = U_eval[i]
        location = userLocation[user].split(',')
        lat, lon = float(location[0]), float(location[1])
        latlon_true.append([lat, lon])
        prediction = str(y_pred[i])
        lat_pred, lon_pred = classLatMedian[prediction], classLonMedian[prediction]
        latlon_pred.append([lat_pred, lon_pred])
        distances.append(haversine((lat, lon), (lat_pred, lon_pred)))
    acc_at_161 = 100 * len([d for d in distances if d < 161]) / float(len(distances))
    return np.mean(distances), np.median(distances), acc_at_161, distances, latlon_true, latlon_pred
