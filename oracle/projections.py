"""Oracle map projections (numpy float64) -- TEST INFRASTRUCTURE ONLY.

The reference gets every lon/lat <-> x/y transform from pyproj/PROJ through
``pyresample.geometry.AreaDefinition.get_lonlats`` (call sites:
satpy/modifiers/angles.py:434-441, satpy/resample/kdtree.py:79-87).  Neither
pyproj nor pyresample is in /root/reference or installed here, so this module
restates the published PROJ algorithms (Snyder 1987, "Map Projections -- A
Working Manual"; PROJ ``geos.cpp`` for the geostationary view) for the
projections the benchmark configs need.

Anchors (tests/test_oracle_golden.py):
  * ``merc`` inverse: through the SunZenithCorrector golden vectors
    (satpy/tests/test_modifiers.py:34-110,116-176);
  * ``geos`` inverse: against the reference's own formula
    ``_lonlat_from_geos_angle`` (satpy/readers/core/utils.py:136-157);
  * ``lcc``/``laea``/``stere``/``merc`` forward+inverse: Snyder's numerical
    examples (known answers);
  * every projection: forward(inverse(x)) == x round trips.

Conventions: lon/lat in degrees, x/y in metres (degrees for ``longlat``).
Points that cannot be projected come back as ``inf`` like pyproj's default.
"""
from __future__ import annotations

import math

import numpy as np

ELLIPSOIDS = {
    # name: (a, rf)  (rf = 0 => sphere); values as in PROJ's ellps table
    "WGS84": (6378137.0, 298.257223563),
    "GRS80": (6378137.0, 298.257222101),
    "WGS72": (6378135.0, 298.26),
    "intl": (6378388.0, 297.0),
    "bessel": (6377397.155, 299.1528128),
    "clrk66": (6378206.4, 294.9786982138982),  # b = 6356583.8
    "sphere": (6370997.0, 0.0),
}

_EPS10 = 1e-10
_HALFPI = math.pi / 2


def ellipsoid_from_dict(proj: dict) -> tuple[float, float]:
    """Return ``(a, es)`` following PROJ's parameter precedence (R > a/b/rf/f > ellps > datum)."""
    if "R" in proj:
        return float(proj["R"]), 0.0
    a = None
    es = None
    name = proj.get("ellps")
    if name is None and proj.get("datum") is not None:
        name = {"WGS84": "WGS84", "NAD83": "GRS80", "NAD27": "clrk66"}[proj["datum"]]
    if name is None and "a" not in proj:
        name = "WGS84"
    if name is not None:
        a0, rf = ELLIPSOIDS[name]
        a = a0
        es = 0.0 if rf == 0 else (2.0 / rf - 1.0 / rf ** 2)
    if "a" in proj:
        a = float(proj["a"])
        if es is None:
            es = 0.0
    if "b" in proj:
        b = float(proj["b"])
        es = 1.0 - (b * b) / (a * a)
    elif "rf" in proj:
        rf = float(proj["rf"])
        es = 2.0 / rf - 1.0 / rf ** 2
    elif "f" in proj:
        f = float(proj["f"])
        es = 2.0 * f - f * f
    elif "es" in proj:
        es = float(proj["es"])
    return a, es


def _tsfn(phi, sinphi, e):
    """Snyder (15-9): t = tan(pi/4 - phi/2) / ((1 - e sin phi)/(1 + e sin phi))^(e/2)."""
    con = e * sinphi
    return np.tan(0.5 * (_HALFPI - phi)) / np.power((1.0 - con) / (1.0 + con), 0.5 * e)


def _msfn(sinphi, cosphi, es):
    """Snyder (14-15): m = cos phi / sqrt(1 - e^2 sin^2 phi)."""
    return cosphi / np.sqrt(1.0 - es * sinphi * sinphi)


def _phi2(ts, e):
    """Snyder (7-9): iterate the conformal-latitude inverse (PROJ ``pj_phi2``, tol 1e-10, 15 iterations)."""
    ts = np.asarray(ts, dtype=np.float64)
    eccnth = 0.5 * e
    phi = _HALFPI - 2.0 * np.arctan(ts)
    for _ in range(15):
        con = e * np.sin(phi)
        dphi = _HALFPI - 2.0 * np.arctan(ts * np.power((1.0 - con) / (1.0 + con), eccnth)) - phi
        phi = phi + dphi
        if np.all(np.abs(dphi) <= _EPS10):
            break
    return phi


def _qsfn(sinphi, e, one_es):
    """Snyder (3-12): authalic q."""
    if e < 1e-7:
        return 2.0 * sinphi
    con = e * sinphi
    return one_es * (sinphi / (1.0 - con * con) - (0.5 / e) * np.log((1.0 - con) / (1.0 + con)))


def _authlat(beta, es):
    """Snyder (3-18): geodetic latitude from authalic latitude (PROJ ``pj_authlat`` series)."""
    es2 = es * es
    es3 = es2 * es
    apa0 = es / 3.0 + es2 * (31.0 / 180.0) + es3 * (517.0 / 5040.0)
    apa1 = es2 * (23.0 / 360.0) + es3 * (251.0 / 3780.0)
    apa2 = es3 * (761.0 / 45360.0)
    t = beta + beta
    return beta + apa0 * np.sin(t) + apa1 * np.sin(t + t) + apa2 * np.sin(t + t + t)


def _wrap_lon_rad(lam):
    """PROJ ``adjlon``: wrap to [-pi, pi] leaving values already inside untouched."""
    lam = np.asarray(lam, dtype=np.float64)
    out = lam.copy()
    big = np.abs(lam) > math.pi + 1e-12
    if np.any(big):
        w = lam[big] + math.pi
        w = w - 2 * math.pi * np.floor(w / (2 * math.pi))
        out[big] = w - math.pi
    return out


class Proj:
    """One projection instance built from a PROJ-style parameter dict."""

    def __init__(self, proj: dict):
        self.params = dict(proj)
        self.name = proj["proj"]
        if self.name in ("latlong", "longlat", "latlon", "lonlat"):
            self.name = "longlat"
        self.a, self.es = ellipsoid_from_dict(proj)
        self.e = math.sqrt(self.es)
        self.one_es = 1.0 - self.es
        self.lam0 = math.radians(float(proj.get("lon_0", 0.0)))
        self.phi0 = math.radians(float(proj.get("lat_0", 0.0)))
        self.x0 = float(proj.get("x_0", 0.0))
        self.y0 = float(proj.get("y_0", 0.0))
        self.k0 = float(proj.get("k_0", proj.get("k", 1.0)))
        getattr(self, "_setup_" + self.name)()

    # ------------------------------------------------------------------ API
    def forward(self, lon, lat):
        lon = np.asarray(lon, dtype=np.float64)
        lat = np.asarray(lat, dtype=np.float64)
        if self.name == "longlat":
            return lon.copy(), lat.copy()
        lam = _wrap_lon_rad(np.radians(lon) - self.lam0)
        phi = np.radians(lat)
        with np.errstate(all="ignore"):
            x, y = getattr(self, "_fwd_" + self.name)(lam, phi)
        bad = ~(np.isfinite(x) & np.isfinite(y)) | (np.abs(lat) > 90.0 + 1e-12)
        x = np.where(bad, np.inf, self.a * x + self.x0)
        y = np.where(bad, np.inf, self.a * y + self.y0)
        return x, y

    def inverse(self, x, y):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if self.name == "longlat":
            return x.copy(), y.copy()
        xn = (x - self.x0) / self.a
        yn = (y - self.y0) / self.a
        with np.errstate(all="ignore"):
            lam, phi = getattr(self, "_inv_" + self.name)(xn, yn)
        bad = ~(np.isfinite(lam) & np.isfinite(phi))
        lam = _wrap_lon_rad(np.where(bad, 0.0, lam) + self.lam0)
        lon = np.where(bad, np.inf, np.degrees(lam))
        lat = np.where(bad, np.inf, np.degrees(np.where(bad, 0.0, phi)))
        return lon, lat

    # -------------------------------------------------------------- longlat
    def _setup_longlat(self):
        pass

    # ------------------------------------------------------------------ eqc
    # PROJ eqc.cpp: spherical only (es forced to 0); x = lam*cos(lat_ts), y = phi - phi0.
    def _setup_eqc(self):
        self.rc = math.cos(math.radians(float(self.params.get("lat_ts", 0.0))))

    def _fwd_eqc(self, lam, phi):
        return self.rc * lam, phi - self.phi0

    def _inv_eqc(self, x, y):
        return x / self.rc, y + self.phi0

    # ----------------------------------------------------------------- merc
    # Snyder ch. 7 (7-6..7-9); PROJ merc.cpp.
    def _setup_merc(self):
        if "lat_ts" in self.params:
            ts = math.radians(float(self.params["lat_ts"]))
            if self.es != 0.0:
                self.k0 = _msfn(math.sin(ts), math.cos(ts), self.es)
            else:
                self.k0 = math.cos(ts)

    def _fwd_merc(self, lam, phi):
        if self.es != 0.0:
            return self.k0 * lam, -self.k0 * np.log(_tsfn(phi, np.sin(phi), self.e))
        return self.k0 * lam, self.k0 * np.log(np.tan(math.pi / 4 + 0.5 * phi))

    def _inv_merc(self, x, y):
        if self.es != 0.0:
            return x / self.k0, _phi2(np.exp(-y / self.k0), self.e)
        return x / self.k0, _HALFPI - 2.0 * np.arctan(np.exp(-y / self.k0))

    # ------------------------------------------------------------------ lcc
    # Snyder ch. 15 (15-1..15-11); PROJ lcc.cpp (lat_2 defaults to lat_1, lat_0 defaults to lat_1).
    def _setup_lcc(self):
        p = self.params
        phi1 = math.radians(float(p["lat_1"]))
        phi2 = math.radians(float(p.get("lat_2", p["lat_1"])))
        if "lat_0" not in p:
            self.phi0 = phi1
        sinphi = math.sin(phi1)
        cosphi = math.cos(phi1)
        secant = abs(phi1 - phi2) >= _EPS10
        n = sinphi
        if self.es != 0.0:
            m1 = _msfn(sinphi, cosphi, self.es)
            ml1 = float(_tsfn(phi1, sinphi, self.e))
            if secant:
                sinphi2 = math.sin(phi2)
                n = math.log(m1 / _msfn(sinphi2, math.cos(phi2), self.es))
                n /= math.log(ml1 / float(_tsfn(phi2, sinphi2, self.e)))
            self.c = m1 * math.pow(ml1, -n) / n
            self.rho0 = 0.0 if abs(abs(self.phi0) - _HALFPI) < _EPS10 else \
                self.c * math.pow(float(_tsfn(self.phi0, math.sin(self.phi0), self.e)), n)
        else:
            if secant:
                n = math.log(cosphi / math.cos(phi2)) / math.log(
                    math.tan(math.pi / 4 + 0.5 * phi2) / math.tan(math.pi / 4 + 0.5 * phi1))
            self.c = cosphi * math.pow(math.tan(math.pi / 4 + 0.5 * phi1), n) / n
            self.rho0 = 0.0 if abs(abs(self.phi0) - _HALFPI) < _EPS10 else \
                self.c * math.pow(math.tan(math.pi / 4 + 0.5 * self.phi0), -n)
        self.n = n

    def _fwd_lcc(self, lam, phi):
        if self.es != 0.0:
            rho = self.c * np.power(_tsfn(phi, np.sin(phi), self.e), self.n)
        else:
            rho = self.c * np.power(np.tan(math.pi / 4 + 0.5 * phi), -self.n)
        lam = lam * self.n
        return self.k0 * (rho * np.sin(lam)), self.k0 * (self.rho0 - rho * np.cos(lam))

    def _inv_lcc(self, x, y):
        x = x / self.k0
        y = self.rho0 - y / self.k0
        rho = np.hypot(x, y)
        if self.n < 0.0:
            rho = -rho
            x = -x
            y = -y
        if self.es != 0.0:
            phi = _phi2(np.power(rho / self.c, 1.0 / self.n), self.e)
        else:
            phi = 2.0 * np.arctan(np.power(self.c / rho, 1.0 / self.n)) - _HALFPI
        lam = np.arctan2(x, y) / self.n
        return lam, phi

    # ---------------------------------------------------------------- stere
    # Snyder ch. 21 polar aspect (21-33..21-40); PROJ stere.cpp (polar modes only here).
    def _setup_stere(self):
        p = self.params
        if abs(abs(self.phi0) - _HALFPI) >= _EPS10:
            raise NotImplementedError("oracle: only polar stereographic (lat_0 = +-90) is restated")
        self.north = self.phi0 > 0
        phits = abs(math.radians(float(p["lat_ts"]))) if "lat_ts" in p else _HALFPI
        if self.es != 0.0:
            if abs(phits - _HALFPI) < _EPS10:
                self.akm1 = 2.0 * self.k0 / math.sqrt(
                    math.pow(1 + self.e, 1 + self.e) * math.pow(1 - self.e, 1 - self.e))
            else:
                s = math.sin(phits)
                self.akm1 = math.cos(phits) / float(_tsfn(phits, s, self.e))
                self.akm1 /= math.sqrt(1.0 - self.es * s * s)
        else:
            self.akm1 = (math.cos(phits) / math.tan(math.pi / 4 - 0.5 * phits)
                         if abs(phits - _HALFPI) >= _EPS10 else 2.0 * self.k0)

    def _fwd_stere(self, lam, phi):
        coslam = np.cos(lam)
        sinlam = np.sin(lam)
        if not self.north:
            phi = -phi
            coslam = -coslam
        if self.es != 0.0:
            rho = self.akm1 * _tsfn(phi, np.sin(phi), self.e)
        else:
            rho = self.akm1 * np.tan(math.pi / 4 - 0.5 * phi)
        return rho * sinlam, -rho * coslam

    def _inv_stere(self, x, y):
        rho = np.hypot(x, y)
        if self.north:
            y = -y
        if self.es != 0.0:
            phi = _phi2(rho / self.akm1, self.e)
        else:
            phi = _HALFPI - 2.0 * np.arctan(rho / self.akm1)
        if not self.north:
            phi = -phi
        lam = np.where(rho == 0.0, 0.0, np.arctan2(x, y))
        return lam, phi

    # ----------------------------------------------------------------- laea
    # Snyder ch. 24 (24-17..24-26 ellipsoid, 24-2..24-4 sphere); PROJ laea.cpp.
    def _setup_laea(self):
        t = abs(self.phi0)
        if abs(t - _HALFPI) < _EPS10:
            self.mode = "N" if self.phi0 > 0 else "S"
        elif t < _EPS10:
            self.mode = "EQ"
        else:
            self.mode = "OB"
        if self.es != 0.0:
            self.qp = float(_qsfn(1.0, self.e, self.one_es))
            self.rq = math.sqrt(0.5 * self.qp)
            sinphi = math.sin(self.phi0)
            self.sinb1 = float(_qsfn(sinphi, self.e, self.one_es)) / self.qp
            self.cosb1 = math.sqrt(max(0.0, 1.0 - self.sinb1 * self.sinb1))
            if self.mode == "OB":
                self.dd = math.cos(self.phi0) / (math.sqrt(1.0 - self.es * sinphi * sinphi) * self.rq * self.cosb1)
            else:
                self.dd = 1.0 / self.rq if self.mode == "EQ" else 1.0
            self.xmf = self.rq * self.dd
            self.ymf = self.rq / self.dd
        else:
            self.sinb1 = math.sin(self.phi0)
            self.cosb1 = math.cos(self.phi0)

    def _fwd_laea(self, lam, phi):
        coslam = np.cos(lam)
        sinlam = np.sin(lam)
        sinphi = np.sin(phi)
        if self.es != 0.0:
            q = _qsfn(sinphi, self.e, self.one_es)
            if self.mode in ("OB", "EQ"):
                sinb = q / self.qp
                cosb = np.sqrt(np.maximum(0.0, 1.0 - sinb * sinb))
                if self.mode == "OB":
                    b = 1.0 + self.sinb1 * sinb + self.cosb1 * cosb * coslam
                    b = np.sqrt(2.0 / b)
                    x = self.xmf * b * cosb * sinlam
                    y = self.ymf * b * (self.cosb1 * sinb - self.sinb1 * cosb * coslam)
                else:
                    b = np.sqrt(2.0 / (1.0 + cosb * coslam))
                    x = self.xmf * b * cosb * sinlam
                    y = b * sinb * self.ymf
                return x, y
            if self.mode == "N":
                qq = self.qp - q
                b = np.sqrt(np.maximum(qq, 0.0))
                return b * sinlam, -b * coslam
            qq = self.qp + q
            b = np.sqrt(np.maximum(qq, 0.0))
            return b * sinlam, b * coslam
        cosphi = np.cos(phi)
        if self.mode in ("OB", "EQ"):
            kk = 1.0 + self.sinb1 * sinphi + self.cosb1 * cosphi * coslam
            kk = np.where(kk <= _EPS10, np.nan, kk)
            kk = np.sqrt(2.0 / kk)
            return kk * cosphi * sinlam, kk * (self.cosb1 * sinphi - self.sinb1 * cosphi * coslam)
        if self.mode == "N":
            yk = 2.0 * np.sin(math.pi / 4 - 0.5 * phi)
            return yk * sinlam, -yk * coslam
        yk = 2.0 * np.cos(math.pi / 4 - 0.5 * phi)
        return yk * sinlam, yk * coslam

    def _inv_laea(self, x, y):
        if self.es != 0.0:
            if self.mode in ("OB", "EQ"):
                x = x / self.dd
                y = y * self.dd
                rho = np.hypot(x, y)
                sce = 2.0 * np.arcsin(0.5 * rho / self.rq)
                cce = np.cos(sce)
                sce = np.sin(sce)
                xx = x * sce
                safe_rho = np.where(rho < _EPS10, 1.0, rho)
                if self.mode == "OB":
                    ab = cce * self.sinb1 + y * sce * self.cosb1 / safe_rho
                    yy = safe_rho * self.cosb1 * cce - y * self.sinb1 * sce
                else:
                    ab = y * sce / safe_rho
                    yy = safe_rho * cce
                lam = np.arctan2(xx, yy)
                phi = _authlat(np.arcsin(ab), self.es)
                lam = np.where(rho < _EPS10, 0.0, lam)
                phi = np.where(rho < _EPS10, self.phi0, phi)
                phi = np.where(0.5 * rho / self.rq > 1.0, np.nan, phi)
                return lam, phi
            if self.mode == "N":
                y = -y
            q = x * x + y * y
            ab = 1.0 - q / self.qp
            if self.mode == "S":
                ab = -ab
            lam = np.arctan2(x, y)
            phi = _authlat(np.arcsin(ab), self.es)
            return np.where(q == 0.0, 0.0, lam), np.where(q == 0.0, self.phi0, phi)
        rh = np.hypot(x, y)
        phi = 0.5 * rh
        phi = np.where(phi > 1.0, np.nan, phi)
        phi = 2.0 * np.arcsin(phi)
        sinz = np.sin(phi)
        cosz = np.cos(phi)
        if self.mode == "EQ":
            ph = np.where(rh <= _EPS10, 0.0, np.arcsin(y * sinz / np.where(rh <= _EPS10, 1.0, rh)))
            xx = x * sinz
            yy = cosz * rh
        elif self.mode == "OB":
            safe = np.where(rh <= _EPS10, 1.0, rh)
            ph = np.where(rh <= _EPS10, self.phi0,
                          np.arcsin(cosz * self.sinb1 + y * sinz * self.cosb1 / safe))
            xx = x * sinz * self.cosb1
            yy = (cosz - np.sin(ph) * self.sinb1) * rh
        elif self.mode == "N":
            yy = -y
            xx = x
            ph = _HALFPI - phi
        else:
            yy = y
            xx = x
            ph = phi - _HALFPI
        lam = np.where((yy == 0.0) & (self.mode in ("EQ", "OB")), 0.0, np.arctan2(xx, yy))
        return lam, ph

    # ----------------------------------------------------------------- geos
    # PROJ geos.cpp (ellipsoidal + spherical share this form with radius_p = 1 on a sphere).
    # Cross-checked against satpy/readers/core/utils.py:136-157 for sweep='y'.
    def _setup_geos(self):
        h = float(self.params["h"])
        self.flip_axis = str(self.params.get("sweep", "y")) == "x"
        self.radius_g_1 = h / self.a
        self.radius_g = 1.0 + self.radius_g_1
        self.c_geos = self.radius_g * self.radius_g - 1.0
        self.radius_p = math.sqrt(self.one_es)
        self.radius_p2 = self.one_es
        self.radius_p_inv2 = 1.0 / self.one_es

    def _fwd_geos(self, lam, phi):
        phi = np.arctan(self.radius_p2 * np.tan(phi))
        r = self.radius_p / np.hypot(self.radius_p * np.cos(phi), np.sin(phi))
        vx = r * np.cos(lam) * np.cos(phi)
        vy = r * np.sin(lam) * np.cos(phi)
        vz = r * np.sin(phi)
        hidden = ((self.radius_g - vx) * vx - vy * vy - vz * vz * self.radius_p_inv2) < 0.0
        tmp = self.radius_g - vx
        if self.flip_axis:
            x = self.radius_g_1 * np.arctan(vy / np.hypot(vz, tmp))
            y = self.radius_g_1 * np.arctan(vz / tmp)
        else:
            x = self.radius_g_1 * np.arctan(vy / tmp)
            y = self.radius_g_1 * np.arctan(vz / np.hypot(vy, tmp))
        return np.where(hidden, np.nan, x), np.where(hidden, np.nan, y)

    def _inv_geos(self, x, y):
        vx = -1.0
        if self.flip_axis:
            vz = np.tan(y / self.radius_g_1)
            vy = np.tan(x / self.radius_g_1) * np.hypot(1.0, vz)
        else:
            vy = np.tan(x / self.radius_g_1)
            vz = np.tan(y / self.radius_g_1) * np.hypot(1.0, vy)
        a = vz / self.radius_p
        a = vy * vy + a * a + vx * vx
        b = 2.0 * self.radius_g * vx
        det = b * b - 4.0 * a * self.c_geos
        k = (-b - np.sqrt(det)) / (2.0 * a)
        vx = self.radius_g + k * vx
        vy = vy * k
        vz = vz * k
        lam = np.arctan2(vy, vx)
        phi = np.arctan(vz * np.cos(lam) / vx)
        phi = np.arctan(self.radius_p_inv2 * np.tan(phi))
        bad = det < 0.0
        return np.where(bad, np.nan, lam), np.where(bad, np.nan, phi)


class Area:
    """Minimal restatement of ``pyresample.geometry.AreaDefinition`` pixel-centre geometry.

    ``area_extent = (x_ll, y_ll, x_ur, y_ur)`` are the OUTER pixel edges
    (satpy/tests/test_resample.py:66-73, satpy/readers/core/abi.py:271-291);
    row 0 is the northernmost row.
    """

    def __init__(self, proj: dict, width: int, height: int, area_extent):
        self.proj_dict = dict(proj)
        self.proj = Proj(proj)
        self.width = int(width)
        self.height = int(height)
        self.area_extent = tuple(float(v) for v in area_extent)
        self.pixel_size_x = (self.area_extent[2] - self.area_extent[0]) / self.width
        self.pixel_size_y = (self.area_extent[3] - self.area_extent[1]) / self.height
        self.shape = (self.height, self.width)

    def proj_vectors(self, rows=None, cols=None):
        if cols is None:
            cols = np.arange(self.width)
        if rows is None:
            rows = np.arange(self.height)
        # pyresample ``AreaDefinition._get_proj_vectors``: arange(n) * pixel_size + pixel_upper_left
        x_ul = self.area_extent[0] + self.pixel_size_x / 2.0
        y_ul = self.area_extent[3] - self.pixel_size_y / 2.0
        x = np.asarray(cols, dtype=np.float64) * self.pixel_size_x + x_ul
        y = np.asarray(rows, dtype=np.float64) * -self.pixel_size_y + y_ul
        return x, y

    def get_lonlats(self, row_slice: slice | None = None):
        """lon/lat (deg, float64) of pixel centres; ``inf`` where the projection has no answer."""
        rows = np.arange(self.height)[row_slice] if row_slice is not None else None
        x, y = self.proj_vectors(rows=rows)
        xx, yy = np.meshgrid(x, y)
        return self.proj.inverse(xx, yy)


class Swath:
    """Minimal ``SwathDefinition``: explicit lon/lat arrays (degrees)."""

    def __init__(self, lons, lats):
        self.lons = np.asarray(lons)
        self.lats = np.asarray(lats)
        self.shape = self.lons.shape
        self.height, self.width = self.shape

    def get_lonlats(self, row_slice: slice | None = None):
        if row_slice is not None:
            return (np.asarray(self.lons[row_slice], dtype=np.float64),
                    np.asarray(self.lats[row_slice], dtype=np.float64))
        return np.asarray(self.lons, dtype=np.float64), np.asarray(self.lats, dtype=np.float64)
