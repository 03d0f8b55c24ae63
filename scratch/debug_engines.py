import sys, numpy as np
sys.path.insert(0, '.')
from satpy_b200 import demo
from satpy_b200.resample import B200NearestResampler
from oracle.projections import Area as OArea
from oracle import nearest as onn
sname, ss = "goes_east_abi_f_2km", 900/5424
dname, ds = "global_eqc_500m", 1500/80150
radius = 60000.0
src, dst = demo.make_area(sname, ss), demo.make_area(dname, ds)
out = {}
for m in ("bucket", "fixed_grid"):
    r = B200NearestResampler(src, dst); r.precompute(radius_of_influence=radius, method=m)
    out[m] = r.neighbour_info()
a = out["bucket"]["index_array"][..., 0]; b = out["fixed_grid"]["index_array"][..., 0]
bad = np.argwhere(a != b)
print("mismatches", bad)
osrc, odst = OArea(*demo.area_params(sname, ss)), OArea(*demo.area_params(dname, ds))
slon, slat = osrc.get_lonlats(); dlon, dlat = odst.get_lonlats()
vin = onn.valid_lonlat(slon, slat)
sxyz = onn.lonlat2xyz(slon[vin], slat[vin])
W = src.width
raw_of = np.flatnonzero(vin.ravel())
for (r_, c_) in bad:
    q = onn.lonlat2xyz(dlon[r_, c_], dlat[r_, c_])[0]
    d = np.sqrt(((sxyz - q) ** 2).sum(-1))
    best = np.argsort(d)[:4]
    print("target", r_, c_, dlon[r_, c_], dlat[r_, c_], "bucket", a[r_, c_], "fixed", b[r_, c_])
    for k in best:
        print("   cand comp", k, "raw", raw_of[k], divmod(raw_of[k], W), "d", repr(d[k]))
    for nm, idx in (("bucket", a[r_, c_]), ("fixed", b[r_, c_])):
        if idx >= 0:
            print("  ", nm, "d =", repr(d[idx]), "raw", divmod(raw_of[idx], W))
