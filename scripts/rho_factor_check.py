import numpy as np
a = 6378137.0; b = 6356752.314; h = 35786023.0; D = a + h
def rho(theta, psi, far=False):
    # ray from S=(D,0,0) dir u; ellipsoid scaled: z' = z*a/b
    u = np.array([-np.cos(theta), np.sin(theta)*np.cos(psi), np.sin(theta)*np.sin(psi)])
    k = a/b
    S = np.array([D, 0, 0.0])
    us = u*np.array([1,1,k])[:,None] if u.ndim>1 else u*np.array([1,1,k])
    A = (us*us).sum(0); B = 2*(S[0]*us[0]); C = D*D - a*a
    disc = B*B - 4*A*C
    ok = disc >= 0
    sq = np.sqrt(np.where(ok, disc, 0))
    t = (-B - sq)/(2*A) if not far else (-B + sq)/(2*A)
    return np.where(ok, t, np.nan)
def kfac(r, c1, cap):
    s = (r/h)**2 - 1.0
    return np.minimum(cap, 1.0 + c1*s)
for px in (14e-6, 56e-6, 112e-6, 250e-6):
  for c1, cap in ((0.36, 1.10), (0.40,1.12), (0.30, 1.10)):
    worst = np.inf
    delta = 3.0*px
    for psi in np.linspace(0, np.pi/2, 7):
        th = np.linspace(0, 0.1525, 400001)
        r_near = rho(th, np.full_like(th, psi))
        r_far = rho(th, np.full_like(th, psi), far=True)
        vis = np.isfinite(r_near)
        # candidates: min rho over directions within delta: sample inward theta-delta (same psi) and sideways
        rmin = np.full_like(th, np.inf)
        for dth, dpsi_fac in ((-delta, 0.0), (-delta*0.7, 0.7), (-delta*0.7, -0.7), (0, 1.0), (0, -1.0)):
            th2 = np.abs(th + dth)
            with np.errstate(divide='ignore', invalid='ignore'):
                psi2 = psi + np.where(th > 0, dpsi_fac*delta/np.maximum(th, 1e-9), 0.0)
            r2 = rho(th2, psi2)
            rmin = np.where(np.isfinite(r2), np.minimum(rmin, r2), rmin)
        for rT, extra in ((r_near, vis), (r_far, vis & ((r_far - r_near) < 300e3))):
            k = kfac(rT, c1, cap)
            m = extra & np.isfinite(rmin) & np.isfinite(rT)
            ratio = (rmin[m]/h)/k[m]
            worst = min(worst, ratio.min())
    print(f"px {px*1e6:.0f} urad c1 {c1} cap {cap}: min rho_cand/(h k) = {worst:.6f}")
