// K2 (Love) — eigenfunctions, energy integrals, group velocity and phase-velocity Frechet kernels
// for one (model, period, c), one thread each.
//
// Replaces /root/reference/src/SWD/slegn96.f90: shfunc :179-246, varl :248-328, hskl :330-351,
// emat :353-370, up :372-445, energy :447-629 and the suffix sum of slegn96 :759-766.
//
// Re-design: the up-sweep keeps (uu,tt,exl) per layer in thread-local memory; the top-down
// rescale of shfunc, the energy integrals and the dc/dh boundary terms are fused in one pass.
// kern layout: [4][n] with stride ks, order dcda(=0), dcdb, dcdr, dcdh.
#pragma once
#include "common.cuh"
#include "swd_roots.cuh"

namespace rfs {

struct VarL {
  double rb, xkb, cosq, yl, zl, eexl;
};
RFS_DEVINL VarL varl_dev(double zb, double omega, double wvno, double dpth) {
  VarL o;
  o.xkb = omega / zb;
  o.rb = sqrt((wvno + o.xkb) * fabs(wvno - o.xkb));
  const double q = o.rb * dpth;
  o.eexl = 0.0;
  if (wvno < o.xkb) {
    double sinq;
    sincos_cb(q, &sinq, &o.cosq);
    o.yl = sinq / o.rb;
    o.zl = -o.rb * sinq;
  } else if (wvno == o.xkb) {
    o.cosq = 1.0;
    o.yl = dpth;
    o.zl = 0.0;
  } else {
    o.eexl = q;
    double fac = 0.0;
    if (q < 18.0) fac = exp_neg(2.0 * q);
    o.cosq = (1.0 + fac) * 0.5;
    const double sinq = (1.0 - fac) * 0.5;
    o.yl = sinq / o.rb;
    o.zl = o.rb * sinq;
  }
  return o;
}

template <int NMAX>
RFS_DEVINL void love_solve(const SwdModel &M, long long b, double T, double c, double *ugr_out,
                           double *__restrict__ kern, long long ks) {
  const int mmax = M.n;
  const double omega = (2.0 * RFS_PI32) / T;
  const double wvno = omega / c;
  const double omega2 = omega * omega, wvno2 = wvno * wvno;
  double ul[NMAX * 3];  // [m][0]=uu [1]=tt [2]=exl

  // ---------------- up (:372-445)
  {
    const int m = mmax - 1;
    const double zb = M.ld(F_B, m, b), zr = M.ld(F_RHO, m, b);
    ul[m * 3 + 0] = 1.0;
    if (zb > 0.01) {
      const VarL v = varl_dev(zb, omega, wvno, 0.0);
      ul[m * 3 + 1] = -(zr * zb * zb) * v.rb;
    } else {
      ul[m * 3 + 1] = 0.0;
    }
    ul[m * 3 + 2] = 0.0;
  }
  // fluid layers carry no SH motion: they are skipped everywhere (iwat tests of up/shfunc/energy);
  // only a contiguous block of water layers at the top is supported (k0 = first solid layer)
  int k0 = 0;
  while (k0 < mmax - 1 && !(M.ld(F_B, k0, b) > 0.0)) {
    ul[k0 * 3 + 0] = 0.0;
    ul[k0 * 3 + 1] = 0.0;
    ul[k0 * 3 + 2] = 0.0;
    k0++;
  }
  for (int k = mmax - 2; k >= k0; k--) {
    const double zb = M.ld(F_B, k, b), zr = M.ld(F_RHO, k, b), dpth = M.ld(F_D, k, b);
    const VarL v = varl_dev(zb, omega, wvno, dpth);
    const double mu = zr * zb * zb;
    const double a11 = v.cosq, a22 = v.cosq, a12 = -(v.yl / mu), a21 = -(v.zl * mu);
    const double u1 = ul[(k + 1) * 3 + 0], t1 = ul[(k + 1) * 3 + 1];
    const double amp0 = a11 * u1 + a12 * t1;
    const double str0 = a21 * u1 + a22 * t1;
    double rr = fmax(fabs(amp0), fabs(str0));
    if (rr < 1.e-30) rr = 1.0;
    ul[k * 3 + 2] = log(rr) + v.eexl;
    ul[k * 3 + 0] = amp0 / rr;
    ul[k * 3 + 1] = str0 / rr;
  }
  // ---------------- shfunc rescale (:212-240): V(0) = 1, T(0) = 0
  {
    double ext = 0.0;
    ul[1] = 0.0;
    for (int k = max(1, k0); k < mmax; k++) {
      ext = ext + ul[(k - 1) * 3 + 2];
      double fact = 0.0;
      if (ext < 80.0) fact = 1. / exp(ext);
      ul[k * 3 + 0] *= fact;
      ul[k * 3 + 1] *= fact;
    }
    double umax = ul[0];
    if (umax == 0.0) {
      for (int k = 1; k < mmax; k++)
        if (fabs(ul[k * 3 + 0]) > fabs(umax)) umax = ul[k * 3 + 0];
    }
    if (fabs(umax) > 0.0) {
      for (int k = k0; k < mmax; k++) {
        ul[k * 3 + 0] /= umax;
        ul[k * 3 + 1] /= umax;
      }
    }
  }
  // ---------------- energy (:447-629)
  const double cph = omega / wvno;
  double sumi0 = 0.0, sumi1 = 0.0, sumi2 = 0.0;
  double zr_prev = 0.0, xmu_prev = 0.0;
  for (int k = 0; k < k0; k++)
    for (int p = 0; p < 4; p++) kern[((long long)p * mmax + k) * ks] = 0.0;
  for (int k = k0; k < mmax; k++) {
    const double zb = M.ld(F_B, k, b), zr = M.ld(F_RHO, k, b), dpth = M.ld(F_D, k, b);
    const double xmu = zr * zb * zb;
    const VarL v = varl_dev(zb, omega, wvno, dpth);
    double rb = v.rb;
    if (rb < 1.0e-10) rb = 1.0e-10;
    const double uk = ul[k * 3 + 0], tk = ul[k * 3 + 1];
    double upup, dupdup;
    if (k == mmax - 1) {
      upup = (0.5 / rb) * uk * uk;
      dupdup = (0.5 * rb) * uk * uk;
    } else {
      const bool osc = wvno < v.xkb;
      const cd nub = osc ? cd(0.0, rb) : cd(rb, 0.0);
      const cd xnub = xmu * nub;
      const cd iwx = cinv(wvno * xnub);
      const double hw = 0.5 / wvno;
      const double u1 = ul[(k + 1) * 3 + 0], t1 = ul[(k + 1) * 3 + 1];
      const cd km1dn = hw * uk - (0.5 * iwx) * tk;
      const cd kmup = hw * u1 + (0.5 * iwx) * t1;
      const cd f3 = nub * dpth;
      cd exqq = (f3.x < 40.0) ? cexp_b(-2.0 * f3) : cd(0.0);
      const cd f = (1.0 - exqq) / (2.0 * nub);
      exqq = (f3.x < 75.0) ? cexp_b(-1.0 * f3) : cd(0.0);
      const cd g = dpth * exqq;
      const double w2 = wvno * wvno;
      const cd f1 = f * (w2 * (kmup * kmup) + w2 * (km1dn * km1dn));
      const cd f2 = g * ((w2 + w2) * (kmup * km1dn));
      upup = (f1 + f2).x;
      dupdup = ((nub * nub) * (f1 - f2)).x;
    }
    sumi0 += zr * upup;
    sumi1 += xmu * upup;
    sumi2 += xmu * dupdup;
    kern[(0LL * mmax + k) * ks] = 0.0;  // Love has no vp sensitivity (reference leaves it unset)
    kern[(1LL * mmax + k) * ks] = cph * zr * zb * upup + cph * zr * zb * dupdup / wvno2;
    kern[(2LL * mmax + k) * ks] =
        0.5 * cph * (-cph * cph * upup + zb * zb * upup + zb * zb * dupdup / wvno2);
    // boundary term of dc/dh (:588-607)
    double drho, dmu, dvdz;
    if (k == k0) {
      drho = zr;
      dmu = xmu;
      dvdz = 0.0;
    } else {
      drho = zr - zr_prev;
      dmu = xmu - xmu_prev;
      dvdz = tk * tk * (1.0 / xmu - 1.0 / xmu_prev);
    }
    kern[(3LL * mmax + k) * ks] = uk * uk * (omega2 * drho - wvno2 * dmu) + dvdz;
    zr_prev = zr;
    xmu_prev = xmu;
  }
  const double ugr = sumi1 / (cph * sumi0);
  const double ale = 0.5 / sumi1;
  const double fac = ale * cph / wvno2;
  double suffix = 0.0;
  for (int k = mmax - 1; k >= 0; k--) {
    kern[(1LL * mmax + k) * ks] = kern[(1LL * mmax + k) * ks] / sumi1;
    kern[(2LL * mmax + k) * ks] = kern[(2LL * mmax + k) * ks] / sumi1;
    double dfac = fac * kern[(3LL * mmax + k) * ks];
    if (fabs(dfac) < 1.0e-38) dfac = 0.0;
    kern[(3LL * mmax + k) * ks] = suffix;
    suffix += dfac * M.ld(F_DTP, k, b);  // dtp = 1 on a flat earth (splove :660-664 otherwise)
  }
  *ugr_out = ugr;
}

}  // namespace rfs
