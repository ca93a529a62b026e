// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/README.md).
//
// Restates /root/reference/src/SWD/slegn96.f90 (Love eigenfunctions, energy integrals, group
// velocity, phase/group Frechet kernels):
//   bldsph :107-167  shfunc :179-246  varl :248-328  hskl :330-351  emat :353-370  up :372-445
//   energy :447-629  splove :631-670  slegn96 :672-783  slegnpu :785-919
#include "oracle.hpp"
#include <cmath>
#include <vector>

namespace oracle {
namespace {

struct LW {
  int mmax = 0;
  std::vector<double> zd, zb, zrho, xmu, uu, tt, dcdb, dcdh, dcdr, exl, vtp, dtp, rtp;
  std::vector<int> iwat;
  double uu0[4] = {0, 0, 0, 0};
  double cosq = 0, sinq = 0, yl = 0, zl = 0, mu = 0;
  double sumi0 = 0, sumi1 = 0, sumi2 = 0, flagr = 0, ale = 0, ugr = 0;
  void alloc(int n) {
    mmax = n;
    for (auto *v : {&zd, &zb, &zrho, &xmu, &uu, &tt, &dcdb, &dcdh, &dcdr, &exl}) v->assign(n, 0.0);
    iwat.assign(n, 0);
  }
};

// slegn96.f90:107-167
void bldsph(LW &S) {
  const int mmax = S.mmax;
  S.vtp.assign(mmax, 0.0);
  S.dtp.assign(mmax, 0.0);
  S.rtp.assign(mmax, 0.0);
  double ar = 6371.0, dr = 0.0, r0 = ar;
  S.zd[mmax - 1] = 1.0;
  for (int i = 0; i < mmax; i++) {
    dr = dr + S.zd[i];
    double r1 = ar - dr;
    double z0 = ar * std::log(ar / r0);
    double z1 = ar * std::log(ar / r1);
    S.dtp[i] = ar / r0;
    double tmp = (2.0 * ar) / (r0 + r1);
    S.vtp[i] = tmp;
    S.rtp[i] = 1.0 / (tmp * tmp * tmp * tmp * tmp);  // tmp**(-5)
    S.zb[i] = S.zb[i] * tmp;
    S.zrho[i] = S.zrho[i] * S.rtp[i];
    S.zd[i] = z1 - z0;
    r0 = r1;
  }
  S.zd[mmax - 1] = 0.0;
}

// slegn96.f90:248-328 (m 1-based)
void varl(LW &S, int m, double &rb, double omega, double wvno, double &xkb, double dpth,
          double &eexl) {
  xkb = omega / S.zb[m - 1];
  double wvnop = wvno + xkb;
  double wvnom = std::fabs(wvno - xkb);
  rb = std::sqrt(wvnop * wvnom);
  double q = rb * dpth;
  S.mu = S.zrho[m - 1] * S.zb[m - 1] * S.zb[m - 1];
  eexl = 0.0;
  if (wvno < xkb) {
    S.sinq = std::sin(q);
    S.yl = S.sinq / rb;
    S.zl = -rb * S.sinq;
    S.cosq = std::cos(q);
  } else if (wvno == xkb) {
    S.cosq = 1.0;
    S.yl = dpth;
    S.zl = 0.0;
  } else {
    eexl = q;
    double fac = 0.0;
    if (q < 18.0) fac = std::exp(-2.0 * q);
    S.cosq = (1.0 + fac) * 0.5;
    S.sinq = (1.0 - fac) * 0.5;
    S.yl = S.sinq / rb;
    S.zl = rb * S.sinq;
  }
}

// slegn96.f90:330-351
void hskl(LW &S, double hl[2][2], int iwat) {
  if (iwat == 0) {
    hl[0][0] = S.cosq;
    hl[0][1] = S.yl / S.mu;
    hl[1][0] = S.zl * S.mu;
    hl[1][1] = S.cosq;
  } else {
    hl[0][0] = 1.0;
    hl[0][1] = 0.0;
    hl[1][0] = 0.0;
    hl[1][1] = 1.0;
  }
}

// slegn96.f90:372-445
void up(LW &S, double omega, double wvno, double &fl) {
  const int mmax = S.mmax;
  double rb, xkb, eexl, hl[2][2];
  if (S.zb[mmax - 1] > 0.01) {
    double dpth = 0.0;
    varl(S, mmax, rb, omega, wvno, xkb, dpth, eexl);
    S.uu[mmax - 1] = 1.0;
    S.tt[mmax - 1] = -S.xmu[mmax - 1] * rb;
  } else {
    S.uu[mmax - 1] = 1.0;
    S.tt[mmax - 1] = 0.0;
  }
  S.exl[mmax - 1] = 0.0;
  double ttlast = 0.0;
  for (int k = mmax - 1; k >= 1; k--) {
    if (S.iwat[k - 1] == 0) {
      double dpth = S.zd[k - 1];
      varl(S, k, rb, omega, wvno, xkb, dpth, eexl);
      hskl(S, hl, S.iwat[k - 1]);
      int k1 = k + 1;
      double a11 = hl[0][0], a22 = hl[1][1], a12 = -hl[0][1], a21 = -hl[1][0];
      double amp0 = a11 * S.uu[k1 - 1] + a12 * S.tt[k1 - 1];
      double str0 = a21 * S.uu[k1 - 1] + a22 * S.tt[k1 - 1];
      double rr = std::fabs(amp0), ss = std::fabs(str0);
      if (ss > rr) rr = ss;
      if (rr < 1.e-30) rr = 1.0;
      S.exl[k - 1] = std::log(rr) + eexl;
      S.uu[k - 1] = amp0 / rr;
      S.tt[k - 1] = str0 / rr;
      ttlast = S.tt[k - 1];
    }
  }
  fl = ttlast;
}

// slegn96.f90:179-246
void shfunc(LW &S, double omega, double wvno) {
  const int mmax = S.mmax;
  double fl;
  up(S, omega, wvno, fl);
  S.uu0[0] = 1.0;
  S.uu0[1] = fl;
  S.uu0[2] = 0.0;
  S.uu0[3] = 0.0;
  double ext = 0.0;
  double umax = S.uu[0];
  S.tt[0] = 0.0;
  for (int k = 2; k <= mmax; k++) {
    if (S.iwat[k - 1] == 0) {
      ext = ext + S.exl[k - 2];
      double fact = 0.0;
      if (ext < 80.0) fact = 1. / std::exp(ext);
      S.uu[k - 1] = S.uu[k - 1] * fact;
      S.tt[k - 1] = S.tt[k - 1] * fact;
    } else {
      S.uu[k - 1] = 0.0;
      S.tt[k - 1] = 0.0;
    }
    if (std::fabs(S.uu[k - 1]) > std::fabs(umax)) umax = S.uu[k - 1];
  }
  if (S.uu[0] != 0.0) umax = S.uu[0];
  if (std::fabs(umax) > 0.0) {
    for (int k = 1; k <= mmax; k++) {
      if (S.iwat[k - 1] == 0) {
        S.uu[k - 1] = S.uu[k - 1] / umax;
        S.tt[k - 1] = S.tt[k - 1] / umax;
      }
    }
  }
}

// slegn96.f90:447-629 (source/receiver eigenfunction outputs Eut.. are unused by the callers)
void energy(LW &S, double omega, double wvno) {
  const int mmax = S.mmax;
  double c = omega / wvno, omega2 = omega * omega, wvno2 = wvno * wvno;
  S.sumi0 = S.sumi1 = S.sumi2 = 0.0;
  for (int k = 1; k <= mmax; k++) {
    if (S.iwat[k - 1] == 0) {
      double zb = S.zb[k - 1], zrho = S.zrho[k - 1];
      double TN = zrho * zb * zb, TL = zrho * zb * zb;
      double VSHH = zb, VSHV = zb;
      int k1 = k + 1;
      double drho = zrho, dpth = S.zd[k - 1];
      double rb, xkb, eexl;
      varl(S, k, rb, omega, wvno, xkb, dpth, eexl);
      double dmu = S.xmu[k - 1];
      if (rb < 1.0e-10) rb = 1.0e-10;
      double upup, dupdup;
      if (k == mmax) {
        upup = (0.5 / rb) * S.uu[mmax - 1] * S.uu[mmax - 1];
        dupdup = (0.5 * rb) * S.uu[mmax - 1] * S.uu[mmax - 1];
      } else {
        cplx nub = cplx(rb, 0.0);
        if (wvno < xkb) nub = cplx(0.0, rb);
        cplx xnub = dmu * nub;
        // emat :353-370
        cplx einvl[2][2], el[2][2];
        einvl[0][0] = 0.5 / wvno;
        einvl[0][1] = 0.5 / (wvno * xnub);
        einvl[1][0] = 0.5 / wvno;
        einvl[1][1] = -0.5 / (wvno * xnub);
        el[0][0] = wvno;
        el[0][1] = wvno;
        el[1][0] = wvno * xnub;
        el[1][1] = -wvno * xnub;
        cplx km1dn = einvl[1][0] * S.uu[k - 1] + einvl[1][1] * S.tt[k - 1];
        cplx kmup = einvl[0][0] * S.uu[k1 - 1] + einvl[0][1] * S.tt[k1 - 1];
        cplx f3 = nub * dpth;
        cplx exqq = 0.0;
        if (f3.real() < 40.0) exqq = std::exp(-2.0 * f3);
        cplx f = (1.0 - exqq) / (2.0 * nub);
        exqq = 0.0;
        if (f3.real() < 75.0) exqq = std::exp(-f3);
        cplx g = dpth * exqq;
        cplx f1 = f * (el[0][0] * el[0][0] * kmup * kmup + el[0][1] * el[0][1] * km1dn * km1dn);
        cplx f2 = g * (el[0][0] * el[0][1] + el[0][0] * el[0][1]) * kmup * km1dn;
        upup = (f1 + f2).real();
        dupdup = (nub * nub * (f1 - f2)).real();
      }
      S.sumi0 += drho * upup;
      S.sumi1 += TN * upup;
      S.sumi2 += TL * dupdup;
      double DCDBH = c * drho * VSHH * upup;
      double DCDBV = c * drho * VSHV * dupdup / wvno2;
      S.dcdb[k - 1] = DCDBH + DCDBV;
      double DCDRSH = 0.5 * c * (-c * c * upup + VSHH * VSHH * upup + VSHV * VSHV * dupdup / wvno2);
      S.dcdr[k - 1] = DCDRSH;
    } else {
      S.dcdb[k - 1] = 0.0;
      S.dcdr[k - 1] = 0.0;
    }
  }
  for (int k = 1; k <= mmax; k++) {
    if (S.iwat[k - 1] == 0) {
      S.dcdb[k - 1] = S.dcdb[k - 1] / S.sumi1;
      S.dcdr[k - 1] = S.dcdr[k - 1] / S.sumi1;
    } else {
      S.dcdb[k - 1] = 0.0;
      S.dcdr[k - 1] = 0.0;
    }
  }
  S.flagr = omega2 * S.sumi0 - wvno2 * S.sumi1 - S.sumi2;
  S.ugr = S.sumi1 / (c * S.sumi0);
  S.ale = 0.5 / S.sumi1;
  double fac = S.ale * c / wvno2;
  int llflag = 0;
  for (int k = 1; k <= mmax; k++) {
    if (S.iwat[k - 1] == 0) {
      double drho, dmu, dvdz;
      if (llflag == 0) {
        drho = S.zrho[k - 1];
        dmu = S.xmu[k - 1];
        dvdz = 0.0;
      } else {
        drho = S.zrho[k - 1] - S.zrho[k - 2];
        dmu = S.xmu[k - 1] - S.xmu[k - 2];
        dvdz = S.tt[k - 1] * S.tt[k - 1] * (1.0 / S.xmu[k - 1] - 1.0 / S.xmu[k - 2]);
      }
      double dfac = fac * (S.uu[k - 1] * S.uu[k - 1] * (omega2 * drho - wvno2 * dmu) + dvdz);
      if (std::fabs(dfac) < 1.0e-38)
        S.dcdh[k - 1] = 0.0;
      else
        S.dcdh[k - 1] = dfac;
      llflag = llflag + 1;
    } else {
      S.dcdh[k - 1] = 0.0;
    }
  }
}

// slegn96.f90:631-670
void splove(LW &S, double om, double c, double &csph, double &usph, double ugr) {
  double a = 6371.0;
  double x = 3.0 * c / (2. * a * om);
  double tm = std::sqrt(1. + x * x);
  double tm3 = tm * tm * tm;
  for (int i = 0; i < S.mmax; i++) {
    S.dcdb[i] = S.dcdb[i] * S.vtp[i] / tm3;
    S.dcdh[i] = S.dcdh[i] * S.dtp[i] / tm3;
    S.dcdr[i] = S.dcdr[i] * S.rtp[i] / tm3;
  }
  csph = c / tm;
  usph = ugr * tm;
}

void setup(LW &S, const float *thk, const float *vs, const float *rhom, int nlayer, int iflsph,
           bool eq_zero_test) {
  S.alloc(nlayer);
  for (int i = 0; i < nlayer; i++) {
    S.zb[i] = (double)vs[i];
    S.zrho[i] = (double)rhom[i];
    S.zd[i] = (double)thk[i];
  }
  for (int i = 0; i < nlayer; i++) {
    if (eq_zero_test)
      S.iwat[i] = (S.zb[i] == 0.0) ? 1 : 0;  // slegn96 :707-713
    else
      S.iwat[i] = (S.zb[i] > 0.0) ? 0 : 1;   // slegnpu :824-830
  }
  if (iflsph > 0) bldsph(S);
  for (int i = 0; i < nlayer; i++) S.xmu[i] = S.zrho[i] * S.zb[i] * S.zb[i];
}

}  // namespace

void slegn96(const float *thk, const float *vs, const float *rhom, int nlayer, double *t,
             double *cp, double *cg, double *disp, double *stress, double *dc2db, double *dc2dh,
             double *dc2dr, int iflsph) {
  LW S;
  setup(S, thk, vs, rhom, nlayer, iflsph, true);
  const int mmax = S.mmax;
  // `pi = 3.1415926535898` is a REAL*4 literal (:689) -> float32 pi
  double twopi = 2.0 * (double)3.1415926535898f;
  double omega = twopi / *t;
  double c = *cp;
  double wvno = omega / c;
  shfunc(S, omega, wvno);
  energy(S, omega, wvno);
  double csph, usph;
  if (iflsph > 0) {
    splove(S, omega, c, csph, usph, S.ugr);
  } else {
    csph = c;
    usph = S.ugr;
  }
  if (std::fabs(S.ugr) < 1.0e-36) S.ugr = 0.0;  // (after usph was taken, as in the reference)
  for (int i = 0; i < mmax; i++) dc2dh[i] = S.dcdh[i];
  for (int i = 1; i <= mmax - 1; i++) {
    double sums = 0.0;
    for (int j = i + 1; j <= mmax; j++) sums += dc2dh[j - 1];
    S.dcdh[i - 1] = sums;
  }
  S.dcdh[mmax - 1] = 0.0;
  for (int i = 0; i < mmax; i++) {
    disp[i] = S.uu[i];
    stress[i] = S.uu[i];  // reference typo `stress(:) = uu(:)` (:770); output unused
    dc2db[i] = S.dcdb[i];
    dc2dr[i] = S.dcdr[i];
    dc2dh[i] = S.dcdh[i];
  }
  *cp = csph;
  *cg = usph;
}

void slegnpu(const float *thk, const float *vs, const float *rhom, int nlayer, double *t,
             double *cp, double *cg, double *disp, double *stress, double *t1, double *cp1,
             double *t2, double *cp2, double *dc2db, double *dc2dh, double *dc2dr, double *du2db,
             double *du2dh, double *du2dr, int iflsph, bool stale_first_term) {
  LW S;
  setup(S, thk, vs, rhom, nlayer, iflsph, false);
  const int mmax = S.mmax;
  std::vector<double> b1(mmax), b2(mmax), h1(mmax), h2(mmax), r1(mmax), r2(mmax);
  double twopi = 2.0 * PI32;  // atan(1.0)*4.0, :804
  double omega = twopi / *t;
  double c = *cp;
  double wvno = omega / c;
  shfunc(S, omega, wvno);
  energy(S, omega, wvno);
  *cg = S.ugr;
  for (int i = 0; i < mmax; i++) {
    dc2db[i] = S.dcdb[i];
    stress[i] = S.tt[i];
    dc2dr[i] = S.dcdr[i];
    disp[i] = S.uu[i];
    dc2dh[i] = S.dcdh[i];
  }
  omega = twopi / *t1;
  c = *cp1;
  wvno = omega / c;
  shfunc(S, omega, wvno);
  energy(S, omega, wvno);
  for (int i = 0; i < mmax; i++) {
    b1[i] = S.dcdb[i];
    r1[i] = S.dcdr[i];
    h1[i] = S.dcdh[i];
  }
  omega = twopi / *t2;
  c = *cp2;
  wvno = omega / c;
  shfunc(S, omega, wvno);
  energy(S, omega, wvno);
  for (int i = 0; i < mmax; i++) {
    b2[i] = S.dcdb[i];
    r2[i] = S.dcdr[i];
    h2[i] = S.dcdh[i];
  }
  double uc1 = *cg / *cp;
  for (int i = 0; i < mmax; i++) {
    // :876-878 — first term uses the module arrays (T2 solve) in the reference.
    double fb = stale_first_term ? S.dcdb[i] : dc2db[i];
    double fr = stale_first_term ? S.dcdr[i] : dc2dr[i];
    double fh = stale_first_term ? S.dcdh[i] : dc2dh[i];
    du2db[i] = uc1 * (2.0 - uc1) * fb - uc1 * uc1 * *t * (b2[i] - b1[i]) / (*t2 - *t1);
    du2dr[i] = uc1 * (2.0 - uc1) * fr - uc1 * uc1 * *t * (r2[i] - r1[i]) / (*t2 - *t1);
    du2dh[i] = uc1 * (2.0 - uc1) * fh - uc1 * uc1 * *t * (h2[i] - h1[i]) / (*t2 - *t1);
  }
  if (iflsph > 0) {
    double ar = 6371.0;
    omega = twopi / *t;
    double x = 3.0 * *cp / (2. * ar * omega);
    double tm = std::sqrt(1. + x * x);
    double y = 1.5 / (ar * omega);
    double tm1 = y * y / tm;
    double tm3 = tm * tm * tm;
    for (int i = 0; i < mmax; i++) {
      du2db[i] = (tm * du2db[i] + *cg * *cp * dc2db[i] * tm1) * S.vtp[i];
      du2dr[i] = (tm * du2dr[i] + *cg * *cp * dc2dr[i] * tm1) * S.rtp[i];
      du2dh[i] = (tm * du2dh[i] + *cg * *cp * dc2dh[i] * tm1) * S.dtp[i];
      dc2db[i] = dc2db[i] / tm3 * S.vtp[i];
      dc2dr[i] = dc2dr[i] / tm3 * S.rtp[i];
      dc2dh[i] = dc2dh[i] / tm3 * S.dtp[i];
    }
    *cp = *cp / tm;
    *cg = *cg * tm;
  }
  for (int i = 1; i <= mmax - 1; i++) {
    double s1 = 0.0, s2 = 0.0;
    for (int j = i + 1; j <= mmax; j++) {
      s1 += dc2dh[j - 1];
      s2 += du2dh[j - 1];
    }
    dc2dh[i - 1] = s1;
    du2dh[i - 1] = s2;
  }
  dc2dh[mmax - 1] = 0.0;
  du2dh[mmax - 1] = 0.0;
}

}  // namespace oracle
