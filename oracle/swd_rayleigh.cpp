// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/README.md).
//
// Restates /root/reference/src/SWD/sregn96.f90 (Rayleigh eigenfunctions, energy integrals,
// group velocity and phase/group Frechet kernels):
//   bldsph :133-187   svfunc :196-402   up :404-492     dnka :494-650   evalg :652-829
//   varsv :831-915    hska :917-991     down :993-1063  energy :1065-1201
//   intijr :1203-1323 ffunc/gfunc/h1func/h2func :1325-1403  normc :1405-1434
//   getdcdh :1436-1535 getmat :1537-1589 sprayl :1591-1635
//   sregn96 :1637-1745 sregnpu :1747-1888
// The Fortran module-global state becomes the struct `RW`.
#include "oracle.hpp"
#include <cmath>
#include <vector>

namespace oracle {
namespace {

struct RW {
  int mmax = 0;
  std::vector<double> zd, zrho, za, zb, xmu, xlam;
  double uu0[4] = {0, 0, 0, 0};
  std::vector<double> ur, uz, tz, tr, dcda, dcdb, dcdr, dcdh;
  std::vector<int> iwat;
  double sumi0 = 0, sumi1 = 0, sumi2 = 0, sumi3 = 0, flagr = 0, are = 0, ugr = 0;
  std::vector<double> vtp, dtp, rtp;
  std::vector<double> exe, exa;
  std::vector<double> cd;  // [mmax][5]
  std::vector<double> vv;  // [mmax][4]
  bool allfluid = false;
  cplx ra, rb;
  cplx e[4][4], einv[4][4];

  void alloc(int n) {
    mmax = n;
    for (auto *v : {&zd, &zrho, &za, &zb, &xmu, &xlam, &ur, &uz, &tz, &tr, &dcda, &dcdb, &dcdr,
                    &dcdh, &exe, &exa})
      v->assign(n, 0.0);
    iwat.assign(n, 0);
    cd.assign((size_t)n * 5, 0.0);
    vv.assign((size_t)n * 4, 0.0);
  }
  double &CD(int m, int j) { return cd[(size_t)(m - 1) * 5 + (j - 1)]; }  // 1-based
  double &VV(int m, int j) { return vv[(size_t)(m - 1) * 4 + (j - 1)]; }
};

// sregn96.f90:133-187
void bldsph(RW &S) {
  const int mmax = S.mmax;
  double ar = 6371.0, dr = 0.0, r0 = ar;
  S.vtp.assign(mmax, 0.0);
  S.dtp.assign(mmax, 0.0);
  S.rtp.assign(mmax, 0.0);
  for (int i = 0; i < mmax; i++) {
    if (i == mmax - 1)
      dr = dr + 1.0;
    else
      dr = dr + S.zd[i];
    double r1 = ar - dr;
    double z0 = ar * std::log(ar / r0);
    double z1 = ar * std::log(ar / r1);
    double tmp = (2.0 * ar) / (r0 + r1);
    S.vtp[i] = tmp;
    S.rtp[i] = std::pow(tmp, (double)(-2.275f));  // tmp**(-2.275): REAL*4 exponent promoted
    S.dtp[i] = ar / r0;
    S.za[i] = S.za[i] * tmp;
    S.zb[i] = S.zb[i] * tmp;
    S.zrho[i] = S.zrho[i] * S.rtp[i];
    S.zd[i] = z1 - z0;
    r0 = r1;
  }
  S.zd[mmax - 1] = 0.0;
}

// sregn96.f90:1405-1434
void normc(double *ee, double &ex, int nmat) {
  ex = 0.0;
  double t1 = 0.0;
  for (int i = 0; i < nmat; i++)
    if (std::fabs(ee[i]) > t1) t1 = std::fabs(ee[i]);
  if (t1 < 1.e-40) t1 = 1.0;
  for (int i = 0; i < nmat; i++) ee[i] = ee[i] / t1;
  ex = std::log(t1);
}

// sregn96.f90:831-915
void varsv(cplx p, cplx q, cplx rp, cplx rsv, double &cosp, double &cosq, double &rsinp,
           double &rsinq, double &sinpr, double &sinqr, double &pex, double &svex, int iwat,
           double zd) {
  const double small = (double)1.0e-5f;  // `1.0e-5` REAL*4 literal
  pex = 0.0;
  svex = 0.0;
  double pr = p.real(), pi = p.imag(), qr = q.real(), qi = q.imag();
  pex = pr;
  cplx epp = cplx(std::cos(pi), std::sin(pi)) / 2.0;
  cplx epm = std::conj(epp);
  double pfac;
  if (pr < 30.)
    pfac = std::exp(-2. * pr);
  else
    pfac = 0.0;
  cosp = (epp + pfac * epm).real();
  cplx sinp = epp - pfac * epm;
  rsinp = (rp * sinp).real();
  if (std::fabs(pr) < small && std::abs(rp) < small)
    sinpr = zd;
  else
    sinpr = (sinp / rp).real();
  if (iwat == 1) {
    cosq = 1.0;
    rsinq = 0.0;
    sinqr = 0.0;
  } else {
    svex = qr;
    cplx eqp = cplx(std::cos(qi), std::sin(qi)) / 2.0;
    cplx eqm = std::conj(eqp);
    double svfac;
    if (qr < 30.)
      svfac = std::exp(-2. * qr);
    else
      svfac = 0.0;
    cosq = (eqp + svfac * eqm).real();
    cplx sinq = eqp - svfac * eqm;
    rsinq = (rsv * sinq).real();
    if (std::fabs(qr) < small && std::abs(rsv) < small)
      sinqr = zd;
    else
      sinqr = (sinq / rsv).real();
  }
}

// sregn96.f90:494-650 — reduced 5x5 Dunkin compound matrix (1-based ca via CA()).
void dnka(double ca[5][5], double cosp, double rsinp, double sinpr, double cossv, double rsinsv,
          double sinsvr, float rho, float b, int iwat, double ex, double exa, double wvno,
          double wvno2, double om2) {
#define CA(i, j) ca[(i)-1][(j)-1]
  if (iwat == 1) {
    for (int j = 0; j < 5; j++)
      for (int i = 0; i < 5; i++) ca[i][j] = 0.0;
    double dfac;
    if (ex > 35.0)
      dfac = 0.0;
    else
      dfac = std::exp(-ex);
    CA(3, 3) = dfac;
    CA(1, 1) = cosp;
    CA(5, 5) = cosp;
    CA(1, 2) = -rsinp / ((double)rho * om2);
    CA(2, 1) = -(double)rho * sinpr * om2;
    CA(2, 2) = cosp;
    CA(4, 4) = cosp;
    CA(4, 5) = CA(1, 2);
    CA(5, 4) = CA(2, 1);
  } else {
    double a0;
    if (exa < 60.0)
      a0 = std::exp(-exa);
    else
      a0 = 0.0;
    double cpcq = cosp * cossv;
    double cpy = cosp * sinsvr;
    double cpz = cosp * rsinsv;
    double cqw = cossv * sinpr;
    double cqx = cossv * rsinp;
    double xy = rsinp * sinsvr;
    double xz = rsinp * rsinsv;
    double wy = sinpr * sinsvr;
    double wz = sinpr * rsinsv;
    float rho2 = rho * rho;                              // REAL*4
    double gam = (double)(2.0f * b * b) * wvno2 / om2;   // 2.0*b*b evaluated in REAL*4
    double gam2 = gam * gam;
    double gamm1 = gam - 1.;
    double gamm2 = gamm1 * gamm1;
    double cqww2 = cqw * wvno2;
    double cqxw2 = cqx / wvno2;
    double gg1 = gam * gamm1;
    double a0c = 2.0 * (a0 - cpcq);
    double xz2 = xz / wvno2;
    double gxz2 = gam * xz2;
    double g2xz2 = gam2 * xz2;
    double a0cgg1 = a0c * (gam + gamm1);
    double wy2 = wy * wvno2;
    double g2wy2 = gamm2 * wy2;
    double g1wy2 = gamm1 * wy2;
    double rom2 = (double)rho * om2;
    double temp = a0c * gg1 + g2xz2 + g2wy2;
    CA(3, 3) = a0 + temp + temp;
    CA(1, 1) = cpcq - temp;
    CA(1, 2) = (-cqx + wvno2 * cpy) / rom2;
    temp = 0.5 * a0cgg1 + gxz2 + g1wy2;
    CA(1, 3) = wvno * temp / rom2;
    CA(1, 4) = (-cqww2 + cpz) / rom2;
    temp = wvno2 * (a0c + wy2) + xz;
    CA(1, 5) = -temp / ((double)rho2 * om2 * om2);
    CA(2, 1) = (-gamm2 * cqw + gam2 * cpz / wvno2) * (double)rho * om2;
    CA(2, 2) = cpcq;
    CA(2, 3) = (gamm1 * cqww2 - gam * cpz) / wvno;
    CA(2, 4) = -wz;
    CA(2, 5) = CA(1, 4);
    temp = 0.5 * a0cgg1 * gg1 + gam2 * gxz2 + gamm2 * g1wy2;
    CA(3, 1) = -2.0 * temp * (double)rho * om2 / wvno;
    CA(3, 2) = -wvno * (gam * cqxw2 - gamm1 * cpy) * 2.0;
    CA(3, 4) = -2.0 * CA(2, 3);
    CA(3, 5) = -2.0 * CA(1, 3);
    CA(4, 1) = (-gam2 * cqxw2 + gamm2 * cpy) * (double)rho * om2;
    CA(4, 2) = -xy;
    CA(4, 3) = -CA(3, 2) / 2.0;
    CA(4, 4) = CA(2, 2);
    CA(4, 5) = CA(1, 2);
    temp = gamm2 * (a0c * gam2 + g2wy2) + gam2 * g2xz2;
    CA(5, 1) = -(double)rho2 * om2 * om2 * temp / wvno2;
    CA(5, 2) = CA(4, 1);
    CA(5, 3) = -CA(3, 1) / 2.0;
    CA(5, 4) = CA(2, 1);
    CA(5, 5) = CA(1, 1);
  }
#undef CA
}

// sregn96.f90:652-829 — half-space vector gbr(inp,1:5) and the E / E^-1 matrices of layer m.
void evalg(RW &S, int jbdry, int m, cplx gbr[5], double wvno, double om, double om2,
           double wvno2) {
#define E(i, j) S.e[(i)-1][(j)-1]
#define EINV(i, j) S.einv[(i)-1][(j)-1]
  const double zr = S.zrho[m - 1];
  double xka = om / S.za[m - 1];
  double xkb;
  if (S.zb[m - 1] > 0.0)
    xkb = om / S.zb[m - 1];
  else
    xkb = 0.0;
  S.ra = std::sqrt(cplx(wvno2 - xka * xka, 0.0));
  S.rb = std::sqrt(cplx(wvno2 - xkb * xkb, 0.0));
  const cplx ra = S.ra, rb = S.rb;
  double gam = S.zb[m - 1] * wvno / om;
  gam = 2.0 * (gam * gam);
  double gamm1 = gam - 1.0;
  if (jbdry < 0) {
    for (int i = 0; i < 5; i++) gbr[i] = 0.0;
    if (S.zb[m - 1] > 0.0) {
      gbr[0] = 1.0;
    } else {
      if (S.allfluid)
        gbr[0] = 1.0;
      else
        gbr[3] = 1.0;
    }
  } else if (jbdry == 0) {
    if (S.iwat[m - 1] == 0) {
      E(1, 1) = wvno;
      E(1, 2) = rb;
      E(1, 3) = wvno;
      E(1, 4) = -rb;
      E(2, 1) = ra;
      E(2, 2) = wvno;
      E(2, 3) = -ra;
      E(2, 4) = wvno;
      E(3, 1) = zr * om2 * gamm1;
      E(3, 2) = zr * om2 * gam * rb / wvno;
      E(3, 3) = zr * om2 * gamm1;
      E(3, 4) = -zr * om2 * gam * rb / wvno;
      E(4, 1) = zr * om2 * gam * ra / wvno;
      E(4, 2) = zr * om2 * gamm1;
      E(4, 3) = -zr * om2 * gam * ra / wvno;
      E(4, 4) = zr * om2 * gamm1;

      EINV(1, 1) = 0.5 * gam / wvno;
      EINV(1, 2) = -0.5 * gamm1 / ra;
      EINV(1, 3) = -0.5 / (zr * om2);
      EINV(1, 4) = 0.5 * wvno / (zr * om2 * ra);
      EINV(2, 1) = -0.5 * gamm1 / rb;
      EINV(2, 2) = 0.5 * gam / wvno;
      EINV(2, 3) = 0.5 * wvno / (zr * om2 * rb);
      EINV(2, 4) = -0.5 / (zr * om2);
      EINV(3, 1) = 0.5 * gam / wvno;
      EINV(3, 2) = 0.5 * gamm1 / ra;
      EINV(3, 3) = -0.5 / (zr * om2);
      EINV(3, 4) = -0.5 * wvno / (zr * om2 * ra);
      EINV(4, 1) = 0.5 * gamm1 / rb;
      EINV(4, 2) = 0.5 * gam / wvno;
      EINV(4, 3) = -0.5 * wvno / (zr * om2 * rb);
      EINV(4, 4) = -0.5 / (zr * om2);

      gbr[0] = (zr * zr) * om2 * om2 * (-gam * gam * ra * rb + wvno2 * gamm1 * gamm1);
      gbr[1] = -zr * (wvno2 * ra) * om2;
      gbr[2] = -zr * (-gam * ra * rb + wvno2 * gamm1) * om2 * wvno;
      gbr[3] = zr * (wvno2 * rb) * om2;
      gbr[4] = wvno2 * (wvno2 - ra * rb);
      cplx den = (-zr * zr * om2 * om2 * wvno2 * ra * rb);
      for (int i = 0; i < 5; i++) gbr[i] = 0.25 * gbr[i] / den;
    } else {
      for (int i = 0; i < 5; i++) gbr[i] = 0.0;
      if (S.allfluid) {
        gbr[0] = 0.5 / ra;
        gbr[1] = cplx(0.5, 0.0) / (-zr * om2);
      } else {
        gbr[3] = (0.5 * zr * om2) / ra;
        gbr[4] = cplx(-0.5, 0.0);
      }
      for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
          S.e[i][j] = 0.0;
          S.einv[i][j] = 0.0;
        }
      E(1, 1) = ra;
      E(1, 2) = -ra;
      E(2, 1) = -zr * om2;
      E(2, 2) = -zr * om2;
      EINV(1, 1) = 0.5 / ra;
      EINV(1, 2) = -0.5 / (zr * om2);
      EINV(2, 1) = -0.5 / ra;
      EINV(2, 2) = -0.5 / (zr * om2);
    }
  } else {
    for (int i = 0; i < 5; i++) gbr[i] = 0.0;
    if (S.zb[m - 1] > 0.0) {
      gbr[4] = 1.0;
    } else {
      if (S.allfluid)
        gbr[1] = 1.0;
      else
        gbr[4] = 1.0;
    }
  }
#undef E
#undef EINV
}

// sregn96.f90:917-991 — 4x4 Haskell propagator (1-based AA()).
void hska(double aa[4][4], double cosp, double rsinp, double sinpr, double tcossv, double trsinsv,
          double tsinsvr, float rho, float b, int iwat, double pex, double svex, double wvno,
          double wvno2, double om2) {
#define AA(i, j) aa[(i)-1][(j)-1]
  if (iwat == 1) {
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) aa[i][j] = 0.0;
    double dfac;
    if (pex > 35.0)
      dfac = 0.0;
    else
      dfac = std::exp(-pex);
    AA(1, 1) = dfac;
    AA(4, 4) = dfac;
    AA(2, 2) = cosp;
    AA(3, 3) = cosp;
    AA(2, 3) = -rsinp / ((double)rho * om2);
    AA(3, 2) = -(double)rho * om2 * sinpr;
  } else {
    double dfac;
    if ((pex - svex) > 70.0)
      dfac = 0.0;
    else
      dfac = std::exp(svex - pex);
    double cossv = dfac * tcossv;
    double rsinsv = dfac * trsinsv;
    double sinsvr = dfac * tsinsvr;
    double gam = (double)(2.0f * b * b) * wvno2 / om2;
    double gamm1 = gam - 1.0;
    double rom2 = (double)rho * om2;
    AA(1, 1) = cossv + gam * (cosp - cossv);
    AA(1, 2) = -wvno * gamm1 * sinpr + gam * rsinsv / wvno;
    AA(1, 3) = -wvno * (cosp - cossv) / rom2;
    AA(1, 4) = (wvno2 * sinpr - rsinsv) / rom2;
    AA(2, 1) = gam * rsinp / wvno - wvno * gamm1 * sinsvr;
    AA(2, 2) = cosp - gam * (cosp - cossv);
    AA(2, 3) = (-rsinp + wvno2 * sinsvr) / rom2;
    AA(2, 4) = -AA(1, 3);
    AA(3, 1) = rom2 * gam * gamm1 * (cosp - cossv) / wvno;
    AA(3, 2) = rom2 * (-gamm1 * gamm1 * sinpr + gam * gam * rsinsv / wvno2);
    AA(3, 3) = AA(2, 2);
    AA(3, 4) = -AA(1, 2);
    AA(4, 1) = rom2 * (gam * gam * rsinp / wvno2 - gamm1 * gamm1 * sinsvr);
    AA(4, 2) = -AA(3, 1);
    AA(4, 3) = -AA(2, 1);
    AA(4, 4) = AA(1, 1);
  }
#undef AA
}

// sregn96.f90:404-492
void up(RW &S, double omega, double wvno, double &fr) {
  const int mmax = S.mmax;
  double wvno2 = wvno * wvno, om2 = omega * omega;
  cplx gbr[5];
  evalg(S, 0, mmax, gbr, wvno, omega, om2, wvno2);
  for (int j = 1; j <= 5; j++) S.CD(mmax, j) = gbr[j - 1].real();
  S.exe[mmax - 1] = 0.0;
  double exsum = 0.0;
  double ca[5][5], ee[5];
  for (int m = mmax - 1; m >= 1; m--) {
    double xka = omega / S.za[m - 1];
    double xkb = (S.zb[m - 1] > 0.0) ? omega / S.zb[m - 1] : 0.0;
    cplx rp = std::sqrt(cplx(wvno2 - xka * xka, 0.0));
    cplx rsv = std::sqrt(cplx(wvno2 - xkb * xkb, 0.0));
    cplx p = rp * S.zd[m - 1];
    cplx q = rsv * S.zd[m - 1];
    double cosp, cossv, rsinp, rsinsv, sinpr, sinsvr, pex, svex;
    varsv(p, q, rp, rsv, cosp, cossv, rsinp, rsinsv, sinpr, sinsvr, pex, svex, S.iwat[m - 1],
          S.zd[m - 1]);
    dnka(ca, cosp, rsinp, sinpr, cossv, rsinsv, sinsvr, (float)S.zrho[m - 1], (float)S.zb[m - 1],
         S.iwat[m - 1], pex, pex + svex, wvno, wvno2, om2);
    for (int i = 0; i < 5; i++) {
      double cr = 0.0;
      for (int j = 0; j < 5; j++) cr = cr + S.CD(m + 1, j + 1) * ca[j][i];
      ee[i] = cr;
    }
    double exn = 0.0;
    normc(ee, exn, 5);
    exsum = exsum + pex + svex + exn;
    S.exe[m - 1] = exsum;
    for (int i = 0; i < 5; i++) S.CD(m, i + 1) = ee[i];
  }
  fr = S.CD(1, 1);
}

// sregn96.f90:993-1063
void down(RW &S, double omega, double wvno) {
  const int mmax = S.mmax;
  double om2 = omega * omega, wvno2 = wvno * wvno;
  for (int i = 1; i <= 4; i++) S.VV(1, i) = (i == 1) ? 1.0 : 0.0;
  S.exa[0] = 0.0;
  double exsum = 0.0;
  double aa[4][4], aa0[4];
  for (int m = 1; m <= mmax - 1; m++) {
    double xka = omega / S.za[m - 1];
    double xkb = (S.zb[m - 1] > 0.0) ? omega / S.zb[m - 1] : 0.0;
    cplx rp = std::sqrt(cplx(wvno2 - xka * xka, 0.0));
    cplx rsv = std::sqrt(cplx(wvno2 - xkb * xkb, 0.0));
    cplx p = rp * S.zd[m - 1];
    cplx q = rsv * S.zd[m - 1];
    double cosp, cossv, rsinp, rsinsv, sinpr, sinsvr, pex, svex;
    varsv(p, q, rp, rsv, cosp, cossv, rsinp, rsinsv, sinpr, sinsvr, pex, svex, S.iwat[m - 1],
          S.zd[m - 1]);
    hska(aa, cosp, rsinp, sinpr, cossv, rsinsv, sinsvr, (float)S.zrho[m - 1], (float)S.zb[m - 1],
         S.iwat[m - 1], pex, svex, wvno, wvno2, om2);
    for (int i = 0; i < 4; i++) {
      double cc = 0.0;
      for (int j = 0; j < 4; j++) cc = cc + aa[i][j] * S.VV(m, j + 1);
      aa0[i] = cc;
    }
    double ex2 = 0.0;
    normc(aa0, ex2, 4);
    exsum = exsum + pex + ex2;
    S.exa[m] = exsum;
    for (int i = 0; i < 4; i++) S.VV(m + 1, i + 1) = aa0[i];
  }
}

// sregn96.f90:196-402
void svfunc(RW &S, double omega, double wvno) {
  const int mmax = S.mmax;
  double fr;
  up(S, omega, wvno, fr);
  down(S, omega, wvno);
  double f1213 = -S.CD(1, 2);
  S.ur[0] = S.CD(1, 3) / S.CD(1, 2);
  S.uz[0] = 1.0;
  S.tz[0] = 0.0;
  S.tr[0] = 0.0;
  S.uu0[0] = S.ur[0];
  S.uu0[1] = 1.0;
  S.uu0[2] = fr;
  S.uu0[3] = fr;
  for (int i = 2; i <= mmax; i++) {
    double cd1 = S.CD(i, 1), cd2 = S.CD(i, 2), cd3 = S.CD(i, 3), cd4 = -S.CD(i, 3),
           cd5 = S.CD(i, 4), cd6 = S.CD(i, 5);
    double tz1 = -S.VV(i, 4), tz2 = -S.VV(i, 3), tz3 = S.VV(i, 2), tz4 = S.VV(i, 1);
    double uu1 = tz2 * cd6 - tz3 * cd5 + tz4 * cd4;
    double uu2 = -tz1 * cd6 + tz3 * cd3 - tz4 * cd2;
    double uu3 = tz1 * cd5 - tz2 * cd3 + tz4 * cd1;
    double uu4 = -tz1 * cd4 + tz2 * cd2 - tz3 * cd1;
    double ext = S.exa[i - 1] + S.exe[i - 1] - S.exe[0];
    if (ext > -80.0 && ext < 80.0) {
      double fact = std::exp(ext);
      S.ur[i - 1] = uu1 * fact / f1213;
      S.uz[i - 1] = uu2 * fact / f1213;
      S.tz[i - 1] = uu3 * fact / f1213;
      S.tr[i - 1] = uu4 * fact / f1213;
    } else {
      S.ur[i - 1] = 0.0;
      S.uz[i - 1] = 0.0;
      S.tz[i - 1] = 0.0;
      S.tr[i - 1] = 0.0;
    }
  }
  if (!S.allfluid) {
    int jwat = 0;
    for (int i = 1; i <= mmax; i++) {
      if (S.iwat[i - 1] > 0)
        jwat = i;
      else
        break;
    }
    for (int i = 1; i <= jwat; i++) {
      S.ur[i - 1] = 0.0;
      S.tr[i - 1] = 0.0;
    }
  }
}

// sregn96.f90:1325-1403
cplx ffunc(cplx nub, double dm) {
  if (std::abs(nub) < 1.0e-08) return dm;
  cplx argcd = nub * dm;
  cplx exqq;
  if (argcd.real() < 40.0)
    exqq = std::exp(-2.0 * argcd);
  else
    exqq = 0.0;
  return (1.0 - exqq) / (2.0 * nub);
}
cplx gfunc(cplx nub, double dm) {
  cplx argcd = nub * dm;
  if (argcd.real() < 75) return std::exp(-argcd) * dm;
  return cplx(0.0, 0.0);
}
cplx h1func(cplx nua, cplx nub, double dm) {
  if (std::abs(nub + nua) < 1.0e-08) return dm;
  cplx argcd = (nua + nub) * dm;
  cplx exqq;
  if (argcd.real() < 40.0)
    exqq = std::exp(-argcd);
  else
    exqq = 0.0;
  return (1.0 - exqq) / (nub + nua);
}
cplx h2func(cplx nua, cplx nub, double dm) {
  if (std::abs(nub - nua) < 1.0e-08) return dm;
  cplx argcd = nua * dm;
  cplx exqp, exqq;
  if (argcd.real() < 40.0)
    exqp = std::exp(-argcd);
  else
    exqp = 0.0;
  argcd = nub * dm;
  if (argcd.real() < 40.0)
    exqq = std::exp(-argcd);
  else
    exqq = 0.0;
  return (exqq - exqp) / (nua - nub);
}

// sregn96.f90:1203-1323
double intijr(RW &S, int i, int j, int m, int typelyr, double om, double om2, double wvno,
              double wvno2) {
#define E(a, b) S.e[(a)-1][(b)-1]
#define EINV(a, b) S.einv[(a)-1][(b)-1]
  cplx gbr[5];
  evalg(S, 0, m, gbr, wvno, om, om2, wvno2);
  const cplx ra = S.ra, rb = S.rb;
  const double *ur = S.ur.data() - 1, *uz = S.uz.data() - 1, *tz = S.tz.data() - 1,
               *tr = S.tr.data() - 1;  // 1-based views
  const double zdm = S.zd[m - 1];
  cplx cintijr, kmpu, kmsu, km1pd, km1sd;
  if (S.iwat[m - 1] == 1) {
    if (typelyr < 0) {
      kmpu = EINV(1, 1) * uz[m + 1] + EINV(1, 2) * tz[m + 1];
      cintijr = E(i, 1) * E(j, 1) * kmpu * kmpu / (2.0 * ra);
    } else if (typelyr == 0) {
      km1pd = EINV(2, 1) * uz[m] + EINV(2, 2) * tz[m];
      kmpu = EINV(1, 1) * uz[m + 1] + EINV(1, 2) * tz[m + 1];
      cplx FA = ffunc(ra, zdm), GA = gfunc(ra, zdm);
      cintijr = E(i, 1) * E(j, 1) * kmpu * kmpu * FA +
                (E(i, 1) * E(j, 2) + E(i, 2) * E(j, 1)) * kmpu * km1pd * GA +
                E(i, 2) * E(j, 2) * km1pd * km1pd * FA;
    } else {
      km1pd = EINV(2, 1) * uz[m] + EINV(2, 2) * tz[m];
      cintijr = E(i, 2) * E(j, 2) * km1pd * km1pd / (2.0 * ra);
    }
  } else {
    if (typelyr < 0) {
      kmpu = EINV(1, 1) * ur[m] + EINV(1, 2) * uz[m] + EINV(1, 3) * tz[m] + EINV(1, 4) * tr[m];
      kmsu = EINV(2, 1) * ur[m] + EINV(2, 2) * uz[m] + EINV(2, 3) * tz[m] + EINV(2, 4) * tr[m];
      cintijr = E(i, 1) * E(j, 1) * kmpu * kmpu / (2.0 * ra) +
                (E(i, 1) * E(j, 2) + E(i, 2) * E(j, 1)) * kmpu * kmsu / (ra + rb) +
                E(i, 2) * E(j, 2) * kmsu * kmsu / (2.0 * rb);
    } else if (typelyr == 0) {
      km1pd = EINV(3, 1) * ur[m] + EINV(3, 2) * uz[m] + EINV(3, 3) * tz[m] + EINV(3, 4) * tr[m];
      km1sd = EINV(4, 1) * ur[m] + EINV(4, 2) * uz[m] + EINV(4, 3) * tz[m] + EINV(4, 4) * tr[m];
      kmpu = EINV(1, 1) * ur[m + 1] + EINV(1, 2) * uz[m + 1] + EINV(1, 3) * tz[m + 1] +
             EINV(1, 4) * tr[m + 1];
      kmsu = EINV(2, 1) * ur[m + 1] + EINV(2, 2) * uz[m + 1] + EINV(2, 3) * tz[m + 1] +
             EINV(2, 4) * tr[m + 1];
      cplx FA = ffunc(ra, zdm), GA = gfunc(ra, zdm), FB = ffunc(rb, zdm), GB = gfunc(rb, zdm);
      cplx H1 = h1func(ra, rb, zdm), H2 = h2func(ra, rb, zdm);
      cintijr = E(i, 1) * E(j, 1) * kmpu * kmpu * FA + E(i, 3) * E(j, 3) * km1pd * km1pd * FA +
                E(i, 2) * E(j, 2) * kmsu * kmsu * FB + E(i, 4) * E(j, 4) * km1sd * km1sd * FB +
                H1 * ((E(i, 1) * E(j, 2) + E(i, 2) * E(j, 1)) * kmpu * kmsu +
                      (E(i, 3) * E(j, 4) + E(i, 4) * E(j, 3)) * km1pd * km1sd) +
                H2 * ((E(i, 1) * E(j, 4) + E(i, 4) * E(j, 1)) * kmpu * km1sd +
                      (E(i, 2) * E(j, 3) + E(i, 3) * E(j, 2)) * km1pd * kmsu) +
                GA * (E(i, 1) * E(j, 3) + E(i, 3) * E(j, 1)) * kmpu * km1pd +
                GB * (E(i, 2) * E(j, 4) + E(i, 4) * E(j, 2)) * kmsu * km1sd;
    } else {
      km1pd = EINV(3, 1) * ur[m] + EINV(3, 2) * uz[m] + EINV(3, 3) * tz[m] + EINV(3, 4) * tr[m];
      km1sd = EINV(4, 1) * ur[m] + EINV(4, 2) * uz[m] + EINV(4, 3) * tz[m] + EINV(4, 4) * tr[m];
      cintijr = E(i, 3) * E(j, 3) * km1pd * km1pd / (2.0 * ra) +
                (E(i, 3) * E(j, 4) + E(i, 4) * E(j, 3)) * km1pd * km1sd / (ra + rb) +
                E(i, 4) * E(j, 4) * km1sd * km1sd / (2.0 * rb);
    }
  }
  return cintijr.real();
#undef E
#undef EINV
}

// sregn96.f90:1537-1589
void getmat(RW &S, int m, double wvno, double om, double &a12, double &a14, double &a21,
            double &a23, double &ah, double &av, double &bh, double &bv, double &eta,
            double &rho, double &TA, double &TC, double &TF, double &TL, double &TN, int iwat) {
  const double za = S.za[m - 1], zb = S.zb[m - 1], zr = S.zrho[m - 1];
  if (iwat == 1) {
    ah = za;
    av = za;
    bh = 0.0;
    bv = 0.0;
    rho = zr;
    eta = 1.0;
    TL = 0.0;
    TN = 0.0;
    TC = zr * za * za;
    TA = zr * za * za;
    TF = TA - 2. * TN;
    a12 = -(wvno * wvno - om * om / (ah * ah)) / (rho * om * om);
  } else {
    ah = za;
    av = za;
    bh = zb;
    bv = zb;
    rho = zr;
    eta = 1.0;
    TL = zr * zb * zb;
    TN = zr * zb * zb;
    TC = zr * za * za;
    TA = zr * za * za;
    TF = TA - 2. * TN;
    a12 = -wvno;
    a14 = 1.0 / TL;
    a21 = wvno * TF / TC;
    a23 = 1.0 / TC;
  }
}

// sregn96.f90:1436-1535
void getdcdh(RW &S, double om2, double wvno, double wvno2, double fac) {
  const int mmax = S.mmax;
  const double *ur = S.ur.data() - 1, *uz = S.uz.data() - 1, *tz = S.tz.data() - 1,
               *tr = S.tr.data() - 1, *xmu = S.xmu.data() - 1, *xlam = S.xlam.data() - 1,
               *zrho = S.zrho.data() - 1;
  const int *iwat = S.iwat.data() - 1;
  for (int m = 1; m <= mmax; m++) {
    double tuz = uz[m], ttz = tz[m], ttr = tr[m], tur;
    if (iwat[m] == 1)
      tur = -wvno * ttz / (zrho[m] * om2);
    else
      tur = ur[m];
    double gfac1, gfac2, gfac3, gfac4, gfac5, gfac6;
    if (m == 1) {
      double drho = zrho[1] - 0.0, dmu = xmu[1] - 0.0, dlm = xlam[1] - 0.0;
      double dl2mu = dlm + dmu + dmu;
      double xl2mp = xlam[m] + xmu[m] + xmu[m];
      double duzdzp = (ttz + wvno * xlam[m] * tur) / xl2mp;
      double durdzp;
      if (iwat[m] == 1)
        durdzp = wvno * tuz;
      else
        durdzp = (ttr / xmu[m]) - wvno * tuz;
      double drur2 = tur * tur * drho;
      double dlur2 = tur * tur * dl2mu;
      gfac1 = om2 * drho * tuz * tuz;
      gfac2 = om2 * drur2;
      gfac3 = -wvno2 * dmu * tuz * tuz;
      gfac4 = -wvno2 * dlur2;
      gfac5 = (xl2mp * duzdzp * duzdzp);
      gfac6 = (xmu[m] * durdzp * durdzp);
    } else {
      double drho = zrho[m] - zrho[m - 1];
      double dmu = xmu[m] - xmu[m - 1];
      double dlm = xlam[m] - xlam[m - 1];
      double dl2mu = dlm + dmu + dmu;
      double xl2mp = xlam[m] + xmu[m] + xmu[m];
      double xl2mm = xlam[m - 1] + xmu[m - 1] + xmu[m - 1];
      double duzdzp = (ttz + wvno * xlam[m] * tur) / xl2mp;
      double durdzp, durdzm, drur2, dlur2, duzdzm;
      if (xmu[m] == 0.0)
        durdzp = wvno * tuz;
      else
        durdzp = (ttr / xmu[m]) - wvno * tuz;
      if (xmu[m - 1] == 0.0)
        durdzm = wvno * tuz;
      else
        durdzm = (ttr / xmu[m - 1]) - wvno * tuz;
      if (iwat[m - 1] == 1 && iwat[m] == 0) {
        double URB = -wvno * tz[m] / (zrho[m - 1] * om2);
        drur2 = tur * tur * zrho[m] - URB * URB * zrho[m - 1];
        dlur2 = tur * tur * xl2mp - URB * URB * xl2mm;
        duzdzm = (ttz + wvno * xlam[m - 1] * URB) / xlam[m - 1];
      } else if (iwat[m - 1] == 1 && iwat[m] == 1) {
        double URB = -wvno * tz[m] / (zrho[m - 1] * om2);
        drur2 = tur * tur * zrho[m] - URB * URB * zrho[m - 1];
        dlur2 = tur * tur * xl2mp - URB * URB * xl2mm;
        duzdzm = (ttz + wvno * xlam[m - 1] * URB) / xl2mm;
      } else {
        drur2 = tur * tur * drho;
        dlur2 = tur * tur * dl2mu;
        duzdzm = (ttz + wvno * xlam[m - 1] * tur) / xl2mm;
      }
      gfac1 = om2 * drho * tuz * tuz;
      gfac2 = om2 * drur2;
      gfac3 = -wvno2 * dmu * tuz * tuz;
      gfac4 = -wvno2 * dlur2;
      gfac5 = (xl2mp * duzdzp * duzdzp - xl2mm * duzdzm * duzdzm);
      gfac6 = (xmu[m] * durdzp * durdzp - xmu[m - 1] * durdzm * durdzm);
    }
    double dfac = fac * (gfac1 + gfac2 + gfac3 + gfac4 + gfac5 + gfac6);
    if (std::fabs(dfac) < 1.0e-38) dfac = 0.0;
    S.dcdh[m - 1] = dfac;
  }
}

// sregn96.f90:1065-1201
void energy(RW &S, double om, double wvno) {
  const int mmax = S.mmax;
  S.sumi0 = S.sumi1 = S.sumi2 = S.sumi3 = 0.0;
  double c = om / wvno, om2 = om * om, wvno2 = wvno * wvno;
  for (int m = 1; m <= mmax; m++) {
    double a12 = 0, a14 = 0, a21 = 0, a23 = 0, ah, av, bh, bv, eta, rho, TA, TC, TF, TL, TN;
    getmat(S, m, wvno, om, a12, a14, a21, a23, ah, av, bh, bv, eta, rho, TA, TC, TF, TL, TN,
           S.iwat[m - 1]);
    int typelyr = (m == mmax) ? 1 : 0;
    double INT11 = intijr(S, 1, 1, m, typelyr, om, om2, wvno, wvno2);
    double INT13 = intijr(S, 1, 3, m, typelyr, om, om2, wvno, wvno2);
    double INT22 = intijr(S, 2, 2, m, typelyr, om, om2, wvno, wvno2);
    double INT24 = intijr(S, 2, 4, m, typelyr, om, om2, wvno, wvno2);
    double INT33 = intijr(S, 3, 3, m, typelyr, om, om2, wvno, wvno2);
    double INT44 = intijr(S, 4, 4, m, typelyr, om, om2, wvno, wvno2);
    if (S.iwat[m - 1] == 1) {
      double w = wvno / (rho * om2);
      double URUR = INT22 * w * w;
      double UZUZ = INT11;
      double URDUZ = -(wvno / (rho * om2)) * a12 * INT22;
      double DUZDUZ = a12 * a12 * INT22;
      S.sumi0 += rho * (URUR + UZUZ);
      S.sumi1 += TA * URUR;
      S.sumi2 -= TF * URDUZ;
      S.sumi3 += TC * DUZDUZ;
      double facah = rho * ah * (URUR - 2. * eta * URDUZ / wvno);
      double facav = rho * av * DUZDUZ / wvno2;
      S.dcda[m - 1] = facah + facav;
      double facr = -0.5 * c * c * (URUR + UZUZ);
      S.dcdr[m - 1] = 0.5 * (av * facav + ah * facah) / rho + facr;
      S.ur[m - 1] = -wvno * S.tz[m - 1] / (rho * om2);
    } else {
      double URUR = INT11;
      double UZUZ = INT22;
      double DURDUR = a12 * a12 * INT22 + 2. * a12 * a14 * INT24 + a14 * a14 * INT44;
      double DUZDUZ = a21 * a21 * INT11 + 2. * a21 * a23 * INT13 + a23 * a23 * INT33;
      double URDUZ = a21 * INT11 + a23 * INT13;
      double UZDUR = a12 * INT22 + a14 * INT24;
      S.sumi0 += rho * (URUR + UZUZ);
      S.sumi1 += TL * UZUZ + TA * URUR;
      S.sumi2 += TL * UZDUR - TF * URDUZ;
      S.sumi3 += TL * DURDUR + TC * DUZDUZ;
      double facah = rho * ah * (URUR - 2. * eta * URDUZ / wvno);
      double facav = rho * av * DUZDUZ / wvno2;
      double facbh = 0.0;
      double facbv =
          rho * bv * (UZUZ + 2. * UZDUR / wvno + DURDUR / wvno2 + 4. * eta * URDUZ / wvno);
      S.dcda[m - 1] = facah + facav;
      S.dcdb[m - 1] = facbv + facbh;
      double facr = -0.5 * c * c * (URUR + UZUZ);
      S.dcdr[m - 1] = 0.5 * (av * facav + ah * facah + bv * facbv) / rho + facr;
    }
  }
  S.flagr = om2 * S.sumi0 - wvno2 * S.sumi1 - 2.0 * wvno * S.sumi2 - S.sumi3;
  S.ugr = (wvno * S.sumi1 + S.sumi2) / (om * S.sumi0);
  S.are = wvno / (2.0 * om * S.ugr * S.sumi0);
  double fac = S.are * c / wvno2;
  for (int m = 0; m < mmax; m++) {
    S.dcda[m] = S.dcda[m] / (S.ugr * S.sumi0);
    S.dcdb[m] = S.dcdb[m] / (S.ugr * S.sumi0);
    S.dcdr[m] = S.dcdr[m] / (S.ugr * S.sumi0);
  }
  getdcdh(S, om2, wvno, wvno2, fac);
}

// sregn96.f90:1591-1635
void sprayl(RW &S, double om, double c, double &csph, double &usph, double ugr) {
  double ar = 6371.0;
  double x = c / (2. * ar * om);
  double tm = std::sqrt(1. + x * x);
  double tm3 = tm * tm * tm;
  for (int i = 0; i < S.mmax; i++) {
    S.dcda[i] = S.dcda[i] * S.vtp[i] / tm3;
    S.dcdb[i] = S.dcdb[i] * S.vtp[i] / tm3;
    S.dcdh[i] = S.dcdh[i] * S.dtp[i] / tm3;
    S.dcdr[i] = S.dcdr[i] * S.rtp[i] / tm3;
  }
  usph = ugr * tm;
  csph = c / tm;
}

// common prologue of sregn96/sregnpu (:1657-1683, :1776-1801)
void setup(RW &S, const float *thk, const float *vp, const float *vs, const float *rhom,
           int nlayer, int iflsph) {
  S.alloc(nlayer);
  for (int i = 0; i < nlayer; i++) {
    S.zb[i] = (double)vs[i];
    S.za[i] = (double)vp[i];
    S.zrho[i] = (double)rhom[i];
    S.zd[i] = (double)thk[i];
  }
  S.allfluid = true;
  for (int i = 0; i < nlayer; i++) {
    if (S.zb[i] > 0.0) {
      S.allfluid = false;
      S.iwat[i] = 0;
    } else
      S.iwat[i] = 1;
  }
  if (iflsph > 0) bldsph(S);
  for (int i = 0; i < nlayer; i++) {
    S.xmu[i] = S.zrho[i] * S.zb[i] * S.zb[i];
    S.xlam[i] = S.zrho[i] * S.za[i] * S.za[i] - 2 * S.xmu[i];
  }
}

}  // namespace

void sregn96(const float *thk, const float *vp, const float *vs, const float *rhom, int nlayer,
             double *t, double *cp, double *cg, double *dispu, double *dispw, double *stressu,
             double *stressw, double *dc2da, double *dc2db, double *dc2dh, double *dc2dr,
             int iflsph) {
  RW S;
  setup(S, thk, vp, vs, rhom, nlayer, iflsph);
  const int mmax = S.mmax;
  double twopi = 2.0 * PI32;
  double omega = twopi / *t;
  double c = *cp;
  S.ugr = c;
  double wvno = omega / c;
  svfunc(S, omega, wvno);
  energy(S, omega, wvno);
  if (std::fabs(S.uu0[0]) < 1.0e-36) S.uu0[0] = 0.0;
  if (std::fabs(S.uu0[2]) < 1.0e-36) S.uu0[2] = 0.0;
  if (std::fabs(c) < 1.0e-36) c = 0.0;
  if (std::fabs(S.ugr) < 1.0e-36) S.ugr = 0.0;
  double csph, usph;
  if (iflsph > 0) {
    sprayl(S, omega, c, csph, usph, S.ugr);
  } else {
    csph = c;
    usph = S.ugr;
  }
  for (int i = 1; i <= mmax - 1; i++) {
    double sums = 0.0;
    for (int j = i + 1; j <= mmax; j++) sums += S.dcdh[j - 1];
    S.dcdh[i - 1] = sums;
  }
  S.dcdh[mmax - 1] = 0.0;
  for (int i = 0; i < mmax; i++) {
    dc2da[i] = S.dcda[i];
    dc2db[i] = S.dcdb[i];
    dc2dr[i] = S.dcdr[i];
    dc2dh[i] = S.dcdh[i];
    dispu[i] = S.ur[i];
    dispw[i] = S.uz[i];
    stressu[i] = S.tr[i];
    stressw[i] = S.tz[i];
  }
  *cp = csph;
  *cg = usph;
}

void sregnpu(const float *thk, const float *vp, const float *vs, const float *rhom, int nlayer,
             double *t, double *cp, double *cg, double *dispu, double *dispw, double *stressu,
             double *stressw, double *t1, double *cp1, double *t2, double *cp2, double *dc2da,
             double *dc2db, double *dc2dh, double *dc2dr, double *du2da, double *du2db,
             double *du2dh, double *du2dr, int iflsph, bool stale_first_term) {
  RW S;
  setup(S, thk, vp, vs, rhom, nlayer, iflsph);
  const int mmax = S.mmax;
  std::vector<double> a1(mmax), a2(mmax), b1(mmax), b2(mmax), h1(mmax), h2(mmax), r1(mmax),
      r2(mmax);
  double twopi = 2.0 * PI32;
  // solve at (t, cp)
  double omega = twopi / *t;
  double c = *cp;
  double wvno = omega / c;
  svfunc(S, omega, wvno);
  energy(S, omega, wvno);
  *cg = S.ugr;
  for (int i = 0; i < mmax; i++) {
    dc2da[i] = S.dcda[i];
    dc2db[i] = S.dcdb[i];
    dc2dr[i] = S.dcdr[i];
    dc2dh[i] = S.dcdh[i];
    dispu[i] = S.ur[i];
    dispw[i] = S.uz[i];
    stressu[i] = S.tr[i];
    stressw[i] = S.tz[i];
  }
  // solve at (t1, cp1)
  omega = twopi / *t1;
  c = *cp1;
  wvno = omega / c;
  svfunc(S, omega, wvno);
  energy(S, omega, wvno);
  for (int i = 0; i < mmax; i++) {
    a1[i] = S.dcda[i];
    b1[i] = S.dcdb[i];
    r1[i] = S.dcdr[i];
    h1[i] = S.dcdh[i];
  }
  // solve at (t2, cp2)
  omega = twopi / *t2;
  c = *cp2;
  wvno = omega / c;
  svfunc(S, omega, wvno);
  energy(S, omega, wvno);
  for (int i = 0; i < mmax; i++) {
    a2[i] = S.dcda[i];
    b2[i] = S.dcdb[i];
    r2[i] = S.dcdr[i];
    h2[i] = S.dcdh[i];
  }
  // dU/dm, :1839-1844.  The reference's first term uses the MODULE arrays dcda.. which at this
  // point hold the T2 solve (stale); the "fixed" form uses the T solve (dc2da..).
  double uc1 = *cg / *cp;
  for (int i = 0; i < mmax; i++) {
    double fa = stale_first_term ? S.dcda[i] : dc2da[i];
    double fb = stale_first_term ? S.dcdb[i] : dc2db[i];
    double fr = stale_first_term ? S.dcdr[i] : dc2dr[i];
    double fh = stale_first_term ? S.dcdh[i] : dc2dh[i];
    du2da[i] = uc1 * (2.0 - uc1) * fa - uc1 * uc1 * *t * (a2[i] - a1[i]) / (*t2 - *t1);
    du2db[i] = uc1 * (2.0 - uc1) * fb - uc1 * uc1 * *t * (b2[i] - b1[i]) / (*t2 - *t1);
    du2dr[i] = uc1 * (2.0 - uc1) * fr - uc1 * uc1 * *t * (r2[i] - r1[i]) / (*t2 - *t1);
    du2dh[i] = uc1 * (2.0 - uc1) * fh - uc1 * uc1 * *t * (h2[i] - h1[i]) / (*t2 - *t1);
  }
  if (iflsph > 0) {
    double ar = 6371.0;
    omega = twopi / *t;
    double x = *cp / (2. * ar * omega);
    double tm = std::sqrt(1. + x * x);
    double y = 0.5 / (ar * omega);
    double tm1 = y * y / tm;
    double tm3 = tm * tm * tm;
    for (int i = 0; i < mmax; i++) {
      du2da[i] = (tm * du2da[i] + *cg * *cp * dc2da[i] * tm1) * S.vtp[i];
      du2db[i] = (tm * du2db[i] + *cg * *cp * dc2db[i] * tm1) * S.vtp[i];
      du2dr[i] = (tm * du2dr[i] + *cg * *cp * dc2dr[i] * tm1) * S.rtp[i];
      du2dh[i] = (tm * du2dh[i] + *cg * *cp * dc2dh[i] * tm1) * S.dtp[i];
      dc2da[i] = dc2da[i] / tm3 * S.vtp[i];
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
