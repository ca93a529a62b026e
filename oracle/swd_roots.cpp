// ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).
// PARITY UNPINNED: the reference ships no golden vectors and cannot be compiled
// in this image (no gfortran / FFTW3); this file is a CPU restatement written by
// reading the reference Fortran.  See oracle/README.md.
//
// Restates the phase-velocity root search of
//   /root/reference/src/SWD/surfdisp96.f
//     surfdisp96 :54-368   gtsolh :375-396   getsol :398-491   sphere :495-564
//     nevill :568-687      half :689-701     dltar :705-723    dltar1 :727-787
//     dltar4 :791-891      var :894-1011     normc :1015-1040  dnka :1044-1088
// including its float32 rounding points (model arrays, start value, betmx,
// output c) and its static `del1st` / `dhalf` state.
#include "oracle.hpp"
#include <cmath>
#include <cstring>
#include <vector>

namespace oracle {

namespace {

inline double dsign(double a, double b) { return std::signbit(b) ? -std::fabs(a) : std::fabs(a); }

struct Mod32 {
  int mmax = 0, llw = 1;
  std::vector<float> d, a, b, rho, rtp, dtp, btp;
  float dhalf = 0.f;  // `save dhalf`, surfdisp96.f:526
};

// surfdisp96.f:375-396 — everything here is REAL*4 (implicit typing).
float gtsolh(float a, float b) {
  float c = 0.95f * b;
  for (int i = 0; i < 5; i++) {
    float gamma = b / a;
    float kappa = c / b;
    float k2 = kappa * kappa;
    float gk2 = (gamma * kappa) * (gamma * kappa);
    float fac1 = std::sqrt(1.0f - gk2);
    float fac2 = std::sqrt(1.0f - k2);
    float fr = (2.0f - k2) * (2.0f - k2) - 4.0f * fac1 * fac2;
    float frp = -4.0f * (2.0f - k2) * kappa + 4.0f * fac2 * gamma * gamma * kappa / fac1 +
                4.0f * fac1 * kappa / fac2;
    frp = frp / b;
    c = c - fr / frp;
  }
  return c;
}

// surfdisp96.f:495-564
void sphere(int ifunc, int iflag, Mod32 &M) {
  const int mmax = M.mmax;
  double ar = 6370.0, dr = 0.0, r0 = ar;
  M.d[mmax - 1] = 1.0f;
  if (iflag == 0) {
    for (int i = 0; i < mmax; i++) {
      M.dtp[i] = M.d[i];
      M.rtp[i] = M.rho[i];
    }
    for (int i = 0; i < mmax; i++) {
      dr = dr + (double)M.d[i];
      double r1 = ar - dr;
      double z0 = ar * std::log(ar / r0);
      double z1 = ar * std::log(ar / r1);
      M.d[i] = (float)(z1 - z0);
      double tmp = (ar + ar) / (r0 + r1);
      M.a[i] = (float)((double)M.a[i] * tmp);
      M.b[i] = (float)((double)M.b[i] * tmp);
      M.btp[i] = (float)tmp;
      r0 = r1;
    }
    M.dhalf = M.d[mmax - 1];
  } else {
    M.d[mmax - 1] = M.dhalf;
    for (int i = 0; i < mmax; i++) {
      if (ifunc == 1) {
        // btp(i)**(-5): REAL*4 base with an integer exponent = repeated multiplication
        const float x = M.btp[i];
        const float x5 = (((x * x) * x) * x) * x;
        M.rho[i] = M.rtp[i] * (1.0f / x5);
      }
      else if (ifunc == 2)
        M.rho[i] = M.rtp[i] * std::pow(M.btp[i], -2.275f);  // REAL*4 pow
    }
  }
  M.d[mmax - 1] = 0.0f;
}

// surfdisp96.f:727-787 — Love (SH) period equation, Haskell 2-vector from the half-space up.
double dltar1(double wvno, double omega, const Mod32 &M) {
  const int mmax = M.mmax, llw = M.llw;
  double beta1 = (double)M.b[mmax - 1];
  double rho1 = (double)M.rho[mmax - 1];
  double xkb = omega / beta1;
  double wvnop = wvno + xkb;
  double wvnom = std::fabs(wvno - xkb);
  double rb = std::sqrt(wvnop * wvnom);
  double e1 = rho1 * rb;
  double e2 = 1.0 / (beta1 * beta1);
  for (int m = mmax - 1; m >= llw; m--) {  // Fortran m = mmax-1 .. llw
    const int i = m - 1;
    beta1 = (double)M.b[i];
    rho1 = (double)M.rho[i];
    double xmu = rho1 * beta1 * beta1;
    xkb = omega / beta1;
    wvnop = wvno + xkb;
    wvnom = std::fabs(wvno - xkb);
    rb = std::sqrt(wvnop * wvnom);
    double q = (double)M.d[i] * rb;
    double sinq, y, z, cosq;
    if (wvno < xkb) {
      sinq = std::sin(q);
      y = sinq / rb;
      z = -rb * sinq;
      cosq = std::cos(q);
    } else if (wvno == xkb) {
      cosq = 1.0;
      y = (double)M.d[i];
      z = 0.0;
    } else {
      double fac = 0.0;
      if (q < 16) fac = std::exp(-2.0 * q);
      cosq = (1.0 + fac) * 0.5;
      sinq = (1.0 - fac) * 0.5;
      y = sinq / rb;
      z = rb * sinq;
    }
    double e10 = e1 * cosq + e2 * xmu * z;
    double e20 = e1 * y / xmu + e2 * cosq;
    double xnor = std::fabs(e10);
    double ynor = std::fabs(e20);
    if (ynor > xnor) xnor = ynor;
    if (xnor < 1.e-40) xnor = 1.0;
    e1 = e10 / xnor;
    e2 = e20 / xnor;
  }
  return e1;
}

struct VarOut {
  double w, cosp, exa, a0, cpcq, cpy, cpz, cqw, cqx, xy, xz, wy, wz;
};

// surfdisp96.f:894-1011
void var(double p, double q, double ra, double rb, double wvno, double xka, double xkb,
         double dpth, VarOut &o) {
  double pex = 0.0, sex = 0.0;
  double sinp, w = 0, x = 0, cosp = 0, sinq, y = 0, z = 0, cosq = 0;
  if (wvno < xka) {
    sinp = std::sin(p);
    w = sinp / ra;
    x = -ra * sinp;
    cosp = std::cos(p);
  } else if (wvno == xka) {
    cosp = 1.0;
    w = dpth;
    x = 0.0;
  } else {
    pex = p;
    double fac = 0.0;
    if (p < 16) fac = std::exp(-2.0 * p);
    cosp = (1.0 + fac) * 0.5;
    sinp = (1.0 - fac) * 0.5;
    w = sinp / ra;
    x = ra * sinp;
  }
  if (wvno < xkb) {
    sinq = std::sin(q);
    y = sinq / rb;
    z = -rb * sinq;
    cosq = std::cos(q);
  } else if (wvno == xkb) {
    cosq = 1.0;
    y = dpth;
    z = 0.0;
  } else {
    sex = q;
    double fac = 0.0;
    if (q < 16) fac = std::exp(-2.0 * q);
    cosq = (1.0 + fac) * 0.5;
    sinq = (1.0 - fac) * 0.5;
    y = sinq / rb;
    z = rb * sinq;
  }
  o.exa = pex + sex;
  o.a0 = 0.0;
  if (o.exa < 60.0) o.a0 = std::exp(-o.exa);
  o.cpcq = cosp * cosq;
  o.cpy = cosp * y;
  o.cpz = cosp * z;
  o.cqw = cosq * w;
  o.cqx = cosq * x;
  o.xy = x * y;
  o.xz = x * z;
  o.wy = w * y;
  o.wz = w * z;
  o.w = w;
  o.cosp = cosp;
  // (the rescaled cosq,y,z of :1005-1010 are locals that never leave `var`)
}

// surfdisp96.f:1044-1088 — Dunkin 5x5 compound matrix, ca[row][col] 0-based.
void dnka(double ca[5][5], double wvno2, double gam, double gammk, double rho, const VarOut &v) {
  const double one = 1.0, two = 2.0;
  double gamm1 = gam - one;
  double twgm1 = gam + gamm1;
  double gmgmk = gam * gammk;
  double gmgm1 = gam * gamm1;
  double gm1sq = gamm1 * gamm1;
  double rho2 = rho * rho;
  double a0pq = v.a0 - v.cpcq;
  ca[0][0] = v.cpcq - two * gmgm1 * a0pq - gmgmk * v.xz - wvno2 * gm1sq * v.wy;
  ca[0][1] = (wvno2 * v.cpy - v.cqx) / rho;
  ca[0][2] = -(twgm1 * a0pq + gammk * v.xz + wvno2 * gamm1 * v.wy) / rho;
  ca[0][3] = (v.cpz - wvno2 * v.cqw) / rho;
  ca[0][4] = -(two * wvno2 * a0pq + v.xz + wvno2 * wvno2 * v.wy) / rho2;
  ca[1][0] = (gmgmk * v.cpz - gm1sq * v.cqw) * rho;
  ca[1][1] = v.cpcq;
  ca[1][2] = gammk * v.cpz - gamm1 * v.cqw;
  ca[1][3] = -v.wz;
  ca[1][4] = ca[0][3];
  ca[3][0] = (gm1sq * v.cpy - gmgmk * v.cqx) * rho;
  ca[3][1] = -v.xy;
  ca[3][2] = gamm1 * v.cpy - gammk * v.cqx;
  ca[3][3] = ca[1][1];
  ca[3][4] = ca[0][1];
  ca[4][0] = -(two * gmgmk * gm1sq * a0pq + gmgmk * gmgmk * v.xz + gm1sq * gm1sq * v.wy) * rho2;
  ca[4][1] = ca[3][0];
  ca[4][2] = -(gammk * gamm1 * twgm1 * a0pq + gam * gammk * gammk * v.xz + gamm1 * gm1sq * v.wy) * rho;
  ca[4][3] = ca[1][0];
  ca[4][4] = ca[0][0];
  double t = -two * wvno2;
  ca[2][0] = t * ca[4][2];
  ca[2][1] = t * ca[3][2];
  ca[2][2] = v.a0 + two * (v.cpcq - ca[0][0]);
  ca[2][3] = t * ca[1][2];
  ca[2][4] = t * ca[0][2];
}

// surfdisp96.f:1015-1040
void normc5(double ee[5], double &ex) {
  ex = 0.0;
  double t1 = 0.0;
  for (int i = 0; i < 5; i++)
    if (std::fabs(ee[i]) > t1) t1 = std::fabs(ee[i]);
  if (t1 < 1.e-40) t1 = 1.0;
  for (int i = 0; i < 5; i++) ee[i] = ee[i] / t1;
  ex = std::log(t1);
}

// surfdisp96.f:791-891 — Rayleigh (P-SV) period equation, Dunkin 5-vector from the half-space up.
double dltar4(double wvno, double omga, const Mod32 &M) {
  const int mmax = M.mmax, llw = M.llw;
  double e[5], ee[5], ca[5][5];
  double omega = omga;
  if (omega < 1.0e-4) omega = 1.0e-4;
  double wvno2 = wvno * wvno;
  double xka = omega / (double)M.a[mmax - 1];
  double xkb = omega / (double)M.b[mmax - 1];
  double wvnop = wvno + xka;
  double wvnom = std::fabs(wvno - xka);
  double ra = std::sqrt(wvnop * wvnom);
  wvnop = wvno + xkb;
  wvnom = std::fabs(wvno - xkb);
  double rb = std::sqrt(wvnop * wvnom);
  double t = (double)M.b[mmax - 1] / omega;
  double gammk = 2.0 * t * t;
  double gam = gammk * wvno2;
  double gamm1 = gam - 1.0;
  double rho1 = (double)M.rho[mmax - 1];
  e[0] = rho1 * rho1 * (gamm1 * gamm1 - gam * gammk * ra * rb);
  e[1] = -rho1 * ra;
  e[2] = rho1 * (gamm1 - gammk * ra * rb);
  e[3] = rho1 * rb;
  e[4] = wvno2 - ra * rb;
  VarOut v;
  for (int m = mmax - 1; m >= llw; m--) {
    const int i = m - 1;
    xka = omega / (double)M.a[i];
    xkb = omega / (double)M.b[i];
    t = (double)M.b[i] / omega;
    gammk = 2.0 * t * t;
    gam = gammk * wvno2;
    wvnop = wvno + xka;
    wvnom = std::fabs(wvno - xka);
    ra = std::sqrt(wvnop * wvnom);
    wvnop = wvno + xkb;
    wvnom = std::fabs(wvno - xkb);
    rb = std::sqrt(wvnop * wvnom);
    double dpth = (double)M.d[i];
    rho1 = (double)M.rho[i];
    double p = ra * dpth;
    double q = rb * dpth;
    var(p, q, ra, rb, wvno, xka, xkb, dpth, v);
    dnka(ca, wvno2, gam, gammk, rho1, v);
    for (int ii = 0; ii < 5; ii++) {
      double cr = 0.0;
      for (int j = 0; j < 5; j++) cr = cr + e[j] * ca[j][ii];
      ee[ii] = cr;
    }
    double exa;
    normc5(ee, exa);
    for (int ii = 0; ii < 5; ii++) e[ii] = ee[ii];
  }
  if (llw != 1) {
    // water layer on top, surfdisp96.f:870-886
    xka = omega / (double)M.a[0];
    wvnop = wvno + xka;
    wvnom = std::fabs(wvno - xka);
    ra = std::sqrt(wvnop * wvnom);
    double dpth = (double)M.d[0];
    rho1 = (double)M.rho[0];
    double p = ra * dpth;
    double znul = 1.0e-05;
    var(p, znul, ra, znul, wvno, xka, znul, dpth, v);
    double w0 = -rho1 * v.w;
    return v.cosp * e[0] + w0 * e[1];
  }
  return e[0];
}

struct Counters {
  long evals = 0;
};

struct Ctx {
  Mod32 M;
  double del1st = 0.0;  // `save del1st`, surfdisp96.f:423
  Counters cnt;
};

// surfdisp96.f:705-723
double dltar(double wvno, double omega, int kk, Ctx &C) {
  C.cnt.evals++;
  if (kk == 1) return dltar1(wvno, omega, C.M);
  return dltar4(wvno, omega, C.M);
}

// surfdisp96.f:689-701
void half(double c1, double c2, double &c3, double &del3, double omega, int ifunc, Ctx &C) {
  c3 = 0.5 * (c1 + c2);
  double wvno = omega / c3;
  del3 = dltar(wvno, omega, ifunc, C);
}

// surfdisp96.f:568-687 — hybrid interval-halving / inverse Neville refinement.
void nevill(double t, double c1, double c2, double del1, double del2, int ifunc, double &cc,
            Ctx &C) {
  const double twopi = 2.0 * 3.141592653589793;
  double x[20], y[20];
  double omega = twopi / t;
  double c3, del3;
  half(c1, c2, c3, del3, omega, ifunc, C);
  int nev = 1;
  int nctrl = 1;
  int m = 1;
  for (;;) {
    nctrl = nctrl + 1;
    if (nctrl >= 100) break;
    if (c3 < std::fmin(c1, c2) || c3 > std::fmax(c1, c2)) {
      nev = 0;
      half(c1, c2, c3, del3, omega, ifunc, C);
    }
    double s13 = del1 - del3;
    double s32 = del3 - del2;
    if (dsign(1.0, del3) * dsign(1.0, del1) < 0.0) {
      c2 = c3;
      del2 = del3;
    } else {
      c1 = c3;
      del1 = del3;
    }
    if (std::fabs(c1 - c2) <= 1.e-6 * c1) break;
    if (dsign(1.0, s13) != dsign(1.0, s32)) nev = 0;
    double ss1 = std::fabs(del1);
    double s1 = (double)0.01f * ss1;  // `0.01*ss1`: REAL*4 literal promoted, :632
    double ss2 = std::fabs(del2);
    double s2 = (double)0.01f * ss2;
    if (s1 > ss2 || s2 > ss1 || nev == 0) {
      half(c1, c2, c3, del3, omega, ifunc, C);
      nev = 1;
      m = 1;
    } else {
      if (nev == 2) {
        x[m] = c3;  // x(m+1)
        y[m] = del3;
      } else {
        x[0] = c1;
        y[0] = del1;
        x[1] = c2;
        y[1] = del2;
        m = 1;
      }
      bool bad = false;
      for (int kk = 1; kk <= m; kk++) {
        int j = m - kk + 1;  // 1-based
        double denom = y[m] - y[j - 1];
        if (std::fabs(denom) < 1.0e-10 * std::fabs(y[m])) {
          bad = true;
          break;
        }
        x[j - 1] = (-y[j - 1] * x[j] + y[m] * x[j - 1]) / denom;
      }
      if (!bad) {
        c3 = x[0];
        double wvno = omega / c3;
        del3 = dltar(wvno, omega, ifunc, C);
        nev = 2;
        m = m + 1;
        if (m > 10) m = 10;
      } else {
        half(c1, c2, c3, del3, omega, ifunc, C);
        nev = 1;
        m = 1;
      }
    }
  }
  cc = c3;
}

// surfdisp96.f:398-491
void getsol(double t1, double &c1, double clow, double dc, double cm, float betmx, int &iret,
            int ifunc, int ifirst, Ctx &C) {
  const double twopi = 2.0 * 3.141592653589793;
  double omega = twopi / t1;
  double wvno = omega / c1;
  double del1 = dltar(wvno, omega, ifunc, C);
  if (ifirst == 1) C.del1st = del1;
  double plmn = dsign(1.0, C.del1st) * dsign(1.0, del1);
  int idir = +1;
  if (ifirst == 1)
    idir = +1;
  else if (plmn >= 0.0)
    idir = +1;
  else
    idir = -1;
  double c2, del2;
  for (;;) {
    if (idir > 0)
      c2 = c1 + dc;
    else
      c2 = c1 - dc;
    if (c2 <= clow) {
      idir = +1;
      c1 = clow;
      continue;  // del1 is NOT re-evaluated (reference behaviour, :467-471)
    }
    omega = twopi / t1;
    wvno = omega / c2;
    del2 = dltar(wvno, omega, ifunc, C);
    if (dsign(1.0, del1) != dsign(1.0, del2)) break;
    c1 = c2;
    del1 = del2;
    if (c1 < cm) {
      iret = -1;
      return;
    }
    if (c1 >= ((double)betmx + dc)) {
      iret = -1;
      return;
    }
  }
  double cn;
  nevill(t1, c1, c2, del1, del2, ifunc, cn, C);
  c1 = cn;
  if (c1 > (double)betmx) {
    iret = -1;
    return;
  }
  iret = 1;
}

}  // namespace

// surfdisp96.f:54-368.  `mode` is 1-based (1 = fundamental).  igr must be 0 (the
// wrapper never passes anything else: surfdisp.cpp:133,163,212,239-241,260,279-281).
void surfdisp96(const float *thkm, const float *vpm, const float *vsm, const float *rhom,
                int nlayer, int iflsph, int iwave, int mode, int igr, int kmax, const double *t,
                double *cg, int *ierr, long *n_evals) {
  (void)igr;
  Ctx C;
  Mod32 &M = C.M;
  const int mmax = nlayer;
  M.mmax = mmax;
  M.d.assign(thkm, thkm + mmax);
  M.a.assign(vpm, vpm + mmax);
  M.b.assign(vsm, vsm + mmax);
  M.rho.assign(rhom, rhom + mmax);
  M.rtp.assign(mmax, 0.f);
  M.dtp.assign(mmax, 0.f);
  M.btp.assign(mmax, 0.f);
  *ierr = 0;
  int idispl = 0, idispr = 0;
  if (iwave == 1)
    idispl = kmax;
  else if (iwave == 2)
    idispr = kmax;
  int iverb[2] = {0, 0};
  const float sone0 = 1.500f;
  const float ddc0 = 0.005f;
  M.llw = 1;
  if (M.b[0] <= 0.0f) M.llw = 2;
  const double one = 1.0e-2;
  if (iflsph == 1) sphere(0, 0, M);
  int jmn = 1, jsol = 1;
  float betmx = -1.e20f, betmn = 1.e20f;
  for (int i = 0; i < mmax; i++) {
    if (M.b[i] > 0.01f && M.b[i] < betmn) {
      betmn = M.b[i];
      jmn = i + 1;
      jsol = 1;
    } else if (M.b[i] <= 0.01f && M.a[i] < betmn) {
      betmn = M.a[i];
      jmn = i + 1;
      jsol = 0;
    }
    if (M.b[i] > betmx) betmx = M.b[i];
  }
  std::vector<double> c(kmax, 0.0), cb(kmax, 0.0);
  for (int ifunc = 1; ifunc <= 2; ifunc++) {
    if (ifunc == 1 && idispl <= 0) continue;
    if (ifunc == 2 && idispr <= 0) continue;
    if (iflsph == 1) sphere(ifunc, 1, M);
    float ddc = ddc0;
    float sone = sone0;
    if (sone < 0.01f) sone = 2.0f;
    double onea = (double)sone;
    float cc1;
    if (jsol == 0)
      cc1 = betmn;
    else
      cc1 = gtsolh(M.a[jmn - 1], M.b[jmn - 1]);
    cc1 = 0.95f * cc1;
    cc1 = 0.90f * cc1;
    double cc = (double)cc1;
    double dc = (double)ddc;
    dc = std::fabs(dc);
    double c1 = cc;
    double cm = cc;
    for (int i = 0; i < kmax; i++) {
      cb[i] = 0.0;
      c[i] = 0.0;
    }
    int ift = 999;
    for (int iq = 1; iq <= mode; iq++) {
      const int is = 1, ie = kmax;
      int k;
      bool failed = false;
      for (k = is; k <= ie; k++) {
        if (k >= ift) {
          failed = true;
          break;
        }
        double t1 = t[k - 1];
        double clow;
        int ifirst;
        if (k == is && iq == 1) {
          c1 = cc;
          clow = cc;
          ifirst = 1;
        } else if (k == is && iq > 1) {
          c1 = c[is - 1] + one * dc;
          clow = c1;
          ifirst = 1;
        } else if (k > is && iq > 1) {
          ifirst = 0;
          clow = c[k - 1] + one * dc;
          c1 = c[k - 2];
          if (c1 < clow) c1 = clow;
        } else {
          ifirst = 0;
          c1 = c[k - 2] - onea * dc;
          clow = cm;
        }
        int iret;
        getsol(t1, c1, clow, dc, cm, betmx, iret, ifunc, ifirst, C);
        if (iret == -1) {
          failed = true;
          break;
        }
        c[k - 1] = c1;
        float cc0 = (float)c[k - 1];  // cg(k) = sngl(c(k)), :302-307
        cg[k - 1] = (double)cc0;
      }
      if (!failed) continue;
      // label 1700 / 1750
      if (iq <= 1) {
        if (iverb[ifunc - 1] == 0) {
          iverb[ifunc - 1] = 1;
          *ierr = 1;
        }
      }
      ift = k;
      for (int i = k; i <= ie; i++) cg[i - 1] = 0.0;
    }
  }
  if (n_evals) *n_evals = C.cnt.evals;
}

}  // namespace oracle
