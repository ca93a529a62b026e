// K2 (Rayleigh) — eigenfunctions, energy integrals, group velocity and phase-velocity Frechet
// kernels for one (model, period, c), one thread each.
//
// Replaces /root/reference/src/SWD/sregn96.f90: svfunc :196-402, up :404-492, dnka :494-650,
// evalg :652-829, varsv :831-915, hska :917-991, down :993-1063, energy :1065-1201,
// intijr :1203-1323, ffunc..h2func :1325-1403, getdcdh :1436-1535, getmat :1537-1589 and the
// suffix-sum of sregn96 :1727-1731.
//
// Re-design (same mathematics, different schedule):
//  * ONE up-sweep stores the normalised Dunkin vectors cd(m,1:5)+exe(m) in thread-local memory
//    (6 doubles/layer); the down-sweep is fused with the eigenfunction assembly, the per-layer
//    energy integrals and the boundary terms of dc/dh, so the Haskell vectors vv(m,:) and the
//    eigenfunctions are never stored (reference: 25 heap arrays per call).
//  * the six intijr integrals of a layer are ONE symmetric bilinear form a_i^T W a_j in the
//    potentials; evalg (E, E^-1) is evaluated once per layer (reference: 6x).
//  * vertical wavenumbers are real or purely imaginary, so varsv is done in real arithmetic.
// Fluid (water) layers follow the iwat branches of dnka / hska / intijr / energy / getdcdh.
#pragma once
#include "common.cuh"
#include "swd_roots.cuh"

namespace rfs {

// varsv (:831-915) for one wavenumber: s = wvno^2 - xk^2 (real)
struct VSV {
  double c, rs, sr, ex;  // cos-like, r*sinh-like, sinh-like/r, exponent
  bool imag;             // vertical wavenumber purely imaginary (oscillatory)
  double r;              // |nu|
  double ri;             // 1/|nu| (0 when nu = 0)
  double e;              // exp(-ex): exp(-2 ex) = e*e and exp(-(ex_p+ex_s)) = e_p*e_s (one exp per half)
};
RFS_DEVINL VSV varsv_half(double s, double zd) {
  VSV o;
  const double small = (double)1.0e-5f;
  const double as = fabs(s);
  const double ri = (as > 0.0) ? rsqrt_pos(as) : 0.0;  // r = |s| rsqrt(|s|), 1/r = rsqrt(|s|)
  o.r = as * ri;
  o.ri = ri;
  if (s >= 0.0) {
    o.imag = false;
    const double pr = o.r * zd;
    o.e = exp_neg(pr);
    const double pfac = (pr < 30.0) ? o.e * o.e : 0.0;
    o.c = 0.5 + pfac * 0.5;
    const double sinp = 0.5 - pfac * 0.5;
    o.rs = o.r * sinp;
    o.sr = (fabs(pr) < small && o.r < small) ? zd : sinp * ri;
    o.ex = pr;
  } else {
    o.imag = true;
    const double pi_ = o.r * zd;
    double sn, cs;
    sincos_cb(pi_, &sn, &cs);
    o.c = cs;
    o.rs = -o.r * sn;
    o.sr = (o.r < small) ? zd : sn * ri;
    o.ex = 0.0;
    o.e = 1.0;
  }
  return o;
}

// max |v_i| on the integer pipe (bit patterns of non-negative doubles order like integers), the
// power of two 2^-k with 2^k <= max < 2^(k+1), and k ln2.  The reference divides the propagated
// vectors by the max itself and adds log(max) to the exponent bookkeeping (normc :1405-1434); a
// power-of-two scale keeps the same bookkeeping exact, costs no divisions and no logarithm.
#define RFS_ABSBITS(v) \
  (((unsigned long long)((unsigned)__double2hiint(v) & 0x7fffffffu) << 32) | (unsigned)__double2loint(v))
RFS_DEVINL double pow2_scale(unsigned long long umax, double &logscale) {
  double t1 = __longlong_as_double((long long)umax);
  if (t1 < 1.e-40) t1 = 1.0;
  const int k = (__double2hiint(t1) >> 20) - 1023;
  logscale = (double)k * 0.6931471805599453;
  return __hiloint2double((1023 - k) << 20, 0);
}

// sregn96.f90 dnka (:494-650), elastic branch: reduced 5x5 compound matrix, unique entries
struct Dnk {
  double c11, c12, c13, c14, c15, c21, c22, c23, c24, c31, c32, c33, c41, c42, c51;
};
RFS_DEVINL Dnk dnka_r(const VSV &P, const VSV &S, double rho, double b, double exa, double wvno,
                      double wvno2, double om2, double iwv, double iwv2, double iom2, double irom2) {
  Dnk o;
  const double a0 = (exa < 60.0) ? P.e * S.e : 0.0;
  const double cpcq = P.c * S.c, cpy = P.c * S.sr, cpz = P.c * S.rs, cqw = S.c * P.sr,
               cqx = S.c * P.rs, xy = P.rs * S.sr, xz = P.rs * S.rs, wy = P.sr * S.sr,
               wz = P.sr * S.rs;
  const float bf = (float)b, rf = (float)rho;
  const double rho2 = (double)__fmul_rn(rf, rf);                        // REAL*4 rho*rho
  const double gam = (double)__fmul_rn(__fmul_rn(2.0f, bf), bf) * wvno2 * iom2;  // 2.0*b*b in REAL*4
  const double gam2 = gam * gam, gamm1 = gam - 1.0, gamm2 = gamm1 * gamm1;
  const double cqww2 = cqw * wvno2, cqxw2 = cqx * iwv2, gg1 = gam * gamm1;
  const double a0c = 2.0 * (a0 - cpcq);
  const double xz2 = xz * iwv2, gxz2 = gam * xz2, g2xz2 = gam2 * xz2;
  const double a0cgg1 = a0c * (gam + gamm1);
  const double wy2 = wy * wvno2, g2wy2 = gamm2 * wy2, g1wy2 = gamm1 * wy2;
  double temp = a0c * gg1 + g2xz2 + g2wy2;
  o.c33 = a0 + temp + temp;
  o.c11 = cpcq - temp;
  o.c12 = (-cqx + wvno2 * cpy) * irom2;
  temp = 0.5 * a0cgg1 + gxz2 + g1wy2;
  o.c13 = wvno * temp * irom2;
  o.c14 = (-cqww2 + cpz) * irom2;
  temp = wvno2 * (a0c + wy2) + xz;
  o.c15 = -temp * (iom2 * iom2) / rho2;
  o.c21 = (-gamm2 * cqw + gam2 * cpz * iwv2) * rho * om2;
  o.c22 = cpcq;
  o.c23 = (gamm1 * cqww2 - gam * cpz) * iwv;
  o.c24 = -wz;
  temp = 0.5 * a0cgg1 * gg1 + gam2 * gxz2 + gamm2 * g1wy2;
  o.c31 = -2.0 * temp * rho * om2 * iwv;
  o.c32 = -wvno * (gam * cqxw2 - gamm1 * cpy) * 2.0;
  o.c41 = (-gam2 * cqxw2 + gamm2 * cpy) * rho * om2;
  o.c42 = -xy;
  temp = gamm2 * (a0c * gam2 + g2wy2) + gam2 * g2xz2;
  o.c51 = -rho2 * om2 * om2 * temp * iwv2;
  return o;
}

RFS_DEVINL cd mk(bool imag, double r) { return imag ? cd(0.0, r) : cd(r, 0.0); }

// ffunc/gfunc/h1func/h2func (:1325-1403)
RFS_DEVINL cd ffunc_d(cd nub, double dm) {
  if (norm2(nub) < 1.0e-16) return cd(dm);
  const cd arg = nub * dm;
  cd exqq = (arg.x < 40.0) ? cexp_b(-2.0 * arg) : cd(0.0);
  return (1.0 - exqq) / (2.0 * nub);
}
RFS_DEVINL cd gfunc_d(cd nub, double dm) {
  const cd arg = nub * dm;
  if (arg.x < 75.0) return cexp_b(-arg) * dm;
  return cd(0.0);
}
RFS_DEVINL cd h1func_d(cd nua, cd nub, double dm) {
  if (cabs(nub + nua) < 1.0e-08) return cd(dm);
  const cd arg = (nua + nub) * dm;
  cd exqq = (arg.x < 40.0) ? cexp_b(-arg) : cd(0.0);
  return (1.0 - exqq) / (nub + nua);
}
RFS_DEVINL cd h2func_d(cd nua, cd nub, double dm) {
  if (cabs(nub - nua) < 1.0e-08) return cd(dm);
  cd arg = nua * dm;
  cd exqp = (arg.x < 40.0) ? cexp_b(-arg) : cd(0.0);
  arg = nub * dm;
  cd exqq = (arg.x < 40.0) ? cexp_b(-arg) : cd(0.0);
  return (exqq - exqp) / (nua - nub);
}

struct Eig4 {
  double ur, uz, tz, tr;
};

// Output of one Rayleigh solve.  kern points at [4][n] doubles with element stride `ks`
// (order: dcda, dcdb, dcdr, dcdh), dcdh already converted to d/d(thickness) (suffix sums).
// SHARED_CD: the up-sweep vectors live in shared memory as cds[(m*6+j)*cstride] (this thread's
// column of a [NMAX*6][blockDim] array: conflict-free); otherwise in thread-local memory.  The varsv
// terms are recomputed in the down-sweep instead of being parked beside them (together 1 152 B per
// thread at NMAX = 8: with 1.8 M threads per launch that spilled 1.7 GB to DRAM per evaluation).
template <int NMAX, bool SHARED_CD = false>
RFS_DEVINL void rayleigh_solve(const SwdModel &M, long long b, double T, double c, double *ugr_out,
                               double *__restrict__ kern, long long ks, double *cds = nullptr,
                               int cstride = 1) {
  const int mmax = M.n;
  const double omega = (2.0 * RFS_PI32) / T;
  const double wvno = omega / c;
  const double wvno2 = wvno * wvno, om2 = omega * omega;
  const double iwv = 1.0 / wvno, iwv2 = iwv * iwv, iomega = 1.0 / omega, iom2 = iomega * iomega;
  const double slow = wvno * iomega;  // 1/c
  // thread-local variant of the up-sweep vectors: [m][0..4] = cd, [m][5] = exe.  The varsv terms are
  // never stored: the down-sweep recomputes them (12 doubles per layer less local-memory traffic)
  double cdl_local[SHARED_CD ? 1 : NMAX * 6];
  double *cdl = SHARED_CD ? cds : cdl_local;
  const int cst = SHARED_CD ? cstride : 1;
#define CDL(i) cdl[(i) * cst]

  // ---------------- up-sweep (:404-492): half-space vector from evalg (:736-768)
  {
    const int m = mmax - 1;
    const double za = M.ld(F_A, m, b), zb = M.ld(F_B, m, b), zr = M.ld(F_RHO, m, b);
    const double xka = omega / za, xkb = omega / zb;
    const double sa = wvno2 - xka * xka, sb = wvno2 - xkb * xkb;
    const cd ra = mk(sa < 0.0, sqrt(fabs(sa))), rb = mk(sb < 0.0, sqrt(fabs(sb)));
    double gam = zb * wvno / omega;
    gam = 2.0 * (gam * gam);
    const double gamm1 = gam - 1.0;
    const cd rarb = ra * rb;
    cd g1 = (zr * zr) * om2 * om2 * (wvno2 * gamm1 * gamm1 - (gam * gam) * rarb);
    cd g2 = -zr * (wvno2 * ra) * om2;
    cd g3 = -zr * (wvno2 * gamm1 - gam * rarb) * om2 * wvno;
    cd g4 = zr * (wvno2 * rb) * om2;
    cd g5 = wvno2 * (wvno2 - rarb);
    const cd den = (-zr * zr * om2 * om2 * wvno2) * rarb;
    const cd q = 0.25 * cinv(den);
    CDL(m * 6 + 0) = (g1 * q).x;
    CDL(m * 6 + 1) = (g2 * q).x;
    CDL(m * 6 + 2) = (g3 * q).x;
    CDL(m * 6 + 3) = (g4 * q).x;
    CDL(m * 6 + 4) = (g5 * q).x;
    CDL(m * 6 + 5) = 0.0;
  }
  double exsum = 0.0;
  for (int m = mmax - 2; m >= 0; m--) {
    const double zb = M.ld(F_B, m, b), zr = M.ld(F_RHO, m, b), zd = M.ld(F_D, m, b);
    const bool wat = !(zb > 0.0);
    const double xka = omega * M.ld(F_IA, m, b), xkb = wat ? 0.0 : omega * M.ld(F_IB, m, b);
    const double irom2 = M.ld(F_IRHO, m, b) * iom2;
    const VSV P = varsv_half(__fma_rn(-xka, xka, wvno2), zd);  // explicit: the down-sweep recomputes it
    VSV S;
    if (wat) {  // fluid layer: no SV wave (varsv :860-880)
      S.c = 1.0;
      S.rs = 0.0;
      S.sr = 0.0;
      S.ex = 0.0;
      S.r = 0.0;
      S.ri = 0.0;
      S.e = 1.0;
      S.imag = false;
    } else {
      S = varsv_half(__fma_rn(-xkb, xkb, wvno2), zd);
    }
    const double d0 = CDL((m + 1) * 6 + 0), d1 = CDL((m + 1) * 6 + 1), d2 = CDL((m + 1) * 6 + 2),
                 d3 = CDL((m + 1) * 6 + 3), d4 = CDL((m + 1) * 6 + 4);
    double n0, n1, n2, n3, n4;
    if (wat) {
      // fluid compound matrix (dnka :555-572): only 9 non-zero entries
      const double dfac = (P.ex > 35.0) ? 0.0 : P.e;
      const double ca12 = -P.rs * irom2, ca21 = -zr * P.sr * om2;
      n0 = d0 * P.c + d1 * ca21;
      n1 = d0 * ca12 + d1 * P.c;
      n2 = d2 * dfac;
      n3 = d3 * P.c + d4 * ca21;
      n4 = d3 * ca12 + d4 * P.c;
    } else {
      const Dnk A = dnka_r(P, S, zr, zb, P.ex + S.ex, wvno, wvno2, om2, iwv, iwv2, iom2, irom2);
      // ee(i) = sum_j cd(m+1,j) ca(j,i); symmetric fill-ins of :620-645
      //   ca(2,5)=c14 ca(3,4)=-2 c23 ca(3,5)=-2 c13 ca(4,3)=-c32/2 ca(4,4)=c22 ca(4,5)=c12
      //   ca(5,2)=c41 ca(5,3)=-c31/2 ca(5,4)=c21 ca(5,5)=c11
      n0 = d0 * A.c11 + d1 * A.c21 + d2 * A.c31 + d3 * A.c41 + d4 * A.c51;
      n1 = d0 * A.c12 + d1 * A.c22 + d2 * A.c32 + d3 * A.c42 + d4 * A.c41;
      n2 = d0 * A.c13 + d1 * A.c23 + d2 * A.c33 + d3 * (-A.c32 / 2.0) + d4 * (-A.c31 / 2.0);
      n3 = d0 * A.c14 + d1 * A.c24 + d2 * (-2.0 * A.c23) + d3 * A.c22 + d4 * A.c21;
      n4 = d0 * A.c15 + d1 * A.c14 + d2 * (-2.0 * A.c13) + d3 * A.c12 + d4 * A.c11;
    }
    double lsc;
    const double sc = pow2_scale(max(max(max(RFS_ABSBITS(n0), RFS_ABSBITS(n1)),
                                         max(RFS_ABSBITS(n2), RFS_ABSBITS(n3))), RFS_ABSBITS(n4)), lsc);
    exsum = exsum + P.ex + S.ex + lsc;
    CDL(m * 6 + 0) = n0 * sc;
    CDL(m * 6 + 1) = n1 * sc;
    CDL(m * 6 + 2) = n2 * sc;
    CDL(m * 6 + 3) = n3 * sc;
    CDL(m * 6 + 4) = n4 * sc;
    CDL(m * 6 + 5) = exsum;
  }

  // ---------------- fused down-sweep / eigenfunctions / energy integrals
  const double f1213 = -CDL(1);
  const double if1213 = 1.0 / f1213;
  const double exe1 = CDL(5);
  Eig4 et;  // eigenfunction at the top of the current layer
  et.ur = CDL(2) / CDL(1);
  et.uz = 1.0;
  et.tz = 0.0;
  et.tr = 0.0;
  // leading water layers (svfunc :318-334): Ur = Tr = 0 at their tops
  bool lead_water = !(M.ld(F_B, 0, b) > 0.0);
  if (lead_water) {
    et.ur = 0.0;
    et.tr = 0.0;
  }
  double v0 = 1.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;  // vv(m,1:4)
  double exa_sum = 0.0;
  double sumi0 = 0.0, sumi1 = 0.0, sumi2 = 0.0, sumi3 = 0.0;
  const double cph = omega / wvno;
  double zr_prev = 0.0, xmu_prev = 0.0, xlam_prev = 0.0, ixmu_prev = 0.0, ixl2m_prev = 0.0,
         irom2_prev = 0.0;
  bool wat_prev = false;
  for (int m = 0; m < mmax; m++) {
    const bool half = (m == mmax - 1);
    const double za = M.ld(F_A, m, b), zb = M.ld(F_B, m, b), zr = M.ld(F_RHO, m, b),
                 zd = M.ld(F_D, m, b);
    const bool wat = !(zb > 0.0);
    const double xmu = zr * zb * zb;
    const double xlam = zr * za * za - 2 * xmu;
    const double xka = omega * M.ld(F_IA, m, b), xkb = wat ? 0.0 : omega * M.ld(F_IB, m, b);
    const double sa = __fma_rn(-xka, xka, wvno2), sb = __fma_rn(-xkb, xkb, wvno2);
    const double rom2 = zr * om2, irho = M.ld(F_IRHO, m, b), irom2 = irho * iom2;
    // 1/(rho b^2), 1/(rho a^2) from the reciprocals of the model block (no divisions)
    const double ibm = wat ? 0.0 : M.ld(F_IB, m, b), iam = M.ld(F_IA, m, b);
    const double ixmu = irho * ibm * ibm, ixl2m = irho * iam * iam;

    // ---- boundary term of dc/dh at the top of layer m (getdcdh :1436-1535)
    double gsum;
    {
      const double tuz = et.uz, ttz = et.tz, ttr = et.tr;
      const double tur = wat ? -wvno * ttz * irom2 : et.ur;
      const double xl2mp = xlam + xmu + xmu;
      const double duzdzp = (ttz + wvno * xlam * tur) * ixl2m;
      const double durdzp = (xmu == 0.0) ? wvno * tuz : (ttr * ixmu) - wvno * tuz;
      if (m == 0) {
        const double drho = zr, dmu = xmu, dl2mu = xlam + dmu + dmu;
        gsum = om2 * drho * tuz * tuz + om2 * (tur * tur * drho) - wvno2 * dmu * tuz * tuz -
               wvno2 * (tur * tur * dl2mu) + (xl2mp * duzdzp * duzdzp) + (xmu * durdzp * durdzp);
      } else {
        const double drho = zr - zr_prev, dmu = xmu - xmu_prev, dlm = xlam - xlam_prev;
        const double dl2mu = dlm + dmu + dmu;
        const double xl2mm = xlam_prev + xmu_prev + xmu_prev;
        const double durdzm = (xmu_prev == 0.0) ? wvno * tuz : (ttr * ixmu_prev) - wvno * tuz;
        double drur2, dlur2, duzdzm;
        if (wat_prev) {
          // Ur is discontinuous across a fluid boundary (:1497-1510)
          const double URB = -wvno * ttz * irom2_prev;
          drur2 = tur * tur * zr - URB * URB * zr_prev;
          dlur2 = tur * tur * xl2mp - URB * URB * xl2mm;
          duzdzm = (ttz + wvno * xlam_prev * URB) / (wat ? xl2mm : xlam_prev);
        } else {
          drur2 = tur * tur * drho;
          dlur2 = tur * tur * dl2mu;
          duzdzm = (ttz + wvno * xlam_prev * tur) * ixl2m_prev;
        }
        gsum = om2 * drho * tuz * tuz + om2 * drur2 - wvno2 * dmu * tuz * tuz - wvno2 * dlur2 +
               (xl2mp * duzdzp * duzdzp - xl2mm * duzdzm * duzdzm) +
               (xmu * durdzp * durdzp - xmu_prev * durdzm * durdzm);
      }
    }
    kern[(3LL * mmax + m) * ks] = gsum;  // scaled by `fac` in the epilogue

    // ---- nu_a, nu_b of this layer; they come from the up-sweep when available
    cd ra, rb;
    double ria, rib;  // 1/|nu_a|, 1/|nu_b| (0 for a vanishing wavenumber)
    VSV Pq, Sq;       // varsv terms of this layer, recomputed exactly as in the up-sweep
    Pq.c = Pq.rs = Pq.sr = Pq.ex = Pq.r = Pq.ri = Pq.e = 0.0;
    Pq.imag = false;
    Sq = Pq;
    if (!half) {
      Pq = varsv_half(sa, zd);
      if (wat) {
        Sq.c = 1.0;
        Sq.rs = 0.0;
        Sq.sr = 0.0;
        Sq.ex = 0.0;
        Sq.r = 0.0;
        Sq.ri = 0.0;
        Sq.e = 1.0;
        Sq.imag = false;
      } else {
        Sq = varsv_half(sb, zd);
      }
      const double sra = Pq.imag ? -Pq.r : Pq.r, srb = Sq.imag ? -Sq.r : Sq.r;
      ra = mk(sra < 0.0 || (sra == 0.0 && sa < 0.0), fabs(sra));
      rb = mk(srb < 0.0 || (srb == 0.0 && sb < 0.0), fabs(srb));
      ria = Pq.ri;
      rib = Sq.ri;
    } else {
      const double asa = fabs(sa), asb = fabs(sb);
      ria = (asa > 0.0) ? rsqrt_pos(asa) : 0.0;
      rib = (asb > 0.0) ? rsqrt_pos(asb) : 0.0;
      ra = mk(sa < 0.0, asa * ria);
      rb = mk(sb < 0.0, asb * rib);
    }
    // 1/nu for a real or purely imaginary nu: 1/r or -i/r
    const cd ira = (ra.y != 0.0) ? cd(0.0, -ria) : cd(ria, 0.0);

    Eig4 eb = et;  // eigenfunction at the bottom of the layer (top of m+1)
    VSV P, S;
    if (!half) {
      P = Pq;
      S = Sq;
      double w0, w1, w2, w3;
      if (wat) {
        // fluid Haskell step (hska :930-944)
        const double dfac = (P.ex > 35.0) ? 0.0 : exp_neg(P.ex);
        const double a23 = -P.rs * irom2, a32 = -rom2 * P.sr;
        w0 = dfac * v0;
        w1 = P.c * v1 + a23 * v2;
        w2 = a32 * v1 + P.c * v2;
        w3 = dfac * v3;
      } else {
        // ---- elastic Haskell step (hska :945-989, down :993-1063)
        const double dfac = ((P.ex - S.ex) > 70.0) ? 0.0 : exp_cb(S.ex - P.ex);
        const double cosp = P.c, rsinp = P.rs, sinpr = P.sr;
        const double cossv = dfac * S.c, rsinsv = dfac * S.rs, sinsvr = dfac * S.sr;
        const float bf = (float)zb;
        const double gmh = (double)__fmul_rn(__fmul_rn(2.0f, bf), bf) * wvno2 * iom2;
        const double gmh1 = gmh - 1.0;
        const double a11 = cossv + gmh * (cosp - cossv);
        const double a12 = -wvno * gmh1 * sinpr + gmh * rsinsv * iwv;
        const double a13 = -wvno * (cosp - cossv) * irom2;
        const double a14 = (wvno2 * sinpr - rsinsv) * irom2;
        const double a21 = gmh * rsinp * iwv - wvno * gmh1 * sinsvr;
        const double a22 = cosp - gmh * (cosp - cossv);
        const double a23 = (-rsinp + wvno2 * sinsvr) * irom2;
        const double a24 = -a13;
        const double a31 = rom2 * gmh * gmh1 * (cosp - cossv) * iwv;
        const double a32 = rom2 * (-gmh1 * gmh1 * sinpr + gmh * gmh * rsinsv * iwv2);
        const double a33 = a22, a34 = -a12;
        const double a41 = rom2 * (gmh * gmh * rsinp * iwv2 - gmh1 * gmh1 * sinsvr);
        const double a42 = -a31, a43 = -a21, a44 = a11;
        w0 = a11 * v0 + a12 * v1 + a13 * v2 + a14 * v3;
        w1 = a21 * v0 + a22 * v1 + a23 * v2 + a24 * v3;
        w2 = a31 * v0 + a32 * v1 + a33 * v2 + a34 * v3;
        w3 = a41 * v0 + a42 * v1 + a43 * v2 + a44 * v3;
      }
      double lsc;
      const double sc = pow2_scale(max(max(RFS_ABSBITS(w0), RFS_ABSBITS(w1)),
                                       max(RFS_ABSBITS(w2), RFS_ABSBITS(w3))), lsc);
      v0 = w0 * sc;
      v1 = w1 * sc;
      v2 = w2 * sc;
      v3 = w3 * sc;
      exa_sum = exa_sum + P.ex + lsc;
      // ---- eigenfunction at the top of layer m+1 (svfunc :268-315)
      const int i = m + 1;
      const double cd1 = CDL(i * 6 + 0), cd2 = CDL(i * 6 + 1), cd3 = CDL(i * 6 + 2), cd4 = -cd3,
                   cd5 = CDL(i * 6 + 3), cd6 = CDL(i * 6 + 4);
      const double tz1 = -v3, tz2 = -v2, tz3 = v1, tz4 = v0;
      const double uu1 = tz2 * cd6 - tz3 * cd5 + tz4 * cd4;
      const double uu2 = -tz1 * cd6 + tz3 * cd3 - tz4 * cd2;
      const double uu3 = tz1 * cd5 - tz2 * cd3 + tz4 * cd1;
      const double uu4 = -tz1 * cd4 + tz2 * cd2 - tz3 * cd1;
      const double ext = exa_sum + CDL(i * 6 + 5) - exe1;
      if (ext > -80.0 && ext < 80.0) {
        const double fact = exp_cb(ext) * if1213;
        eb.ur = uu1 * fact;
        eb.uz = uu2 * fact;
        eb.tz = uu3 * fact;
        eb.tr = uu4 * fact;
      } else {
        eb.ur = eb.uz = eb.tz = eb.tr = 0.0;
      }
      lead_water = lead_water && wat && !(M.ld(F_B, i, b) > 0.0);
      if (lead_water) {
        eb.ur = 0.0;
        eb.tr = 0.0;
      }
    }

    if (wat) {
      // ---- fluid layer: 2x2 potentials (intijr :1246-1266) and energy (energy :1123-1143)
      const cd kmpu = (0.5 * ira) * eb.uz - (0.5 * irom2) * eb.tz;
      const cd km1pd = -1.0 * ((0.5 * ira) * et.uz) - (0.5 * irom2) * et.tz;
      const cd FA = ffunc_d(ra, zd), GA = gfunc_d(ra, zd);
      const cd pp = kmpu * kmpu * FA + km1pd * km1pd * FA, pm = 2.0 * (kmpu * km1pd * GA);
      const double I11 = ((ra * ra) * (pp - pm)).x;
      const double I22 = (rom2 * rom2) * (pp + pm).x;
      const double a12f = -(wvno2 - om2 / (za * za)) / rom2;
      const double wq = wvno / rom2;
      const double URUR = I22 * wq * wq, UZUZ = I11, URDUZ = -wq * a12f * I22,
                   DUZDUZ = a12f * a12f * I22;
      const double TA = zr * za * za;
      sumi0 += zr * (URUR + UZUZ);
      sumi1 += TA * URUR;
      sumi2 -= TA * URDUZ;
      sumi3 += TA * DUZDUZ;
      const double facah = zr * za * (URUR - 2. * URDUZ / wvno);
      const double facav = zr * za * DUZDUZ / wvno2;
      const double facr = -0.5 * cph * cph * (URUR + UZUZ);
      kern[(0LL * mmax + m) * ks] = facah + facav;
      kern[(1LL * mmax + m) * ks] = 0.0;  // reference leaves dcdb of a fluid layer unset
      kern[(2LL * mmax + m) * ks] = 0.5 * (za * facav + za * facah) / zr + facr;
    } else {
      // ---- E, E^-1 of this layer (evalg :736-768)
      double gam = zb * slow;
      gam = 2.0 * (gam * gam);
      const double gamm1 = gam - 1.0;
      const cd irb = (rb.y != 0.0) ? cd(0.0, -rib) : cd(rib, 0.0);
      // EINV rows (acting on [ur, uz, tz, tr])
      const double ei11 = 0.5 * gam * iwv, ei13 = -0.5 * irom2;
      const cd ei12 = (-0.5 * gamm1) * ira, ei14 = (0.5 * wvno * irom2) * ira;
      const cd ei21 = (-0.5 * gamm1) * irb, ei23 = (0.5 * wvno * irom2) * irb;
      // rows 3,4: EINV(3,:) = [ei11, -ei12, ei13, -ei14]; EINV(4,:) = [-ei21, ei11, -ei23, ei13]
      // ---- potentials (intijr :1203-1323)
      // downward coefficients at the top of the layer (rows 3,4 of E^-1), upward at the bottom
      const cd km1pd = ei11 * et.ur - ei12 * et.uz + ei13 * et.tz - ei14 * et.tr;
      const cd km1sd = -1.0 * (ei21 * et.ur) + ei11 * et.uz - ei23 * et.tz + ei13 * et.tr;
      // E columns: E(:,1)=[k, ra, r g1, r g ra/k]  E(:,2)=[rb, k, r g rb/k, r g1]
      //            E(:,3)=[k,-ra, r g1,-r g ra/k]  E(:,4)=[-rb, k,-r g rb/k, r g1]     (r = rho om^2)
      const double rg1 = rom2 * gamm1;
      const cd e41 = (rom2 * gam * iwv) * ra, e32 = (rom2 * gam * iwv) * rb;
      cd a3[4], a4[4];  // a_i3 = E(i,3) km1pd, a_i4 = E(i,4) km1sd
      a3[0] = wvno * km1pd;
      a3[1] = -1.0 * (ra * km1pd);
      a3[2] = rg1 * km1pd;
      a3[3] = -1.0 * (e41 * km1pd);
      a4[0] = -1.0 * (rb * km1sd);
      a4[1] = wvno * km1sd;
      a4[2] = -1.0 * (e32 * km1sd);
      a4[3] = rg1 * km1sd;
      double I11, I13, I22, I24, I33, I44;
      if (half) {
        const cd qa = 0.5 * ira, qb = 0.5 * irb, qab = cinv(ra + rb);
#define RFS_HS(i, j) \
  ((a3[i] * a3[j]) * qa + (a3[i] * a4[j] + a4[i] * a3[j]) * qab + (a4[i] * a4[j]) * qb).x
        I11 = RFS_HS(0, 0);
        I13 = RFS_HS(0, 2);
        I22 = RFS_HS(1, 1);
        I24 = RFS_HS(1, 3);
        I33 = RFS_HS(2, 2);
        I44 = RFS_HS(3, 3);
#undef RFS_HS
      } else {
        const cd kmpu = ei11 * eb.ur + ei12 * eb.uz + ei13 * eb.tz + ei14 * eb.tr;
        const cd kmsu = ei21 * eb.ur + ei11 * eb.uz + ei23 * eb.tz + ei13 * eb.tr;
        cd a1[4], a2[4];
        a1[0] = wvno * kmpu;
        a1[1] = ra * kmpu;
        a1[2] = rg1 * kmpu;
        a1[3] = e41 * kmpu;
        a2[0] = rb * kmsu;
        a2[1] = wvno * kmsu;
        a2[2] = e32 * kmsu;
        a2[3] = rg1 * kmsu;
        // ffunc/gfunc/h1func/h2func (:1325-1403) from TWO complex exponentials: with
        // ea = exp(-nu_a d), eb = exp(-nu_b d): exp(-2 nu d) = e^2, exp(-(nu_a+nu_b) d) = ea eb
        // (the reference evaluates seven); cut-offs as in the reference
        const cd arga = ra * zd, argb = rb * zd;
        const cd ea = (arga.x < 75.0) ? cexp_b(-1.0 * arga) : cd(0.0);
        const cd ebx = (argb.x < 75.0) ? cexp_b(-1.0 * argb) : cd(0.0);
        const cd sab = ra + rb, dab = ra - rb;
        const cd FA = (norm2(ra) < 1.0e-16) ? cd(zd)
                                            : (1.0 - ((arga.x < 40.0) ? ea * ea : cd(0.0))) * (0.5 * ira);
        const cd FB = (norm2(rb) < 1.0e-16) ? cd(zd)
                                            : (1.0 - ((argb.x < 40.0) ? ebx * ebx : cd(0.0))) * (0.5 * irb);
        const cd GA = ea * zd, GB = ebx * zd;
        const cd H1 = (norm2(sab) < 1.0e-16)
                          ? cd(zd)
                          : (1.0 - ((arga.x + argb.x < 40.0) ? ea * ebx : cd(0.0))) * cinv(sab);
        const cd H2 = (norm2(dab) < 1.0e-16)
                          ? cd(zd)
                          : (((argb.x < 40.0) ? ebx : cd(0.0)) - ((arga.x < 40.0) ? ea : cd(0.0))) * cinv(dab);
        // INT_ij = a_i^T W a_j,  W = [[P,Q],[Q,P]], P = [[FA,H1],[H1,FB]], Q = [[GA,H2],[H2,GB]].
        // With s = up + down, t = up - down:  a_i^T W a_j = (s_i^T (P+Q) s_j + t_i^T (P-Q) t_j)/2,
        // two 2x2 forms instead of one 4x4 (56 instead of 88 complex products per layer).
        const cd pa = FA + GA, ph = H1 + H2, pb = FB + GB, ma = FA - GA, mh = H1 - H2, mb = FB - GB;
        cd s1[4], s2[4], t1[4], t2[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          s1[q] = a1[q] + a3[q];
          s2[q] = a2[q] + a4[q];
          t1[q] = a1[q] - a3[q];
          t2[q] = a2[q] - a4[q];
        }
#define RFS_WJ(j)                                   \
  const cd bs1 = pa * s1[j] + ph * s2[j];           \
  const cd bs2 = ph * s1[j] + pb * s2[j];           \
  const cd bt1 = ma * t1[j] + mh * t2[j];           \
  const cd bt2 = mh * t1[j] + mb * t2[j];
#define RFS_RE(u_, v_) ((u_).x * (v_).x - (u_).y * (v_).y)
#define RFS_DOT(i) \
  (0.5 * (RFS_RE(s1[i], bs1) + RFS_RE(s2[i], bs2) + RFS_RE(t1[i], bt1) + RFS_RE(t2[i], bt2)))
        {
          RFS_WJ(0) I11 = RFS_DOT(0);
        }
        {
          RFS_WJ(1) I22 = RFS_DOT(1);
        }
        {
          RFS_WJ(2) I13 = RFS_DOT(0);
          I33 = RFS_DOT(2);
        }
        {
          RFS_WJ(3) I24 = RFS_DOT(1);
          I44 = RFS_DOT(3);
        }
#undef RFS_WJ
#undef RFS_RE
#undef RFS_DOT
      }
      // ---- energy integrals and un-normalised partials (energy :1144-1183, getmat :1537-1589)
      const double TL = zr * zb * zb, TC = zr * za * za, TA = TC, TF = TA - 2. * TL;
      const double a12 = -wvno, a14 = ixmu, a23 = ixl2m, a21 = wvno * TF * a23;
      const double URUR = I11, UZUZ = I22;
      const double DURDUR = a12 * a12 * I22 + 2. * a12 * a14 * I24 + a14 * a14 * I44;
      const double DUZDUZ = a21 * a21 * I11 + 2. * a21 * a23 * I13 + a23 * a23 * I33;
      const double URDUZ = a21 * I11 + a23 * I13;
      const double UZDUR = a12 * I22 + a14 * I24;
      sumi0 += zr * (URUR + UZUZ);
      sumi1 += TL * UZUZ + TA * URUR;
      sumi2 += TL * UZDUR - TF * URDUZ;
      sumi3 += TL * DURDUR + TC * DUZDUZ;
      const double facah = zr * za * (URUR - 2. * URDUZ * iwv);
      const double facav = zr * za * DUZDUZ * iwv2;
      const double facbv = zr * zb * (UZUZ + 2. * UZDUR * iwv + DURDUR * iwv2 + 4. * URDUZ * iwv);
      const double facr = -0.5 * cph * cph * (URUR + UZUZ);
      kern[(0LL * mmax + m) * ks] = facah + facav;
      kern[(1LL * mmax + m) * ks] = facbv;
      kern[(2LL * mmax + m) * ks] = 0.5 * (za * facav + za * facah + zb * facbv) * irho + facr;
    }
    et = eb;
    zr_prev = zr;
    xmu_prev = xmu;
    xlam_prev = xlam;
    ixmu_prev = ixmu;
    ixl2m_prev = ixl2m;
    irom2_prev = irom2;
    wat_prev = wat;
  }
  (void)sumi3;
  // ---------------- epilogue: U, normalisation, boundary -> thickness suffix sums
  const double ugr = (wvno * sumi1 + sumi2) / (omega * sumi0);
  const double are = wvno / (2.0 * omega * ugr * sumi0);
  const double fac = are * cph / wvno2;
  const double inrm = 1.0 / (ugr * sumi0);
  double suffix = 0.0;
  for (int m = mmax - 1; m >= 0; m--) {
    kern[(0LL * mmax + m) * ks] = kern[(0LL * mmax + m) * ks] * inrm;
    kern[(1LL * mmax + m) * ks] = kern[(1LL * mmax + m) * ks] * inrm;
    kern[(2LL * mmax + m) * ks] = kern[(2LL * mmax + m) * ks] * inrm;
    double dfac = fac * kern[(3LL * mmax + m) * ks];
    if (fabs(dfac) < 1.0e-38) dfac = 0.0;
    kern[(3LL * mmax + m) * ks] = suffix;  // dcdh(i) = sum_{j>i} raw(j) dtp(j); dcdh(mmax) = 0
    suffix += dfac * M.ld(F_DTP, m, b);     // dtp = 1 on a flat earth (sprayl :1619-1626 otherwise)
  }
  *ugr_out = ugr;
#undef CDL
}

}  // namespace rfs
