// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/README.md).
//
// Restates the C++ orchestration of /root/reference/src/SWD/surfdisp.cpp:
//   _flat2sphere :16-49   _surfdisp :62-109   _LoveGroup :119-141   _RayleighGroup :151-173
//   _SurfKernel :190-297
// Deviations (documented in DESIGN.md §oracle):
//   * invalid wave type -> returns -2 instead of exit(0);
//   * _LoveGroup's out-of-bounds `vp[kmax]` scratch (surfdisp.cpp:127,132) is replaced by
//     vp[i]=1.732*vs[i] over the n layers (Love ignores vp except in the start value).
//   * dcda for Love is defined as 0 (reference leaves it uninitialised, main.cpp:68).
#include "oracle.hpp"
#include <cmath>
#include <vector>

namespace oracle {

double flat2sphere(double t, double c, const std::string &wavetp) {
  double ar = 6371.0;
  double omega = 2.0 * M_PI / t;
  double tm;
  if (wavetp[0] == 'L')
    tm = 1. + std::pow(1.5 * c / (ar * omega), 2);
  else
    tm = 1. + std::pow(0.5 * c / (ar * omega), 2);
  tm = std::sqrt(tm);
  if (wavetp[1] == 'c') return c / tm;
  return c * tm;
}

int surfdisp(const float *thk, const float *vp, const float *vs, const float *rho, int nlayer,
             const double *t, double *cg, int kmax, const std::string &wavetype, int mode,
             bool sphere, bool keep_flat) {
  int iwave, igr;
  if (wavetype == "Rc") {
    iwave = 2;
    igr = 0;
  } else if (wavetype == "Rg") {
    iwave = 2;
    igr = 1;
  } else if (wavetype == "Lc") {
    iwave = 1;
    igr = 0;
  } else if (wavetype == "Lg") {
    iwave = 1;
    igr = 1;
  } else {
    return -2;
  }
  if (igr != 0) return -2;  // never reached from the wrappers (dead code)
  int ifsph = sphere ? 1 : 0;
  int ierr;
  surfdisp96(thk, vp, vs, rho, nlayer, ifsph, iwave, mode + 1, igr, kmax, t, cg, &ierr);
  if (ierr != 0) {
    for (int i = 0; i < kmax; i++) {
      if (cg[i] == 0.0 || std::isnan(cg[i])) {
        surfdisp96(thk, vp, vs, rho, nlayer, ifsph, iwave, mode + 1, igr, 1, &t[i], &cg[i], &ierr);
        if (ierr != 0) return ierr;
      }
    }
  }
  if (sphere && !keep_flat)
    for (int i = 0; i < kmax; i++) cg[i] = flat2sphere(t[i], cg[i], wavetype);
  return ierr;
}

int love_group(const float *thk, const float *vs, const float *rho, int nlayer, const double *t,
               double *cg, int kmax, int mode, bool sphere) {
  int iflsph = sphere ? 1 : 0;
  std::vector<float> vp(nlayer);
  std::vector<double> cp(kmax), uu(nlayer), tt(nlayer), dcdh(nlayer), dcdr(nlayer), dcdb(nlayer);
  for (int i = 0; i < nlayer; i++) vp[i] = (float)(1.732 * (double)vs[i]);  // double literal, float store
  int ierr = surfdisp(thk, vp.data(), vs, rho, nlayer, t, cp.data(), kmax, "Lc", mode, sphere, true);
  if (ierr == 1) return ierr;
  for (int i = 0; i < kmax; i++) {
    double ti = t[i];
    slegn96(thk, vs, rho, nlayer, &ti, &cp[i], &cg[i], uu.data(), tt.data(), dcdb.data(),
            dcdh.data(), dcdr.data(), iflsph);
  }
  return ierr;
}

int rayleigh_group(const float *thk, const float *vp, const float *vs, const float *rho,
                   int nlayer, const double *t, double *cg, int kmax, int mode, bool sphere) {
  int iflsph = sphere ? 1 : 0;
  std::vector<double> cp(kmax), ur(nlayer), uz(nlayer), tr(nlayer), tz(nlayer), dcdh(nlayer),
      dcda(nlayer), dcdr(nlayer), dcdb(nlayer);
  int ierr = surfdisp(thk, vp, vs, rho, nlayer, t, cp.data(), kmax, "Rc", mode, sphere, true);
  if (ierr == 1) return ierr;
  for (int i = 0; i < kmax; i++) {
    double ti = t[i];
    sregn96(thk, vp, vs, rho, nlayer, &ti, &cp[i], &cg[i], ur.data(), uz.data(), tr.data(),
            tz.data(), dcda.data(), dcdb.data(), dcdh.data(), dcdr.data(), iflsph);
  }
  return ierr;
}

int surf_kernel(const float *thk, const float *vp, const float *vs, const float *rho, int nlayer,
                const double *t, double *c, int nt, double *dcda, double *dcdb, double *dcdr,
                double *dcdh, const std::string &wavetp, int mode, bool sphere,
                bool stale_first_term) {
  bool ok = wavetp == "Rc" || wavetp == "Rg" || wavetp == "Lc" || wavetp == "Lg";
  if (!ok) return -2;
  const bool keep_flat = true;
  int iflsph = sphere ? 1 : 0;
  int ierr;
  const int n = nlayer;
  if (wavetp == "Rc") {
    ierr = surfdisp(thk, vp, vs, rho, nlayer, t, c, nt, wavetp, mode, sphere, keep_flat);
    if (ierr == 1) return ierr;
    double cg;
    std::vector<double> ur(n), uz(n), tr(n), tz(n);
    for (int i = 0; i < nt; i++) {
      int k = i * n;
      double ti = t[i];
      sregn96(thk, vp, vs, rho, n, &ti, c + i, &cg, ur.data(), uz.data(), tr.data(), tz.data(),
              dcda + k, dcdb + k, dcdh + k, dcdr + k, iflsph);
    }
  } else if (wavetp == "Rg") {
    std::vector<double> cp(nt), cp1(nt), cp2(nt), t1(nt), t2(nt);
    for (int i = 0; i < nt; i++) {
      t1[i] = t[i] * (1.0 + 0.05);
      t2[i] = t[i] * (1.0 - 0.05);
    }
    ierr = surfdisp(thk, vp, vs, rho, n, t, cp.data(), nt, "Rc", mode, sphere, keep_flat);
    int ierr1 = surfdisp(thk, vp, vs, rho, n, t1.data(), cp1.data(), nt, "Rc", mode, sphere, keep_flat);
    int ierr2 = surfdisp(thk, vp, vs, rho, n, t2.data(), cp2.data(), nt, "Rc", mode, sphere, keep_flat);
    ierr = (ierr + ierr1 + ierr2) > 0;
    if (ierr == 1) return ierr;
    std::vector<double> ur(n), uz(n), tr(n), tz(n), a1(n), b1(n), r1(n), h1(n);
    for (int i = 0; i < nt; i++) {
      int k = i * n;
      double ti = t[i];
      sregnpu(thk, vp, vs, rho, n, &ti, &cp[i], c + i, ur.data(), uz.data(), tr.data(), tz.data(),
              &t1[i], &cp1[i], &t2[i], &cp2[i], a1.data(), b1.data(), h1.data(), r1.data(),
              dcda + k, dcdb + k, dcdh + k, dcdr + k, iflsph, stale_first_term);
    }
  } else if (wavetp == "Lc") {
    ierr = surfdisp(thk, vp, vs, rho, n, t, c, nt, wavetp, mode, sphere, keep_flat);
    if (ierr == 1) return ierr;
    double cg;
    std::vector<double> uu(n), tt(n);
    for (int i = 0; i < nt; i++) {
      int k = i * n;
      double ti = t[i];
      slegn96(thk, vs, rho, n, &ti, c + i, &cg, uu.data(), tt.data(), dcdb + k, dcdh + k,
              dcdr + k, iflsph);
      for (int j = 0; j < n; j++) dcda[k + j] = 0.0;
    }
  } else {
    std::vector<double> cp(nt), cp1(nt), cp2(nt), t1(nt), t2(nt);
    for (int i = 0; i < nt; i++) {
      t1[i] = t[i] * (1.0 + 0.05);
      t2[i] = t[i] * (1.0 - 0.05);
    }
    ierr = surfdisp(thk, vp, vs, rho, n, t, cp.data(), nt, "Lc", mode, sphere, keep_flat);
    int ierr1 = surfdisp(thk, vp, vs, rho, n, t1.data(), cp1.data(), nt, "Lc", mode, sphere, keep_flat);
    int ierr2 = surfdisp(thk, vp, vs, rho, n, t2.data(), cp2.data(), nt, "Lc", mode, sphere, keep_flat);
    ierr = ierr || ierr1 || ierr2;
    if (ierr == 1) return ierr;
    std::vector<double> uu(n), tt(n), b1(n), r1(n), h1(n);
    for (int i = 0; i < nt; i++) {
      int k = i * n;
      double ti = t[i];
      slegnpu(thk, vs, rho, n, &ti, &cp[i], c + i, uu.data(), tt.data(), &t1[i], &cp1[i], &t2[i],
              &cp2[i], b1.data(), h1.data(), r1.data(), dcdb + k, dcdh + k, dcdr + k, iflsph,
              stale_first_term);
      for (int j = 0; j < n; j++) dcda[k + j] = 0.0;
    }
  }
  return ierr;
}

}  // namespace oracle
