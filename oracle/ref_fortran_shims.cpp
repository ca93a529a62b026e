// ORACLE — TEST INFRASTRUCTURE ONLY.
// The reference's C++ (src/SWD/main.cpp, src/SWD/surfdisp.cpp, src/RF/main.cpp) calls eleven Fortran
// routines through the C ABI (declared in src/SWD/surfdisp.hpp:17-95 and src/RF/rf_cal.hpp:10-52).
// gfortran is absent here, so `make -C oracle ref` links that C++ -- compiled unmodified from
// /root/reference -- against the forwarding functions below, which hand every call to the C++
// restatement of the corresponding Fortran routine in this directory.  The result
// (oracle/_ref/libsurf.so, librf.so) runs the reference's real pybind11 boundary, float32 casts,
// retry loop, group-velocity / kernel drivers and flat->sphere conversion; it pins the restated C++
// layers (swd_driver.cpp, oracle_capi.cpp), not the numerics underneath.
#include "oracle.hpp"

namespace orc = oracle;
using F = float *;
using D = double *;
using CD = const double *;

extern "C" {

// ---- src/SWD/surfdisp96.f, sregn96.f90, slegn96.f90
void surfdisp96_(F h, F a, F b, F r, int n, int sph, int wave, int mode, int igr, int nper, D per,
                 D vel, int *err) {
  orc::surfdisp96(h, a, b, r, n, sph, wave, mode, igr, nper, per, vel, err);
}
void sregn96_(F h, F a, F b, F r, int n, D per, D c, D u, D ur, D uz, D tr, D tz, D ka, D kb, D kh,
              D kr, int sph) {
  orc::sregn96(h, a, b, r, n, per, c, u, ur, uz, tr, tz, ka, kb, kh, kr, sph);
}
void slegn96_(F h, F b, F r, int n, D per, D c, D u, D ut, D tt, D kb, D kh, D kr, int sph) {
  orc::slegn96(h, b, r, n, per, c, u, ut, tt, kb, kh, kr, sph);
}
void sregnpu_(F h, F a, F b, F r, int n, D per, D c, D u, D ur, D uz, D tr, D tz, D p1, D c1, D p2,
              D c2, D ka, D kb, D kh, D kr, D ga, D gb, D gh, D gr, int sph) {
  orc::sregnpu(h, a, b, r, n, per, c, u, ur, uz, tr, tz, p1, c1, p2, c2, ka, kb, kh, kr, ga, gb, gh, gr,
               sph, /*stale_first_term=*/true);
}
void slegnpu_(F h, F b, F r, int n, D per, D c, D u, D ut, D tt, D p1, D c1, D p2, D c2, D kb, D kh,
              D kr, D gb, D gh, D gr, int sph) {
  orc::slegnpu(h, b, r, n, per, c, u, ut, tt, p1, c1, p2, c2, kb, kh, kr, gb, gh, gr, sph,
               /*stale_first_term=*/true);
}

// ---- src/RF/RFModule.f90 (argument order of the Fortran bind(C) interfaces: thk, vp, vs, rho)
void cal_rf_time_(CD h, CD a, CD b, CD r, CD qa, CD qb, int n, int nt, double dt, double p, double g,
                  double t0, int type, D rf) {
  orc::cal_rf_time(h, a, b, r, qa, qb, n, nt, dt, p, g, t0, type, rf);
}
void cal_rf_freq_(CD h, CD a, CD b, CD r, CD qa, CD qb, int n, int nt, double dt, double p, double g,
                  double t0, double wl, int type, D rf) {
  orc::cal_rf_freq(h, a, b, r, qa, qb, n, nt, dt, p, g, t0, wl, type, rf);
}
void cal_rf_par_freq_(CD h, CD a, CD b, CD r, CD qa, CD qb, int n, int nt, double dt, double p,
                      double g, double t0, double wl, int type, int par, D rf, D drf) {
  orc::cal_rf_par_freq(h, a, b, r, qa, qb, n, nt, dt, p, g, t0, wl, type, par, rf, drf);
}
void cal_rf_par_freq_all_(CD h, CD a, CD b, CD r, CD qa, CD qb, int n, int nt, double dt, double p,
                          double g, double t0, double wl, int type, D rf, D drf) {
  orc::cal_rf_par_freq_all(h, a, b, r, qa, qb, n, nt, dt, p, g, t0, wl, type, rf, drf);
}
void cal_rf_par_time_(CD h, CD a, CD b, CD r, CD qa, CD qb, int n, int nt, double dt, double p,
                      double g, double t0, int type, int par, D rf, D drf) {
  orc::cal_rf_par_time(h, a, b, r, qa, qb, n, nt, dt, p, g, t0, type, par, rf, drf);
}
void cal_rf_par_time_all_(CD h, CD a, CD b, CD r, CD qa, CD qb, int n, int nt, double dt, double p,
                          double g, double t0, int type, D rf, D drf) {
  orc::cal_rf_par_time_all(h, a, b, r, qa, qb, n, nt, dt, p, g, t0, type, rf, drf);
}

}  // extern "C"
