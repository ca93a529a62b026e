// ORACLE — TEST INFRASTRUCTURE ONLY.
// Fortran entry points the reference's C++ expects (/root/reference/src/SWD/surfdisp.hpp:17-95,
// /root/reference/src/RF/rf_cal.hpp:10-52), forwarded to the C++ restatements of those Fortran
// routines in this directory.  With these shims the reference's OWN C++ -- src/SWD/main.cpp,
// src/SWD/surfdisp.cpp, src/RF/main.cpp (pybind11 boundary, float32 casts, retry loop,
// _RayleighGroup/_LoveGroup, _SurfKernel, _flat2sphere) -- compiles and runs here unmodified
// (`make -C oracle ref` -> oracle/_ref/libsurf.so, librf.so).  gfortran and FFTW3 are absent, so
// the Fortran underneath remains the restatement: _ref pins the restated C++ layers
// (swd_driver.cpp, oracle_capi.cpp), not the numerics.
#include "oracle.hpp"

extern "C" {

void surfdisp96_(float *thkm, float *vpm, float *vsm, float *rhom, int nlayer, int iflsph, int iwave,
                 int mode, int igr, int kmax, double *t, double *cg, int *ierr) {
  oracle::surfdisp96(thkm, vpm, vsm, rhom, nlayer, iflsph, iwave, mode, igr, kmax, t, cg, ierr);
}

void sregn96_(float *thk, float *vp, float *vs, float *rhom, int nlayer, double *t, double *cp,
              double *cg, double *dispu, double *dispw, double *stressu, double *stressw,
              double *dc2da, double *dc2db, double *dc2dh, double *dc2dr, int iflsph) {
  oracle::sregn96(thk, vp, vs, rhom, nlayer, t, cp, cg, dispu, dispw, stressu, stressw, dc2da, dc2db,
                  dc2dh, dc2dr, iflsph);
}

void slegn96_(float *thk, float *vs, float *rhom, int nlayer, double *t, double *cp, double *cg,
              double *disp, double *stress, double *dc2db, double *dc2dh, double *dc2dr, int iflsph) {
  oracle::slegn96(thk, vs, rhom, nlayer, t, cp, cg, disp, stress, dc2db, dc2dh, dc2dr, iflsph);
}

void slegnpu_(float *thk, float *vs, float *rhom, int nlayer, double *t, double *cp, double *cg,
              double *disp, double *stress, double *t1, double *cp1, double *t2, double *cp2,
              double *dc2db, double *dc2dh, double *dc2dr, double *du2db, double *du2dh,
              double *du2dr, int iflsph) {
  oracle::slegnpu(thk, vs, rhom, nlayer, t, cp, cg, disp, stress, t1, cp1, t2, cp2, dc2db, dc2dh, dc2dr,
                  du2db, du2dh, du2dr, iflsph, true);
}

void sregnpu_(float *thk, float *vp, float *vs, float *rhom, int nlayer, double *t, double *cp,
              double *cg, double *dispu, double *dispw, double *stressu, double *stressw, double *t1,
              double *cp1, double *t2, double *cp2, double *dc2da, double *dc2db, double *dc2dh,
              double *dc2dr, double *du2da, double *du2db, double *du2dh, double *du2dr, int iflsph) {
  oracle::sregnpu(thk, vp, vs, rhom, nlayer, t, cp, cg, dispu, dispw, stressu, stressw, t1, cp1, t2, cp2,
                  dc2da, dc2db, dc2dh, dc2dr, du2da, du2db, du2dh, du2dr, iflsph, true);
}

void cal_rf_time_(const double *thk, const double *vp, const double *vs, const double *rho,
                  const double *qa, const double *qb, int nlayer, int nt, double dt, double ray_p,
                  double gauss, double time_shift, int rf_type, double *rcv_fun) {
  oracle::cal_rf_time(thk, vp, vs, rho, qa, qb, nlayer, nt, dt, ray_p, gauss, time_shift, rf_type, rcv_fun);
}

void cal_rf_freq_(const double *thk, const double *vp, const double *vs, const double *rho,
                  const double *qa, const double *qb, int nlayer, int nt, double dt, double ray_p,
                  double gauss, double time_shift, double water, int rf_type, double *rcv_fun) {
  oracle::cal_rf_freq(thk, vp, vs, rho, qa, qb, nlayer, nt, dt, ray_p, gauss, time_shift, water, rf_type,
                      rcv_fun);
}

void cal_rf_par_freq_(const double *thk, const double *vp, const double *vs, const double *rho,
                      const double *qa, const double *qb, int nlayer, int nt, double dt, double ray_p,
                      double gauss, double time_shift, double water, int rf_type, int par_type,
                      double *rcv_fun, double *rcv_fun_p) {
  oracle::cal_rf_par_freq(thk, vp, vs, rho, qa, qb, nlayer, nt, dt, ray_p, gauss, time_shift, water,
                          rf_type, par_type, rcv_fun, rcv_fun_p);
}

void cal_rf_par_freq_all_(const double *thk, const double *vp, const double *vs, const double *rho,
                          const double *qa, const double *qb, int nlayer, int nt, double dt,
                          double ray_p, double gauss, double time_shift, double water, int rf_type,
                          double *rcv_fun, double *rcv_fun_p) {
  oracle::cal_rf_par_freq_all(thk, vp, vs, rho, qa, qb, nlayer, nt, dt, ray_p, gauss, time_shift, water,
                              rf_type, rcv_fun, rcv_fun_p);
}

void cal_rf_par_time_(const double *thk, const double *vp, const double *vs, const double *rho,
                      const double *qa, const double *qb, int nlayer, int nt, double dt, double ray_p,
                      double gauss, double time_shift, int rf_type, int par_type, double *rcv_fun,
                      double *rcv_fun_p) {
  oracle::cal_rf_par_time(thk, vp, vs, rho, qa, qb, nlayer, nt, dt, ray_p, gauss, time_shift, rf_type,
                          par_type, rcv_fun, rcv_fun_p);
}

void cal_rf_par_time_all_(const double *thk, const double *vp, const double *vs, const double *rho,
                          const double *qa, const double *qb, int nlayer, int nt, double dt,
                          double ray_p, double gauss, double time_shift, int rf_type, double *rcv_fun,
                          double *rcv_fun_p) {
  oracle::cal_rf_par_time_all(thk, vp, vs, rho, qa, qb, nlayer, nt, dt, ray_p, gauss, time_shift, rf_type,
                              rcv_fun, rcv_fun_p);
}

}  // extern "C"
