// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/README.md).
// Internal C++ interface of the CPU restatement of the reference's native path.
// Each function cites the reference file:line it follows in its .cpp.
#pragma once
#include <complex>
#include <string>

namespace oracle {

typedef std::complex<double> cplx;

// float32 value of pi used by sregn96/sregnpu/slegn96/slegnpu and most RF routines
// (`atan(1.0)*4.0` in default REAL; sregn96.f90:1654, slegn96.f90:689, RFModule.f90:28,161,211,277,364)
static const double PI32 = (double)3.14159274101257324f;

// ---- src/SWD/surfdisp96.f
void surfdisp96(const float *thkm, const float *vpm, const float *vsm, const float *rhom,
                int nlayer, int iflsph, int iwave, int mode, int igr, int kmax, const double *t,
                double *cg, int *ierr, long *n_evals = nullptr);

// ---- src/SWD/sregn96.f90
// stale_first_term = true reproduces the reference's use of the T2 solve's kernels in the first
// term of dU/dm (sregn96.f90:1841-1844); false uses the T solve's kernels (the "fixed" form).
void sregn96(const float *thk, const float *vp, const float *vs, const float *rhom, int nlayer,
             double *t, double *cp, double *cg, double *dispu, double *dispw, double *stressu,
             double *stressw, double *dc2da, double *dc2db, double *dc2dh, double *dc2dr,
             int iflsph);
void sregnpu(const float *thk, const float *vp, const float *vs, const float *rhom, int nlayer,
             double *t, double *cp, double *cg, double *dispu, double *dispw, double *stressu,
             double *stressw, double *t1, double *cp1, double *t2, double *cp2, double *dc2da,
             double *dc2db, double *dc2dh, double *dc2dr, double *du2da, double *du2db,
             double *du2dh, double *du2dr, int iflsph, bool stale_first_term = true);

// ---- src/SWD/slegn96.f90
void slegn96(const float *thk, const float *vs, const float *rhom, int nlayer, double *t,
             double *cp, double *cg, double *disp, double *stress, double *dc2db, double *dc2dh,
             double *dc2dr, int iflsph);
void slegnpu(const float *thk, const float *vs, const float *rhom, int nlayer, double *t,
             double *cp, double *cg, double *disp, double *stress, double *t1, double *cp1,
             double *t2, double *cp2, double *dc2db, double *dc2dh, double *dc2dr, double *du2db,
             double *du2dh, double *du2dr, int iflsph, bool stale_first_term = true);

// ---- src/SWD/surfdisp.cpp
double flat2sphere(double t, double c, const std::string &wavetp);
int surfdisp(const float *thk, const float *vp, const float *vs, const float *rho, int nlayer,
             const double *t, double *cg, int kmax, const std::string &wavetype, int mode,
             bool sphere, bool keep_flat);
int love_group(const float *thk, const float *vs, const float *rho, int nlayer, const double *t,
               double *cg, int kmax, int mode, bool sphere);
int rayleigh_group(const float *thk, const float *vp, const float *vs, const float *rho,
                   int nlayer, const double *t, double *cg, int kmax, int mode, bool sphere);
int surf_kernel(const float *thk, const float *vp, const float *vs, const float *rho, int nlayer,
                const double *t, double *c, int nt, double *dcda, double *dcdb, double *dcdr,
                double *dcdh, const std::string &wavetp, int mode, bool sphere,
                bool stale_first_term = true);

// ---- src/RF/RFModule.f90, deconit.f90, fftpack.f90
void rfft(const double *inp, cplx *out, int n);
void irfft(const cplx *inp, double *out, int n);
int nextpow2(int n);
void deconit(const double *u, const double *w, int nt, double dt, double tshift, double f0,
             double *out);
void cal_rf_time(const double *thk, const double *vp, const double *vs, const double *rho,
                 const double *qa, const double *qb, int nlayer, int nt, double dt, double ray_p,
                 double f0, double time_shift, int rf_type, double *rcv_fun);
void cal_rf_freq(const double *thk, const double *vp, const double *vs, const double *rho,
                 const double *qa, const double *qb, int nlayer, int nt, double dt, double ray_p,
                 double f0, double t0, double water, int rf_type, double *rcv_fun);
void cal_rf_par_freq(const double *thk, const double *vp, const double *vs, const double *rho,
                     const double *qa, const double *qb, int nlayer, int nt, double dt,
                     double ray_p, double f0, double t0, double water, int rf_type, int ipar,
                     double *rcv_fun, double *rcv_fun_p);
void cal_rf_par_freq_all(const double *thk, const double *vp, const double *vs, const double *rho,
                         const double *qa, const double *qb, int nlayer, int nt, double dt,
                         double ray_p, double f0, double t0, double water, int rf_type,
                         double *rcv_fun, double *rcv_fun_p);
void cal_rf_par_time(const double *thk, const double *vp, const double *vs, const double *rho,
                     const double *qa, const double *qb, int nlayer, int nt, double dt,
                     double ray_p, double f0, double time_shift, int rf_type, int ipar,
                     double *rcv_fun, double *rcv_fun_p);
void cal_rf_par_time_all(const double *thk, const double *vp, const double *vs, const double *rho,
                         const double *qa, const double *qb, int nlayer, int nt, double dt,
                         double ray_p, double f0, double time_shift, int rf_type, double *rcv_fun,
                         double *rcv_fun_p);

}  // namespace oracle
