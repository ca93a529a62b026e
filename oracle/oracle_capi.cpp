// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/README.md).
//
// C entry points (ctypes-loadable) over the restatement.  They mirror
//   /root/reference/src/SWD/main.cpp:14-82   (libsurf.forward / adjoint_kernel; float32 cast of
//                                             the model at the boundary, :7-9,14,62)
//   /root/reference/src/RF/main.cpp:17-189   (librf.forward / kernel / kernel_all; S-type negates
//                                             time_shift, :35,82,157)
// plus a C++ restatement of the Python glue on the hot path
//   /root/reference/model/model_surf.py:47-79,155-228
//   /root/reference/model/model_rf.py:52-77,137-197
//   /root/reference/model/model_rf_swd_vs_thk.py:66-86
// used as the checker for the fused GPU kernel and as the CPU baseline of bench.py.
#include "oracle.hpp"
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

using namespace oracle;

namespace {

const char *WT[4] = {"Rc", "Rg", "Lc", "Lg"};

void to_f32(const double *x, int n, std::vector<float> &o) {
  o.resize(n);
  for (int i = 0; i < n; i++) o[i] = (float)x[i];
}

// model_surf.py:64-73 / model_rf.py:68-75
void brocher(const double *vs, int n, double *vp, double *rho, double *dadb, double *drda) {
  for (int i = 0; i < n; i++) {
    double b = vs[i];
    double a = 0.9409 + 2.0947 * b - 0.8206 * b * b + 0.2683 * b * b * b - 0.0251 * b * b * b * b;
    vp[i] = a;
    rho[i] = 1.6612 * a - 0.4721 * a * a + 0.0671 * a * a * a - 0.0043 * a * a * a * a +
             0.000106 * a * a * a * a * a;
    if (drda)
      drda[i] = 1.6612 - 0.4721 * 2 * a + 0.0671 * 3 * a * a - 0.0043 * 4 * a * a * a +
                0.000106 * 5 * a * a * a * a;
    if (dadb) dadb[i] = 2.0947 - 0.8206 * 2 * b + 0.2683 * 3 * b * b - 0.0251 * 4 * b * b * b;
  }
}

struct JointCfg {
  int n;
  // swd
  int ntRc, ntRg, ntLc, ntLg;
  const double *tRc, *tRg, *tLc, *tLg;
  int mode, sphere;
  // rf
  double ray_p, dt, gauss, time_shift, water;
  int nt, rf_type, method;  // method 0 time, 1 freq
  double sigma1, sigma2;
  int stale;
};

// SurfWD.misfit_and_grad (model_surf.py:155-228). Returns flag.
int swd_misfit_grad(const JointCfg &c, const double *x, const double *dobs, double *U, double *grad,
                    double *d) {
  const int n = c.n;
  const int ntot = c.ntRc + c.ntRg + c.ntLc + c.ntLg;
  std::vector<double> vp(n), rho(n), dadb(n), drda(n);
  const double *vs = x, *thk = x + n;
  brocher(vs, n, vp.data(), rho.data(), dadb.data(), drda.data());
  std::vector<float> fthk, fvp, fvs, frho;
  to_f32(thk, n, fthk);
  to_f32(vp.data(), n, fvp);
  to_f32(vs, n, fvs);
  to_f32(rho.data(), n, frho);
  std::vector<double> kernel((size_t)ntot * n, 0.0), kthk((size_t)ntot * n, 0.0);
  for (int i = 0; i < ntot; i++) d[i] = 0.0;
  for (int i = 0; i < 2 * n; i++) grad[i] = 0.0;
  *U = 0.0;
  const int nts[4] = {c.ntRc, c.ntRg, c.ntLc, c.ntLg};
  // Quirk kept (model_surf.py:200-201,211-212): Lc and Lg are evaluated on tRc.
  const double *per[4] = {c.tRc, c.tRg, c.tRc, c.tRc};
  int k1 = 0;
  for (int w = 0; w < 4; w++) {
    const int nt = nts[w];
    if (nt <= 0) continue;
    std::vector<double> cg(nt), da((size_t)nt * n), db((size_t)nt * n), dr((size_t)nt * n),
        dh((size_t)nt * n);
    int ierr = surf_kernel(fthk.data(), fvp.data(), fvs.data(), frho.data(), n, per[w], cg.data(),
                           nt, da.data(), db.data(), dr.data(), dh.data(), WT[w], c.mode,
                           c.sphere != 0, c.stale != 0);
    if (ierr == 1) {
      // reference returns (0.0, zeros(n), partially filled d, False); we zero d for determinism
      return 0;
    }
    for (int k = 0; k < nt; k++) {
      d[k1 + k] = cg[k];
      for (int j = 0; j < n; j++) {
        size_t o = (size_t)k * n + j;
        kernel[(size_t)(k1 + k) * n + j] = db[o] + da[o] * dadb[j] + dr[o] * drda[j] * dadb[j];
        kthk[(size_t)(k1 + k) * n + j] = dh[o];
      }
    }
    k1 += nt;
  }
  double s = 0.0;
  for (int k = 0; k < ntot; k++) {
    double r = d[k] - dobs[k];
    s += r * r;
    for (int j = 0; j < n; j++) {
      grad[j] += r * kernel[(size_t)k * n + j];
      grad[n + j] += r * kthk[(size_t)k * n + j];
    }
  }
  *U = 0.5 * s;
  return 1;
}

// ReceiverFunc.misfit_and_grad (model_rf.py:137-197)
void rf_misfit_grad(const JointCfg &c, const double *x, const double *dobs, double *U, double *grad,
                    double *d) {
  const int n = c.n, nt = c.nt;
  std::vector<double> vp(n), rho(n), dadb(n), drda(n), qa(n, 9999.), qb(n, 9999.);
  const double *vs = x, *thk = x + n;
  brocher(vs, n, vp.data(), rho.data(), dadb.data(), drda.data());
  std::vector<double> kl((size_t)4 * n * nt);
  double tshift = c.time_shift;
  if (c.rf_type == 2) tshift = -tshift;
  if (c.method == 1)
    cal_rf_par_freq_all(thk, vp.data(), vs, rho.data(), qa.data(), qb.data(), n, nt, c.dt, c.ray_p,
                        c.gauss, tshift, c.water, c.rf_type, d, kl.data());
  else
    cal_rf_par_time_all(thk, vp.data(), vs, rho.data(), qa.data(), qb.data(), n, nt, c.dt, c.ray_p,
                        c.gauss, tshift, c.rf_type, d, kl.data());
  double s = 0.0;
  for (int it = 0; it < nt; it++) {
    double r = d[it] - dobs[it];
    s += r * r;
  }
  for (int j = 0; j < n; j++) {
    const double *krho = &kl[((size_t)0 * n + j) * nt], *kvp = &kl[((size_t)1 * n + j) * nt],
                 *kvs = &kl[((size_t)2 * n + j) * nt], *kth = &kl[((size_t)3 * n + j) * nt];
    double g1 = 0.0, g2 = 0.0;
    for (int it = 0; it < nt; it++) {
      double r = d[it] - dobs[it];
      double kk = kvs[it] + dadb[j] * kvp[it] + drda[j] * dadb[j] * krho[it];
      g1 += kk * r;
      g2 += kth[it] * r;
    }
    grad[j] = g1;
    grad[n + j] = g2;
  }
  *U = 0.5 * s;
}

// Joint_RF_SWD.misfit_and_grad (model_rf_swd_vs_thk.py:66-86).
// dobs = [rfobs(nt), swdobs(ntswd)].  On SWD failure: U=0, grad=0, dsyn=dobs, flag=0.
int joint_misfit_grad(const JointCfg &c, const double *x, const double *dobs, double *U,
                      double *grad, double *dsyn) {
  const int n = c.n, n1 = c.nt, n2 = c.ntRc + c.ntRg + c.ntLc + c.ntLg;
  std::vector<double> gr(2 * n), gs(2 * n), dr(n1), ds(n2);
  double Ur = 0, Us = 0;
  rf_misfit_grad(c, x, dobs, &Ur, gr.data(), dr.data());
  int flag = swd_misfit_grad(c, x, dobs + n1, &Us, gs.data(), ds.data());
  if (!flag) {
    *U = 0.0;
    for (int i = 0; i < 2 * n; i++) grad[i] = 0.0;
    for (int i = 0; i < n1 + n2; i++) dsyn[i] = dobs[i];
    return 0;
  }
  double q = c.sigma1 / c.sigma2;
  double wt = q * q * n1 / n2;
  *U = Ur + wt * Us;
  for (int i = 0; i < 2 * n; i++) grad[i] = gr[i] + wt * gs[i];
  for (int i = 0; i < n1; i++) dsyn[i] = dr[i];
  for (int i = 0; i < n2; i++) dsyn[n1 + i] = ds[i];
  return 1;
}

template <class F>
void parallel_for(int B, int nthreads, F f) {
  if (nthreads <= 1) {
    for (int b = 0; b < B; b++) f(b);
    return;
  }
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&]() {
      for (;;) {
        int b = next.fetch_add(1);
        if (b >= B) break;
        f(b);
      }
    });
  for (auto &t : th) t.join();
}

}  // namespace

extern "C" {

// libsurf.forward (main.cpp:14-59).  wavetype 0..3 = Rc,Rg,Lc,Lg.  Returns 1 ok / 0 fail / -2 bad arg.
int orc_surf_forward(const double *thk, const double *vp, const double *vs, const double *rho,
                     int n, const double *t, int nt, int wavetype, int mode, int sphere,
                     double *cg) {
  if (wavetype < 0 || wavetype > 3) return -2;
  std::vector<float> fthk, fvp, fvs, frho;
  to_f32(thk, n, fthk);
  to_f32(vp, n, fvp);
  to_f32(vs, n, fvs);
  to_f32(rho, n, frho);
  int ierr;
  if (wavetype == 0)
    ierr = surfdisp(fthk.data(), fvp.data(), fvs.data(), frho.data(), n, t, cg, nt, "Rc", mode,
                    sphere != 0, false);
  else if (wavetype == 1)
    ierr = rayleigh_group(fthk.data(), fvp.data(), fvs.data(), frho.data(), n, t, cg, nt, mode,
                          sphere != 0);
  else if (wavetype == 2)
    ierr = surfdisp(fthk.data(), fvp.data(), fvs.data(), frho.data(), n, t, cg, nt, "Lc", mode,
                    sphere != 0, false);
  else
    ierr = love_group(fthk.data(), fvs.data(), frho.data(), n, t, cg, nt, mode, sphere != 0);
  return ierr == 1 ? 0 : 1;
}

// libsurf.adjoint_kernel (main.cpp:61-82)
int orc_surf_kernel(const double *thk, const double *vp, const double *vs, const double *rho,
                    int n, const double *t, int nt, int wavetype, int mode, int sphere, int stale,
                    double *c, double *dcda, double *dcdb, double *dcdr, double *dcdh) {
  if (wavetype < 0 || wavetype > 3) return -2;
  std::vector<float> fthk, fvp, fvs, frho;
  to_f32(thk, n, fthk);
  to_f32(vp, n, fvp);
  to_f32(vs, n, fvs);
  to_f32(rho, n, frho);
  int ierr = surf_kernel(fthk.data(), fvp.data(), fvs.data(), frho.data(), n, t, c, nt, dcda, dcdb,
                         dcdr, dcdh, WT[wavetype], mode, sphere != 0, stale != 0);
  return ierr == 1 ? 0 : 1;
}

// number of secular-function evaluations surfdisp96 spends (op-count pin for DESIGN.md)
long orc_surfdisp96_evals(const double *thk, const double *vp, const double *vs, const double *rho,
                          int n, const double *t, int nt, int iwave, int mode1) {
  std::vector<float> fthk, fvp, fvs, frho;
  to_f32(thk, n, fthk);
  to_f32(vp, n, fvp);
  to_f32(vs, n, fvs);
  to_f32(rho, n, frho);
  std::vector<double> cg(nt);
  int ierr;
  long ne = 0;
  surfdisp96(fthk.data(), fvp.data(), fvs.data(), frho.data(), n, 0, iwave, mode1, 0, nt, t,
             cg.data(), &ierr, &ne);
  return ne;
}

// librf.forward (main.cpp:17-62).  method 0 time / 1 freq, rf_type 1 P / 2 S.
void orc_rf_forward(const double *thk, const double *rho, const double *vp, const double *vs,
                    const double *qa, const double *qb, int n, double ray_p, int nt, double dt,
                    double gauss, double time_shift, int method, double water, int rf_type,
                    double *rf) {
  if (rf_type == 2) time_shift = -time_shift;
  if (method == 0)
    cal_rf_time(thk, vp, vs, rho, qa, qb, n, nt, dt, ray_p, gauss, time_shift, rf_type, rf);
  else
    cal_rf_freq(thk, vp, vs, rho, qa, qb, n, nt, dt, ray_p, gauss, time_shift, water, rf_type, rf);
}

// librf.kernel (main.cpp:64-136). par_type 1 rho, 2 vp, 3 vs, 4 h.  drf[n][nt]
void orc_rf_kernel(const double *thk, const double *rho, const double *vp, const double *vs,
                   const double *qa, const double *qb, int n, double ray_p, int nt, double dt,
                   double gauss, double time_shift, int method, double water, int rf_type,
                   int par_type, double *rf, double *drf) {
  if (rf_type == 2) time_shift = -time_shift;
  if (method == 0)
    cal_rf_par_time(thk, vp, vs, rho, qa, qb, n, nt, dt, ray_p, gauss, time_shift, rf_type,
                    par_type, rf, drf);
  else
    cal_rf_par_freq(thk, vp, vs, rho, qa, qb, n, nt, dt, ray_p, gauss, time_shift, water, rf_type,
                    par_type, rf, drf);
}

// librf.kernel_all (main.cpp:140-189). drf[4][n][nt], order rho, vp, vs, h
void orc_rf_kernel_all(const double *thk, const double *rho, const double *vp, const double *vs,
                       const double *qa, const double *qb, int n, double ray_p, int nt, double dt,
                       double gauss, double time_shift, int method, double water, int rf_type,
                       double *rf, double *drf) {
  if (rf_type == 2) time_shift = -time_shift;
  if (method == 0)
    cal_rf_par_time_all(thk, vp, vs, rho, qa, qb, n, nt, dt, ray_p, gauss, time_shift, rf_type, rf,
                        drf);
  else
    cal_rf_par_freq_all(thk, vp, vs, rho, qa, qb, n, nt, dt, ray_p, gauss, time_shift, water,
                        rf_type, rf, drf);
}

void orc_rfft(const double *inp, double *out_re_im, int n) {
  std::vector<cplx> o(n / 2 + 1);
  rfft(inp, o.data(), n);
  for (int i = 0; i < n / 2 + 1; i++) {
    out_re_im[2 * i] = o[i].real();
    out_re_im[2 * i + 1] = o[i].imag();
  }
}
void orc_irfft(const double *inp_re_im, double *out, int n) {
  std::vector<cplx> a(n / 2 + 1);
  for (int i = 0; i < n / 2 + 1; i++) a[i] = cplx(inp_re_im[2 * i], inp_re_im[2 * i + 1]);
  irfft(a.data(), out, n);
}

// Batched Joint_RF_SWD.misfit_and_grad over B models, `nthreads` host threads.
// which: 0 joint, 1 RF only (ReceiverFunc.misfit_and_grad), 2 SWD only (SurfWD.misfit_and_grad).
void orc_joint_batch(int B, int n, const double *x, const double *dobs, int ntRc, const double *tRc,
                     int ntRg, const double *tRg, int ntLc, const double *tLc, int ntLg,
                     const double *tLg, int mode, int sphere, double ray_p, int nt, double dt,
                     double gauss, double time_shift, double water, int rf_type, int method,
                     double sigma1, double sigma2, int stale, int which, int nthreads, double *U,
                     double *grad, double *dsyn, int *flag) {
  JointCfg c;
  c.n = n;
  c.ntRc = ntRc;
  c.ntRg = ntRg;
  c.ntLc = ntLc;
  c.ntLg = ntLg;
  c.tRc = tRc;
  c.tRg = tRg;
  c.tLc = tLc;
  c.tLg = tLg;
  c.mode = mode;
  c.sphere = sphere;
  c.ray_p = ray_p;
  c.dt = dt;
  c.gauss = gauss;
  c.time_shift = time_shift;
  c.water = water;
  c.nt = nt;
  c.rf_type = rf_type;
  c.method = method;
  c.sigma1 = sigma1;
  c.sigma2 = sigma2;
  c.stale = stale;
  const int nsw = ntRc + ntRg + ntLc + ntLg;
  const int nd = (which == 0) ? nt + nsw : (which == 1 ? nt : nsw);
  parallel_for(B, nthreads, [&](int b) {
    const double *xb = x + (size_t)b * 2 * n;
    double *gb = grad + (size_t)b * 2 * n, *db = dsyn + (size_t)b * nd;
    if (which == 0)
      flag[b] = joint_misfit_grad(c, xb, dobs, &U[b], gb, db);
    else if (which == 1) {
      rf_misfit_grad(c, xb, dobs, &U[b], gb, db);
      flag[b] = 1;
    } else
      flag[b] = swd_misfit_grad(c, xb, dobs, &U[b], gb, db);
  });
}

}  // extern "C"
