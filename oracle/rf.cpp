// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/README.md).
//
// Restates the receiver-function path of the reference:
//   /root/reference/src/RF/RFModule.f90
//     cal_rf_par_time :11-74   cal_rf_par_time_all :76-142   cal_rf_time :144-191
//     cal_rf_freq :193-255     cal_rf_par_freq :258-341      cal_rf_par_freq_all :343-430
//     cal_response :432-478    cal_response_par :481-589     cal_response_par_all :592-707
//     cal_matrix_a :709-764    cal_matrix_a_par :766-879     cal_E_inv :881-922
//     cal_E_inv_par :924-987
//   /root/reference/src/RF/deconit.f90 (all), /root/reference/src/RF/fftpack.f90 (all)
// FFTW3 (external, absent here) is replaced by a plain radix-2 FFT with FFTW's r2c/c2r
// conventions (forward e^{-i..}, unnormalised; c2r ignores Im of the DC and Nyquist bins).
// Deliberately keeps the reference's O(n^2) product structure: this is the line-faithful
// restatement, not the optimised algorithm.
// Indeterminate reference behaviour resolved here: `va_k` is used uninitialised in
// cal_E_inv_par for ipars==2 (RFModule.f90:933,980); the oracle uses the intended formula of
// the commented line :978-979, va_k = sqrt(p^2 - 1/alpha^2)/p.
#include "oracle.hpp"
#include <cmath>
#include <cstring>
#include <vector>

namespace oracle {

namespace {

const cplx imag_i(0.0, 1.0);

// in-place iterative radix-2, sign = -1 forward, +1 backward, unnormalised
void fft_pow2(std::vector<cplx> &a, int sign) {
  const int n = (int)a.size();
  for (int i = 1, j = 0; i < n; i++) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  for (int len = 2; len <= n; len <<= 1) {
    double ang = sign * 2.0 * M_PI / len;
    for (int i = 0; i < n; i += len) {
      for (int k = 0; k < len / 2; k++) {
        cplx w(std::cos(ang * k), std::sin(ang * k));
        cplx u = a[i + k], v = a[i + k + len / 2] * w;
        a[i + k] = u + v;
        a[i + k + len / 2] = u - v;
      }
    }
  }
}

typedef cplx M4[4][4];

void matmul4(const M4 a, const M4 b, M4 out) {
  M4 t;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      cplx s = 0.0;
      for (int k = 0; k < 4; k++) s += a[i][k] * b[k][j];
      t[i][j] = s;
    }
  std::memcpy(out, t, sizeof(M4));
}

void eye4(M4 a) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) a[i][j] = (i == j) ? 1.0 : 0.0;
}

void scale4(M4 a, cplx s) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) a[i][j] *= s;
}

#define A(i, j) m[(i)-1][(j)-1]

// RFModule.f90:709-764
void cal_matrix_a(cplx omega, double ray_p, double thick, cplx alpha, cplx beta, double rho, M4 m) {
  cplx miu = rho * beta * beta;
  cplx k = omega * ray_p;
  cplx k_alpha = omega / alpha, k_beta = omega / beta;
  cplx v_alpha = std::sqrt(k * k - k_alpha * k_alpha);
  cplx v_beta = std::sqrt(k * k - k_beta * k_beta);
  cplx va_k = std::sqrt(ray_p * ray_p - 1.0 / (alpha * alpha)) / ray_p;
  cplx vb_k = std::sqrt(ray_p * ray_p - 1.0 / (beta * beta)) / ray_p;
  cplx gamma = 2.0 * ray_p * ray_p * beta * beta;
  cplx gamma1 = 1.0 - 1.0 / gamma;
  cplx c_a = std::cosh(v_alpha * thick);
  cplx x_a = va_k * std::sinh(v_alpha * thick);
  cplx y_a = std::sinh(v_alpha * thick) / va_k;
  cplx c_b = std::cosh(v_beta * thick);
  cplx x_b = vb_k * std::sinh(v_beta * thick);
  cplx y_b = std::sinh(v_beta * thick) / vb_k;
  A(1, 1) = c_a - gamma1 * c_b;
  A(1, 2) = gamma1 * y_a - x_b;
  A(1, 3) = (c_b - c_a) / 2.0 / miu;
  A(1, 4) = (x_b - y_a) / 2.0 / miu;
  A(2, 1) = gamma1 * y_b - x_a;
  A(2, 2) = c_b - gamma1 * c_a;
  A(2, 3) = (x_a - y_b) / 2.0 / miu;
  A(2, 4) = (c_a - c_b) / 2.0 / miu;
  A(3, 1) = 2.0 * miu * gamma1 * (c_a - c_b);
  A(3, 2) = 2.0 * miu * (gamma1 * gamma1 * y_a - x_b);
  A(3, 3) = c_b - gamma1 * c_a;
  A(3, 4) = x_b - gamma1 * y_a;
  A(4, 1) = 2.0 * miu * (gamma1 * gamma1 * y_b - x_a);
  A(4, 2) = 2.0 * miu * gamma1 * (c_b - c_a);
  A(4, 3) = x_a - gamma1 * y_b;
  A(4, 4) = c_a - gamma1 * c_b;
  scale4(m, gamma);
}

// RFModule.f90:766-879   ipars: 1 rho, 2 vp, 3 vs, 4 thickness
void cal_matrix_a_par(cplx omega, double ray_p, double thick, cplx alpha, cplx beta, double rho,
                      M4 m, int ipars) {
  cplx miu = rho * beta * beta;
  cplx k = omega * ray_p;
  cplx k_alpha = omega / alpha, k_beta = omega / beta;
  cplx v_alpha = std::sqrt(k * k - k_alpha * k_alpha);
  cplx v_beta = std::sqrt(k * k - k_beta * k_beta);
  cplx gamma = 2. * ray_p * ray_p * beta * beta;
  cplx gamma1 = 1. - 1. / gamma;
  cplx gamma2 = gamma / ((alpha * ray_p) * (alpha * ray_p));
  cplx gamma3 = 1. / (gamma - 2.0);
  cplx va_k = std::sqrt(ray_p * ray_p - 1.0 / (alpha * alpha)) / ray_p;
  cplx vb_k = std::sqrt(ray_p * ray_p - 1.0 / (beta * beta)) / ray_p;
  cplx c_a = std::cosh(v_alpha * thick);
  cplx x_a = va_k * std::sinh(v_alpha * thick);
  cplx y_a = std::sinh(v_alpha * thick) / va_k;
  cplx c_b = std::cosh(v_beta * thick);
  cplx x_b = vb_k * std::sinh(v_beta * thick);
  cplx y_b = std::sinh(v_beta * thick) / vb_k;
  const cplx kt = k * thick;
  const cplx g1sq = gamma1 * gamma1;
  const cplx vak2 = va_k * va_k;
  if (ipars == 3) {
    A(1, 1) = 2. / beta * (gamma * (c_a - c_b) - gamma1 * kt * y_b);
    A(1, 2) = 2. / beta * (gamma * (y_a - x_b) - (kt * c_b + y_b));
    A(1, 3) = kt * y_b / miu / beta;
    A(1, 4) = (kt * c_b + y_b) / miu / beta;
    A(2, 1) = ((y_b - x_a) + gamma1 * gamma3 * (kt * c_b - y_b)) * 2.0 * gamma / beta;
    A(2, 2) = 2. / beta * (gamma * (c_b - c_a) + kt * y_b);
    A(2, 3) = -(kt * c_b - y_b) * gamma * gamma3 / miu / beta;
    A(2, 4) = -A(1, 3);
    A(3, 1) = 4. * miu / beta * ((2.0 * gamma - 1.0) * (c_a - c_b) - gamma1 * kt * y_b);
    A(3, 2) = 4. * miu / beta * ((2.0 * gamma) * (gamma1 * y_a - x_b) - (kt * c_b + y_b));
    A(3, 3) = A(2, 2);
    A(3, 4) = -A(1, 2);
    A(4, 1) = 4. * miu * gamma / beta *
              (2.0 * gamma1 * y_b - 2.0 * x_a + g1sq * gamma3 * (kt * c_b - y_b));
    A(4, 2) = -A(3, 1);
    A(4, 3) = -A(2, 1);
    A(4, 4) = A(1, 1);
  } else if (ipars == 2) {
    A(1, 1) = kt * y_a * gamma2 / alpha;
    A(1, 2) = 1.0 / vak2 / alpha * gamma1 * gamma2 * (kt * c_a - y_a);
    A(1, 3) = -kt * y_a * gamma2 / 2.0 / miu / alpha;
    A(1, 4) = -1.0 / vak2 * (kt * c_a - y_a) * gamma2 / 2.0 / miu / alpha;
    A(2, 1) = -(kt * c_a + y_a) * gamma2 / alpha;
    A(2, 2) = -kt * y_a * gamma1 * gamma2 / alpha;
    A(2, 3) = (kt * c_a + y_a) * gamma2 / 2.0 / miu / alpha;
    A(2, 4) = kt * y_a * gamma2 / 2.0 / miu / alpha;
    A(3, 1) = kt * y_a * gamma1 * gamma2 * 2.0 * miu / alpha;
    A(3, 2) = 2. * miu / alpha * g1sq * gamma2 * (kt * c_a - y_a) / vak2;
    A(3, 3) = -kt * y_a * gamma1 * gamma2 / alpha;
    A(3, 4) = -1. / alpha / vak2 * (kt * c_a - y_a) * gamma1 * gamma2;
    A(4, 1) = -2. * miu / alpha * (kt * c_a + y_a) * gamma2;
    A(4, 2) = -2. * miu / alpha * kt * y_a * gamma1 * gamma2;
    A(4, 3) = (kt * c_a + y_a) * gamma2 / alpha;
    A(4, 4) = kt * y_a / alpha * gamma2;
  } else if (ipars == 1) {
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) m[i][j] = 0.0;
    A(1, 3) = -gamma / (2.0 * rho * miu) * (-c_a + c_b);
    A(1, 4) = -gamma / (2.0 * rho * miu) * (-y_a + x_b);
    A(2, 3) = -gamma / (2.0 * rho * miu) * (x_a - y_b);
    A(2, 4) = -gamma / (2.0 * rho * miu) * (c_a - c_b);
    A(3, 1) = 2. * miu * gamma * gamma1 / rho * (c_a - c_b);
    A(3, 2) = 2. * miu * gamma / rho * (g1sq * y_a - x_b);
    A(4, 1) = 2. * miu * gamma / rho * (-x_a + g1sq * y_b);
    A(4, 2) = 2. * miu * gamma * gamma1 / rho * (-c_a + c_b);
  } else {
    A(1, 1) = (x_a - gamma1 * x_b) * k;
    A(1, 2) = gamma1 * k * c_a - v_beta * vb_k * c_b;
    A(1, 3) = (x_b - x_a) * k / 2.0 / miu;
    A(1, 4) = (v_beta * vb_k * c_b - k * c_a) / 2.0 / miu;
    A(2, 1) = gamma1 * k * c_b - v_alpha * va_k * c_a;
    A(2, 2) = (x_b - gamma1 * x_a) * k;
    A(2, 3) = (v_alpha * va_k * c_a - k * c_b) / 2.0 / miu;
    A(2, 4) = (x_a - x_b) * k / 2.0 / miu;
    A(3, 1) = 2. * miu * gamma1 * k * (x_a - x_b);
    A(3, 2) = 2. * miu * (g1sq * k * c_a - v_beta * vb_k * c_b);
    A(3, 3) = (x_b - gamma1 * x_a) * k;
    A(3, 4) = v_beta * vb_k * c_b - gamma1 * k * c_a;
    A(4, 1) = 2. * miu * (g1sq * k * c_b - v_alpha * va_k * c_a);
    A(4, 2) = 2. * miu * gamma1 * k * (x_b - x_a);
    A(4, 3) = v_alpha * va_k * c_a - gamma1 * k * c_b;
    A(4, 4) = (x_a - gamma1 * x_b) * k;
    scale4(m, gamma);
  }
}

// RFModule.f90:881-922
void cal_E_inv(cplx omega, double ray_p, cplx alpha, cplx beta, double rho, M4 m) {
  (void)omega;
  cplx miu = rho * beta * beta;
  cplx gamma = 2.0 * ray_p * ray_p * beta * beta;
  cplx gamma1 = 1.0 - 1.0 / gamma;
  cplx va_k = std::sqrt(ray_p * ray_p - cplx(1.0, 0.0) / (alpha * alpha)) / ray_p;
  cplx vb_k = std::sqrt(ray_p * ray_p - cplx(1.0, 0.0) / (beta * beta)) / ray_p;
  A(1, 1) = -1.0;
  A(1, 2) = -gamma1 / va_k;
  A(1, 3) = 1.0 / (2. * miu);
  A(1, 4) = 1.0 / (2. * miu * va_k);
  A(2, 1) = gamma1 / vb_k;
  A(2, 2) = 1.0;
  A(2, 3) = -1.0 / (2. * miu * vb_k);
  A(2, 4) = -1.0 / (2. * miu);
  A(3, 1) = 1.0;
  A(3, 2) = -gamma1 / va_k;
  A(3, 3) = -1.0 / (2. * miu);
  A(3, 4) = 1.0 / (2. * miu * va_k);
  A(4, 1) = -gamma1 / vb_k;
  A(4, 2) = 1.0;
  A(4, 3) = 1. / (2. * miu * vb_k);
  A(4, 4) = -1.0 / (2. * miu);
  scale4(m, 0.5 * gamma);
}

// RFModule.f90:924-987
void cal_E_inv_par(cplx omega, double ray_p, cplx alpha, cplx beta, double rho, M4 m, int ipars) {
  cplx miu = rho * beta * beta;
  cplx k = omega * ray_p;
  cplx k_alpha = omega / alpha, k_beta = omega / beta;
  cplx v_alpha = std::sqrt(k * k - k_alpha * k_alpha);
  cplx v_beta = std::sqrt(k * k - k_beta * k_beta);
  cplx gamma = 2.0 * k * k * beta * beta / (omega * omega);
  cplx gamma1 = 1.0 - 1.0 / gamma;
  cplx gamma3 = 1.0 / (gamma - 2.0);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) m[i][j] = 0.0;
  if (ipars == 3) {
    A(1, 1) = -1.0;
    A(1, 2) = -k / v_alpha;
    A(2, 1) = k / v_beta * (1.0 - gamma1 * gamma3);
    A(2, 2) = 1.0;
    A(2, 3) = k * gamma3 / 2.0 / miu / v_beta;
    A(3, 1) = 1.0;
    A(3, 2) = -k / v_alpha;
    A(4, 1) = -k / v_beta * (1.0 - gamma3 * gamma1);
    A(4, 2) = 1.0;
    A(4, 3) = -k * gamma3 / 2.0 / miu / v_beta;
    scale4(m, gamma / beta);
  } else if (ipars == 1) {
    A(1, 3) = -1.0;
    A(1, 4) = -k / v_alpha;
    A(2, 3) = k / v_beta;
    A(2, 4) = 1.0;
    A(3, 3) = 1.0;
    A(3, 4) = -k / v_alpha;
    A(4, 3) = -k / v_beta;
    A(4, 4) = 1.0;
    scale4(m, gamma / 4.0 / rho / miu);
  } else if (ipars == 2) {
    // reference reads an uninitialised va_k here; intended value per the commented line :978-979
    cplx va_k = std::sqrt(ray_p * ray_p - 1.0 / (alpha * alpha)) / ray_p;
    A(1, 2) = gamma1;
    A(1, 4) = -0.5 / miu;
    A(3, 2) = gamma1;
    A(3, 4) = -0.5 / miu;
    scale4(m, beta * beta / (alpha * alpha * alpha) / (va_k * va_k * va_k));
  }
  // ipars == 4: zeros
}
#undef A

inline bool cnan(cplx z) { return std::isnan(std::abs(z)); }

void pick(const M4 a, int rf_type, cplx &R21, cplx &R22) {
  if (rf_type == 1) {
    R22 = a[1][1] * imag_i;
    R21 = a[1][0];
  } else {
    R22 = -a[0][0] * imag_i;
    R21 = a[0][1];
  }
}

// RFModule.f90:432-478
void cal_response(cplx omega, double ray_p, const double *thk, const cplx *alpha,
                  const cplx *beta, const double *rho, int nlayer, int rf_type, cplx &R21,
                  cplx &R22) {
  M4 a_syn, a1, einv;
  eye4(a_syn);
  for (int ilayer = 1; ilayer <= nlayer - 1; ilayer++) {
    int li = nlayer - ilayer - 1;  // 0-based ilayer_inv
    cal_matrix_a(omega, ray_p, thk[li], alpha[li], beta[li], rho[li], a1);
    matmul4(a_syn, a1, a_syn);
  }
  cal_E_inv(omega, ray_p, alpha[nlayer - 1], beta[nlayer - 1], rho[nlayer - 1], einv);
  matmul4(einv, a_syn, a_syn);
  pick(a_syn, rf_type, R21, R22);
}

// RFModule.f90:592-707 (npars = 4; R21_m/R22_m indexed [ipar-1][layer])
// When only_ipar > 0 it behaves as cal_response_par (:481-589) for that parameter.
void cal_response_par_all(cplx omega, double ray_p, const double *thk, const cplx *alpha,
                          const cplx *beta, const double *vp, const double *vs,
                          const double *rho, int nlayer, int rf_type, cplx &R21, cplx &R22,
                          cplx *R21_m, cplx *R22_m, int only_ipar) {
  std::vector<cplx> all_a((size_t)16 * nlayer), all_a_m((size_t)16 * 4 * nlayer);
  auto Aof = [&](int il) { return reinterpret_cast<cplx(*)[4]>(&all_a[(size_t)16 * il]); };
  auto AMof = [&](int ip, int il) {
    return reinterpret_cast<cplx(*)[4]>(&all_a_m[(size_t)16 * (ip * nlayer + il)]);
  };
  M4 einv, einv_par[4], a_syn, a_syn_m;
  for (int il = 0; il < nlayer - 1; il++) {
    cal_matrix_a(omega, ray_p, thk[il], alpha[il], beta[il], rho[il], Aof(il));
    for (int ipar = 1; ipar <= 4; ipar++) {
      if (only_ipar > 0 && ipar != only_ipar) continue;
      cal_matrix_a_par(omega, ray_p, thk[il], alpha[il], beta[il], rho[il], AMof(ipar - 1, il), ipar);
      if (ipar == 3)
        scale4(AMof(ipar - 1, il), beta[il] / vs[il]);
      else if (ipar == 2)
        scale4(AMof(ipar - 1, il), alpha[il] / vp[il]);
    }
  }
  const int nl = nlayer - 1;
  cal_E_inv(omega, ray_p, alpha[nl], beta[nl], rho[nl], einv);
  for (int ipar = 1; ipar <= 4; ipar++) {
    if (only_ipar > 0 && ipar != only_ipar) continue;
    cal_E_inv_par(omega, ray_p, alpha[nl], beta[nl], rho[nl], einv_par[ipar - 1], ipar);
    if (ipar == 3)
      scale4(einv_par[ipar - 1], beta[nl] / vs[nl]);
    else if (ipar == 2)
      scale4(einv_par[ipar - 1], alpha[nl] / vp[nl]);
  }
  eye4(a_syn);
  for (int ilayer = 1; ilayer <= nlayer - 1; ilayer++) matmul4(a_syn, Aof(nlayer - ilayer - 1), a_syn);
  matmul4(einv, a_syn, a_syn);
  pick(a_syn, rf_type, R21, R22);
  if (cnan(R22)) R22 = 0.0;
  if (cnan(R21)) R21 = 0.0;
  for (int ipar = 1; ipar <= 4; ipar++) {
    if (only_ipar > 0 && ipar != only_ipar) continue;
    for (int par_layer = 1; par_layer <= nlayer; par_layer++) {
      eye4(a_syn_m);
      for (int ilayer = 1; ilayer <= nlayer - 1; ilayer++) {
        int ilayer_inv = nlayer - ilayer;
        if (par_layer == ilayer_inv)
          matmul4(a_syn_m, AMof(ipar - 1, ilayer_inv - 1), a_syn_m);
        else
          matmul4(a_syn_m, Aof(ilayer_inv - 1), a_syn_m);
      }
      if (par_layer == nlayer)
        matmul4(einv_par[ipar - 1], a_syn_m, a_syn_m);
      else
        matmul4(einv, a_syn_m, a_syn_m);
      cplx r21, r22;
      pick(a_syn_m, rf_type, r21, r22);
      if (cnan(r22)) r22 = 0.0;
      if (cnan(r21)) r21 = 0.0;
      R22_m[(size_t)(ipar - 1) * nlayer + par_layer - 1] = r22;
      R21_m[(size_t)(ipar - 1) * nlayer + par_layer - 1] = r21;
    }
  }
}

void complex_velocities(const double *vp, const double *vs, const double *qa, const double *qb,
                        int n, std::vector<cplx> &alpha, std::vector<cplx> &beta) {
  alpha.resize(n);
  beta.resize(n);
  for (int i = 0; i < n; i++) {
    alpha[i] = vp[i] * (1.0 + imag_i / (2.0 * qa[i]) + 1.0 / (8.0 * qa[i] * qa[i]));
    beta[i] = vs[i] * (1.0 + imag_i / (2.0 * qb[i]) + 1.0 / (8.0 * qb[i] * qb[i]));
  }
}

// shared tail of the freq-domain routines (RFModule.f90:235-250, 305-336, 392-425)
void freq_rf_core(const double *thk, const double *vp, const double *vs, const double *rho,
                  const double *qa, const double *qb, int nlayer, int nt, double dt, double ray_p,
                  double f0, double t0, double water, int rf_type, int only_ipar, bool want_par,
                  double *rcv_fun, double *rcv_fun_p) {
  const int nft = nextpow2(nt), n2 = nft / 2 + 1;
  std::vector<cplx> alpha, beta;
  complex_velocities(vp, vs, qa, qb, nlayer, alpha, beta);
  const double sigma = 1.0 / dt / nft * 4.;
  std::vector<cplx> R22(n2), R21(n2), spec(n2);
  std::vector<cplx> R22_m, R21_m;
  if (want_par) {
    R22_m.assign((size_t)n2 * 4 * nlayer, 0.0);
    R21_m.assign((size_t)n2 * 4 * nlayer, 0.0);
  }
  std::vector<double> w(n2), gauss(n2), wa(n2), fai(n2), tmp(nft);
  for (int it = 1; it <= n2; it++) {
    if (want_par)
      w[it - 1] = 1.0 / nft / dt * (it - 1) * 2.0 * PI32;
    else
      w[it - 1] = (1.0 / dt / nft) * (it - 1) * 2 * PI32;
    cplx omega(w[it - 1], -sigma);
    if (want_par)
      cal_response_par_all(omega, ray_p, thk, alpha.data(), beta.data(), vp, vs, rho, nlayer,
                           rf_type, R21[it - 1], R22[it - 1], &R21_m[(size_t)(it - 1) * 4 * nlayer],
                           &R22_m[(size_t)(it - 1) * 4 * nlayer], only_ipar);
    else
      cal_response(omega, ray_p, thk, alpha.data(), beta.data(), rho, nlayer, rf_type, R21[it - 1],
                   R22[it - 1]);
  }
  double wmax = 0.0;
  for (int i = 0; i < n2; i++) {
    double x = w[i] / 2 / f0;
    gauss[i] = std::exp(-(x * x));
    wa[i] = (R21[i] * std::conj(R21[i])).real();
    if (i == 0 || wa[i] > wmax) wmax = wa[i];
  }
  for (int i = 0; i < n2; i++) {
    fai[i] = std::fmax(wa[i], water * wmax);
    spec[i] = std::conj(R21[i]) * R22[i] * gauss[i] * std::exp(-imag_i * w[i] * t0) / fai[i];
  }
  irfft(spec.data(), tmp.data(), nft);
  for (int it = 1; it <= nt; it++)
    rcv_fun[it - 1] = tmp[it - 1] / dt * std::exp(sigma * (-t0 + (it - 1) * dt));
  if (!want_par) return;
  std::vector<cplx> R21sq(n2);
  for (int i = 0; i < n2; i++) {
    R21sq[i] = R21[i] * R21[i];
    wa[i] = (R21sq[i] * std::conj(R21sq[i])).real();
    if (i == 0 || wa[i] > wmax) wmax = wa[i];
  }
  for (int i = 0; i < n2; i++) fai[i] = std::fmax(wa[i], water * wmax);
  for (int ipar = 1; ipar <= 4; ipar++) {
    if (only_ipar > 0 && ipar != only_ipar) continue;
    for (int pl = 0; pl < nlayer; pl++) {
      for (int i = 0; i < n2; i++) {
        size_t o = (size_t)i * 4 * nlayer + (size_t)(ipar - 1) * nlayer + pl;
        spec[i] = std::conj(R21sq[i]) * (R22_m[o] * R21[i] - R21_m[o] * R22[i]) * gauss[i] *
                  std::exp(-imag_i * w[i] * t0) / fai[i];
      }
      irfft(spec.data(), tmp.data(), nft);
      // output layout: Fortran (nt, nlayer[, npars]) == C [npars][nlayer][nt]
      double *dst = (only_ipar > 0) ? rcv_fun_p + (size_t)pl * nt
                                    : rcv_fun_p + ((size_t)(ipar - 1) * nlayer + pl) * nt;
      for (int it = 1; it <= nt; it++)
        dst[it - 1] = tmp[it - 1] / dt * std::exp(sigma * (-t0 + (it - 1) * dt));
    }
  }
}

// time-domain variants: RFModule.f90:11-74, 76-142, 144-191
void time_rf_core(const double *thk, const double *vp, const double *vs, const double *rho,
                  const double *qa, const double *qb, int nlayer, int nt, double dt, double ray_p,
                  double f0, double time_shift, int rf_type, int only_ipar, bool want_par,
                  bool pi_is_double, double *rcv_fun, double *rcv_fun_p) {
  const int nft = nextpow2(nt), n2 = nft / 2 + 1;
  std::vector<cplx> alpha, beta;
  complex_velocities(vp, vs, qa, qb, nlayer, alpha, beta);
  std::vector<cplx> R22(n2), R21(n2), num(n2);
  std::vector<cplx> R22_m, R21_m;
  if (want_par) {
    R22_m.assign((size_t)n2 * 4 * nlayer, 0.0);
    R21_m.assign((size_t)n2 * 4 * nlayer, 0.0);
  }
  const double PI = pi_is_double ? (std::atan(1.0) * 4.0) : PI32;  // :94 uses atan(1.0_dp)
  for (int it = 1; it <= n2; it++) {
    cplx omega;
    if (want_par)
      omega = 1.0 / nft / dt * (it - 1) * 2.0 * PI;
    else
      omega = (1.0 / dt / nft) * (it - 1) * 2 * PI;
    if (want_par)
      cal_response_par_all(omega, ray_p, thk, alpha.data(), beta.data(), vp, vs, rho, nlayer,
                           rf_type, R21[it - 1], R22[it - 1], &R21_m[(size_t)(it - 1) * 4 * nlayer],
                           &R22_m[(size_t)(it - 1) * 4 * nlayer], only_ipar);
    else
      cal_response(omega, ray_p, thk, alpha.data(), beta.data(), rho, nlayer, rf_type, R21[it - 1],
                   R22[it - 1]);
  }
  std::vector<double> ux(nft), uz(nft), tmp(nft);
  irfft(R22.data(), ux.data(), nft);
  irfft(R21.data(), uz.data(), nft);
  deconit(ux.data(), uz.data(), nft, dt, time_shift, f0, tmp.data());
  for (int i = 0; i < nt; i++) rcv_fun[i] = tmp[i];
  if (!want_par) return;
  std::vector<cplx> R21sq(n2);
  for (int i = 0; i < n2; i++) R21sq[i] = R21[i] * R21[i];
  irfft(R21sq.data(), uz.data(), nft);
  for (int ipar = 1; ipar <= 4; ipar++) {
    if (only_ipar > 0 && ipar != only_ipar) continue;
    for (int pl = 0; pl < nlayer; pl++) {
      for (int i = 0; i < n2; i++) {
        size_t o = (size_t)i * 4 * nlayer + (size_t)(ipar - 1) * nlayer + pl;
        num[i] = R22_m[o] * R21[i] - R21_m[o] * R22[i];
      }
      irfft(num.data(), ux.data(), nft);
      deconit(ux.data(), uz.data(), nft, dt, time_shift, f0, tmp.data());
      double *dst = (only_ipar > 0) ? rcv_fun_p + (size_t)pl * nt
                                    : rcv_fun_p + ((size_t)(ipar - 1) * nlayer + pl) * nt;
      for (int i = 0; i < nt; i++) dst[i] = tmp[i];
    }
  }
}

// deconit.f90:15-32
void gauss_filter(int nt, double dt, double f0, std::vector<double> &gauss) {
  gauss.resize(nt / 2 + 1);
  for (int i = 1; i <= nt / 2 + 1; i++) {
    double freq = (i - 1) / (nt * dt);
    double x = 2 * PI32 * freq / f0;
    gauss[i - 1] = std::exp(-0.25 * (x * x));
  }
}

// deconit.f90:34-52
void apply_gaussian(double *mydata, double dt, double f0, int nt) {
  std::vector<double> gauss;
  std::vector<cplx> df(nt / 2 + 1);
  gauss_filter(nt, dt, f0, gauss);
  rfft(mydata, df.data(), nt);
  for (int i = 0; i < nt / 2 + 1; i++) df[i] *= gauss[i];
  irfft(df.data(), mydata, nt);
}

// deconit.f90:54-72 — `cmplx(0,1.0)` is a default (REAL*4) complex; the product
// (i-1)/(n*dt)*pi*2*tshift is evaluated in double and multiplied by it.
void shift_data(double *mydata, double dt, double tshift, int n) {
  std::vector<cplx> df(n / 2 + 1);
  rfft(mydata, df.data(), n);
  for (int i = 1; i <= n / 2 + 1; i++) {
    cplx arg = -cplx(0.0, 1.0) * (double)(i - 1) / (n * dt) * PI32 * 2.0 * tshift;
    df[i - 1] = df[i - 1] * std::exp(arg);
  }
  irfft(df.data(), mydata, n);
}

void mycorrelate(const double *a, const double *b, double *out, int n) {
  std::vector<cplx> aft(n / 2 + 1), bft(n / 2 + 1), c(n / 2 + 1);
  rfft(a, aft.data(), n);
  rfft(b, bft.data(), n);
  for (int i = 0; i < n / 2 + 1; i++) c[i] = aft[i] * std::conj(bft[i]);
  irfft(c.data(), out, n);
}

void myconvolve(const double *a, const double *b, double *out, int n) {
  std::vector<cplx> aft(n / 2 + 1), bft(n / 2 + 1), c(n / 2 + 1);
  rfft(a, aft.data(), n);
  rfft(b, bft.data(), n);
  for (int i = 0; i < n / 2 + 1; i++) c[i] = aft[i] * bft[i];
  irfft(c.data(), out, n);
}

}  // namespace

// deconit.f90:1-13
int nextpow2(int n) {
  int nout = 1;
  while (nout < n) nout *= 2;
  return nout;
}

// fftpack.f90:1-21 (FFTW r2c)
void rfft(const double *inp, cplx *out, int n) {
  std::vector<cplx> a(n);
  for (int i = 0; i < n; i++) a[i] = inp[i];
  if (n > 1) fft_pow2(a, -1);
  for (int i = 0; i < n / 2 + 1; i++) out[i] = a[i];
}

// fftpack.f90:23-43 (FFTW c2r, then /n)
void irfft(const cplx *inp, double *out, int n) {
  std::vector<cplx> a(n);
  const int h = n / 2;
  a[0] = cplx(inp[0].real(), 0.0);
  for (int i = 1; i < h; i++) {
    a[i] = inp[i];
    a[n - i] = std::conj(inp[i]);
  }
  if (n > 1) a[h] = cplx(inp[h].real(), 0.0);
  if (n > 1) fft_pow2(a, +1);
  for (int i = 0; i < n; i++) out[i] = a[i].real() / n;
}

// deconit.f90:135-198
void deconit(const double *u, const double *w, int nt, double dt, double tshift, double f0,
             double *out) {
  const int nft = nextpow2(nt);
  std::vector<double> uflt(nft, 0.0), wflt(nft, 0.0), wcopy, p(nft, 0.0), rflt, cuw(nft),
      temp1(nft), temp2(nft);
  for (int i = 0; i < nt; i++) {
    wflt[i] = w[i];
    uflt[i] = u[i];
  }
  wcopy = wflt;
  apply_gaussian(uflt.data(), dt, f0, nft);
  apply_gaussian(wflt.data(), dt, f0, nft);
  double sw = 0.0, su = 0.0;
  for (int i = 0; i < nft; i++) {
    sw += wflt[i] * wflt[i];
    su += uflt[i] * uflt[i];
  }
  double invpw = 1. / sw / dt;
  double invpu = 1. / su / dt;
  double sumsq_i = 1.0, sumsq = 50;
  const double minderr = (double)0.001f;  // REAL*4 literal assigned to real(8)
  double d_error = 100 * invpw + minderr;
  rflt = uflt;
  const int maxiter = 200;
  for (int it = 1; it <= maxiter; it++) {
    if (std::fabs(d_error) <= minderr) break;
    mycorrelate(rflt.data(), wflt.data(), cuw.data(), nft);
    for (int i = 0; i < nft; i++) cuw[i] *= dt;
    int idx = 0;
    double best = -1.0;
    for (int i = 0; i < nft / 2; i++) {  // maxloc: first maximum
      double v = std::fabs(cuw[i]);
      if (v > best) {
        best = v;
        idx = i;
      }
    }
    p[idx] = p[idx] + cuw[idx] * invpw / dt;
    temp1 = p;
    apply_gaussian(temp1.data(), dt, f0, nft);
    myconvolve(temp1.data(), wcopy.data(), temp2.data(), nft);
    double s = 0.0;
    for (int i = 0; i < nft; i++) {
      rflt[i] = uflt[i] - temp2[i] * dt;
      s += rflt[i] * rflt[i];
    }
    sumsq = s * dt * invpu;
    d_error = 100. * (sumsq_i - sumsq);
    sumsq_i = sumsq;
  }
  apply_gaussian(p.data(), dt, f0, nft);
  shift_data(p.data(), dt, tshift, nft);
  for (int i = 0; i < nt; i++) out[i] = p[i];
}

void cal_rf_time(const double *thk, const double *vp, const double *vs, const double *rho,
                 const double *qa, const double *qb, int nlayer, int nt, double dt, double ray_p,
                 double f0, double time_shift, int rf_type, double *rcv_fun) {
  time_rf_core(thk, vp, vs, rho, qa, qb, nlayer, nt, dt, ray_p, f0, time_shift, rf_type, 0, false,
               false, rcv_fun, nullptr);
}
void cal_rf_freq(const double *thk, const double *vp, const double *vs, const double *rho,
                 const double *qa, const double *qb, int nlayer, int nt, double dt, double ray_p,
                 double f0, double t0, double water, int rf_type, double *rcv_fun) {
  freq_rf_core(thk, vp, vs, rho, qa, qb, nlayer, nt, dt, ray_p, f0, t0, water, rf_type, 0, false,
               rcv_fun, nullptr);
}
void cal_rf_par_freq(const double *thk, const double *vp, const double *vs, const double *rho,
                     const double *qa, const double *qb, int nlayer, int nt, double dt,
                     double ray_p, double f0, double t0, double water, int rf_type, int ipar,
                     double *rcv_fun, double *rcv_fun_p) {
  freq_rf_core(thk, vp, vs, rho, qa, qb, nlayer, nt, dt, ray_p, f0, t0, water, rf_type, ipar, true,
               rcv_fun, rcv_fun_p);
}
void cal_rf_par_freq_all(const double *thk, const double *vp, const double *vs, const double *rho,
                         const double *qa, const double *qb, int nlayer, int nt, double dt,
                         double ray_p, double f0, double t0, double water, int rf_type,
                         double *rcv_fun, double *rcv_fun_p) {
  freq_rf_core(thk, vp, vs, rho, qa, qb, nlayer, nt, dt, ray_p, f0, t0, water, rf_type, 0, true,
               rcv_fun, rcv_fun_p);
}
void cal_rf_par_time(const double *thk, const double *vp, const double *vs, const double *rho,
                     const double *qa, const double *qb, int nlayer, int nt, double dt,
                     double ray_p, double f0, double time_shift, int rf_type, int ipar,
                     double *rcv_fun, double *rcv_fun_p) {
  time_rf_core(thk, vp, vs, rho, qa, qb, nlayer, nt, dt, ray_p, f0, time_shift, rf_type, ipar,
               true, false, rcv_fun, rcv_fun_p);
}
void cal_rf_par_time_all(const double *thk, const double *vp, const double *vs, const double *rho,
                         const double *qa, const double *qb, int nlayer, int nt, double dt,
                         double ray_p, double f0, double time_shift, int rf_type, double *rcv_fun,
                         double *rcv_fun_p) {
  time_rf_core(thk, vp, vs, rho, qa, qb, nlayer, nt, dt, ray_p, f0, time_shift, rf_type, 0, true,
               true, rcv_fun, rcv_fun_p);
}

}  // namespace oracle
