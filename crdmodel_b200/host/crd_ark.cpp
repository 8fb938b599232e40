// crd_ark.cpp — explicit adaptive Runge-Kutta driver behind the ARKode-legacy interface.
//
// What the reference uses (src/FHNmodel_torus.cpp:356-373,423,491): ARKodeCreate, ARKodeInit(mem, f,
// NULL, T0, y) -> fully explicit; ARKodeSStolerances(1e-5, 1e-10); ARKodeSetUserData;
// ARKodeSetMaxNumSteps(200000); ARKode(mem, tout, y, &t, ARK_NORMAL); ARKodeFree.  SUNDIALS itself is
// an un-vendored third-party dependency of the reference (API brackets it to 2.6.0-2.7.0, ARKode
// 1.0.x-1.1.0) and its source is not available here, so this file RESTATES the published ARKode 1.x
// explicit algorithm from its documentation (parity with SUNDIALS itself is UNPINNED):
//   * default 4th-order explicit table Zonneveld 5-3-4 (5 stages, embedding order 3),
//   * error weights ewt = 1/(rtol |y| + atol), local error dsm = WRMS(err, ewt), accept if dsm <= 1,
//   * PID step controller k = (0.58, 0.21, 0.1), error bias 1.5, safety 0.96, growth 20 (1e4 on the
//     first step), eta_min 0.1, no change for eta in [1, 1.5], after an error-test failure no growth,
//     and eta <= 0.3 from the second failure on, at most 7 failures per step,
//   * initial step from the CVODE-style ||y''|| iteration (arkHin),
//   * ARK_NORMAL: step past tout and return the cubic Hermite interpolant at tout,
//   * max-steps limit per ARKode call, "too much accuracy" test tolsf = uround * WRMS(y, ewt) > 1.
// Everything is done through the N_Vector ops table and the RHS callback, so the same code drives
// the CPU checker (oracle N_Vector + the reference's own f) and the device path (crd_b200.h).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "crd_ark.h"
#include "crd_pow.h"

namespace {

constexpr int S_MAX = 8;
constexpr double UROUND = DBL_EPSILON;
constexpr double TINY = 1.0e-10;
constexpr double ONEPSM = 1.000001;
constexpr double ONEMSM = 0.999999;

struct ArkMem {
  ARKRhsFn fe = nullptr;
  void *user_data = nullptr;
  double rtol = 0, atol = 0;
  bool tol_set = false, inited = false, setup_done = false;
  long mxstep = 500;  // MXSTEP_DEFAULT
  // method (Zonneveld 5-3-4)
  int s = 5, q = 4, p = 3;
  double A[S_MAX][S_MAX] = {}, b[S_MAX] = {}, b2[S_MAX] = {}, c[S_MAX] = {};
  // controller constants
  double k1 = 0.58, k2 = 0.21, k3 = 0.1, bias = 1.5, safety = 0.96, growth = 20.0, etamx1 = 10000.0,
         etamxf = 0.3, etamin = 0.1, lbound = 1.0, ubound = 1.5;
  int small_nef = 2, maxnef = 7, maxncf = 10;
  // state
  double tn = 0, h = 0, next_h = 0, hold = 0, eta = 1, etamax = 10000.0, hin = 0, hfixed = 0;
  double ehist[2] = {1.0, 1.0};
  long nst = 0, nst_attempts = 0, nfe = 0, netf = 0;
  double ynorm_sq_next = -1;  // sum (y w)^2 of the accepted state when the fused finish provided it
  // vectors
  N_Vector yn = nullptr, yold = nullptr, ycur = nullptr, fnew = nullptr, fold = nullptr, ewt = nullptr,
           tempv = nullptr, sdata = nullptr;
  N_Vector F[S_MAX] = {};
  N_Vector Fp[S_MAX] = {};  // stage derivatives of the current attempt (Fp[0] may alias fnew)
  long nglobal = 0;
  const crd_fused_ops *fused = nullptr;
  bool reuse_first = false;
  bool resident = true;     // use fused->erk_evolve when it is offered and applies
  bool resident_na = false; // it answered "does not apply" once: stop asking
  bool no_stage_finish = false;  // rhs_lincomb_finish answered "does not apply" once
  bool no_stage_pair = false;    // rhs_pair answered "does not apply" once (or was switched off)
  bool stage2_ready = false;     // F[1] holds stage 2 of the NEXT step, evaluated together with f(tn, yn) for the step size stage2_h
  double stage2_h = 0.0;
};

void set_zonneveld(ArkMem *m) {
  m->s = 5; m->q = 4; m->p = 3;
  std::memset(m->A, 0, sizeof m->A);
  m->c[0] = 0.0; m->c[1] = 0.5; m->c[2] = 0.5; m->c[3] = 1.0; m->c[4] = 0.75;
  m->A[1][0] = 0.5;
  m->A[2][1] = 0.5;
  m->A[3][2] = 1.0;
  m->A[4][0] = 5.0 / 32.0; m->A[4][1] = 7.0 / 32.0; m->A[4][2] = 13.0 / 32.0; m->A[4][3] = -1.0 / 32.0;
  m->b[0] = 1.0 / 6.0; m->b[1] = 1.0 / 3.0; m->b[2] = 1.0 / 3.0; m->b[3] = 1.0 / 6.0; m->b[4] = 0.0;
  m->b2[0] = -0.5; m->b2[1] = 7.0 / 3.0; m->b2[2] = 7.0 / 3.0; m->b2[3] = 13.0 / 6.0; m->b2[4] = -16.0 / 3.0;
}

void free_vectors(ArkMem *m) {
  N_Vector *vs[] = {&m->yn, &m->yold, &m->ycur, &m->fnew, &m->fold, &m->ewt, &m->tempv, &m->sdata};
  for (N_Vector *v : vs) { if (*v) N_VDestroy(*v); *v = nullptr; }
  for (int i = 0; i < S_MAX; ++i) { if (m->F[i]) N_VDestroy(m->F[i]); m->F[i] = nullptr; }
}

long global_length(N_Vector v) {
  long lrw = 0, liw = 0;
  v->ops->nvspace(v, &lrw, &liw);
  return lrw;
}

// ewt = 1/(rtol |y| + atol)   (arkEwtSetSS: abs, scale, addconst, min-check, inv)
int ewt_set(ArkMem *m, N_Vector y) {
  N_VAbs(y, m->tempv);
  N_VScale(m->rtol, m->tempv, m->tempv);
  N_VAddConst(m->tempv, m->atol, m->tempv);
  if (N_VMin(m->tempv) <= 0.0) return -1;
  N_VInv(m->tempv, m->ewt);
  return 0;
}

int rhs(ArkMem *m, double t, N_Vector y, N_Vector out) {
  int r = m->fe(t, y, out, m->user_data);
  m->nfe++;
  return r;
}

// z = y0 + sum_j coef[j] X[j]
int assemble(ArkMem *m, N_Vector y0, int n, const double *coef, N_Vector *X, N_Vector z) {
  if (m->fused && m->fused->lincomb && n + 1 <= CRD_ARK_MAX_LINCOMB) {
    double cc[CRD_ARK_MAX_LINCOMB];
    N_Vector XX[CRD_ARK_MAX_LINCOMB];
    cc[0] = 1.0; XX[0] = y0;
    for (int j = 0; j < n; ++j) { cc[j + 1] = coef[j]; XX[j + 1] = X[j]; }
    return m->fused->lincomb(n + 1, cc, XX, z);
  }
  // SUNDIALS 2.x arkSet: sdata = sum_j h A_ij F_j built by successive N_VLinearSum, then y = yn + sdata
  N_VConst(0.0, m->sdata);
  for (int j = 0; j < n; ++j)
    if (coef[j] != 0.0) N_VLinearSum(coef[j], X[j], 1.0, m->sdata, m->sdata);
  N_VLinearSum(1.0, y0, 1.0, m->sdata, z);
  return 0;
}

double upper_bound_h0(ArkMem *m, double tdist) {
  // arkUpperBoundH0: hub = 0.1 tdist, unless f is large: hub = 1 / max_i( |f_i| / (0.1 |y_i| + 1/ewt_i) )
  N_Vector t1 = m->tempv, t2 = m->sdata;
  N_VAbs(m->yn, t2);
  N_VInv(m->ewt, t1);                       // rtol|y| + atol
  N_VLinearSum(0.1, t2, 1.0, t1, t1);
  N_VAbs(m->fnew, t2);
  N_VDiv(t2, t1, t1);
  double hub_inv = N_VMaxNorm(t1);
  double hub = 0.1 * tdist;
  if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
  return hub;
}

// ||y''|| estimate by one Euler probe of size hg (arkYddNorm)
int ydd_norm(ArkMem *m, double hg, double *yddnrm) {
  N_VLinearSum(hg, m->fnew, 1.0, m->yn, m->ycur);
  int r = rhs(m, m->tn + hg, m->ycur, m->F[0]);
  if (r < 0) return ARK_RHSFUNC_FAIL;
  if (r > 0) return 1;
  N_VLinearSum(1.0, m->F[0], -1.0, m->fnew, m->F[0]);
  N_VScale(1.0 / hg, m->F[0], m->F[0]);
  *yddnrm = N_VWrmsNorm(m->F[0], m->ewt);
  return 0;
}

int initial_step(ArkMem *m, double tout) {
  const int MAX_ITERS = 4;
  double tdiff = tout - m->tn;
  if (tdiff == 0.0) return ARK_TOO_CLOSE;
  double sign = tdiff > 0 ? 1.0 : -1.0;
  double tdist = std::fabs(tdiff);
  double tround = UROUND * std::max(std::fabs(m->tn), std::fabs(tout));
  if (tdist < 2.0 * tround) return ARK_TOO_CLOSE;
  double hlb = 100.0 * tround;
  double hub = upper_bound_h0(m, tdist);
  double hg = std::sqrt(hlb * hub);
  if (hub < hlb) { m->h = sign * hg; return 0; }
  double hnew = hg;
  bool ok = false;
  for (int count1 = 1; count1 <= MAX_ITERS; ++count1) {
    double yddnrm = 0;
    bool got = false;
    for (int count2 = 1; count2 <= MAX_ITERS; ++count2) {
      int r = ydd_norm(m, hg * sign, &yddnrm);
      if (r < 0) return r;
      if (r == 0) { got = true; break; }
      hg *= 0.2;
    }
    if (!got) return ARK_REPTD_RHSFUNC_ERR;
    hnew = (yddnrm * hub * hub > 2.0) ? std::sqrt(2.0 / yddnrm) : std::sqrt(hg * hub);
    if (count1 == MAX_ITERS) break;
    double hrat = hnew / hg;
    if (hrat > 0.5 && hrat < 2.0) { ok = true; }
    if (count1 > 1 && hrat > 2.0) { hnew = hg; ok = true; }
    if (ok) break;
    hg = hnew;
  }
  double h0 = 0.5 * hnew;
  if (h0 < hlb) h0 = hlb;
  if (h0 > hub) h0 = hub;
  m->h = sign * h0;
  return 0;
}

// PID controller (arkAdaptPID) + the bounds of arkAdapt; ecur = bias*dsm of the attempt just made
double adapt_eta(ArkMem *m, double dsm) {
  double hcur = m->h;
  double k = (double)m->p;  // embedding order (pq = 0)
  double e1 = std::max(m->bias * dsm, TINY);
  double e2 = std::max(m->ehist[0], TINY);
  double e3 = std::max(m->ehist[1], TINY);
  // crd_pow_pos: x^y as one fixed sequence of IEEE operations, so that the device-resident loop (which runs this controller
  // inside its kernel) chooses the same step sizes bit for bit (libm's and CUDA's pow differ in the last place)
  double h_acc = hcur * crd_pow_pos(e1, -m->k1 / k) * crd_pow_pos(e2, m->k2 / k) * crd_pow_pos(e3, -m->k3 / k);
  double int_dir = hcur / std::fabs(hcur);
  h_acc *= m->safety;
  h_acc = int_dir * std::min(std::fabs(h_acc), std::fabs(m->etamax * hcur));
  h_acc = int_dir * std::max(std::fabs(h_acc), std::fabs(m->etamin * hcur));
  if (std::fabs(h_acc) > std::fabs(hcur * m->lbound * ONEMSM) && std::fabs(h_acc) < std::fabs(hcur * m->ubound * ONEPSM))
    h_acc = hcur;
  return h_acc / hcur;
}

int compute_solution(ArkMem *m, double *dsm) {
  double hb[S_MAX], hd[S_MAX];
  for (int j = 0; j < m->s; ++j) { hb[j] = m->h * m->b[j]; hd[j] = m->h * (m->b[j] - m->b2[j]); }
  if (m->fused && m->fused->erk_finish) {
    double out[2] = {0, 0};
    int r = m->fused->erk_finish(m->s, hb, hd, m->yn, m->Fp, m->ycur, m->rtol, m->atol, out);
    if (r != 0) return ARK_MEM_FAIL;
    *dsm = std::sqrt(out[0] / (double)m->nglobal);
    m->ynorm_sq_next = out[1];
    return 0;
  }
  // arkComputeSolutions: y = yn + sum h b_j F_j ; err = sum h (b_j - b2_j) F_j ; dsm = WRMS(err, ewt)
  N_VScale(1.0, m->yn, m->ycur);
  N_VConst(0.0, m->tempv);
  for (int j = 0; j < m->s; ++j) {
    if (hb[j] != 0.0) N_VLinearSum(hb[j], m->Fp[j], 1.0, m->ycur, m->ycur);
    N_VLinearSum(hd[j], m->Fp[j], 1.0, m->tempv, m->tempv);
  }
  *dsm = N_VWrmsNorm(m->tempv, m->ewt);
  m->ynorm_sq_next = -1;
  return 0;
}

int take_step(ArkMem *m) {
  int nef = 0, ncf = 0;
  for (;;) {
    m->nst_attempts++;
    bool retry_rhs = false;
    for (int is = 0; is < m->s; ++is) m->Fp[is] = m->F[is];
    bool finished = false;   // the last stage's evaluation also produced ynew and the error norm
    double dsm_fused = 0;
    for (int is = 0; is < m->s; ++is) {
      if (is == 0 && m->reuse_first) {
        m->Fp[0] = m->fnew;  // f(tn, yn), evaluated when the previous step completed
        continue;
      }
      if (is == 1 && m->stage2_ready) {
        // evaluated in the same pass as f(tn, yn) when the previous step completed — for exactly this step size?
        m->stage2_ready = false;
        if (m->reuse_first && m->h == m->stage2_h) continue;
      }
      int r;
      if (is == m->s - 1 && is > 0 && !m->no_stage_finish && m->fused && m->fused->rhs_lincomb_finish && m->fused->erk_finish &&
          m->hfixed == 0.0) {
        // last stage + finish in one pass, when every earlier stage enters the last one's state: X = (yn, F_0 .. F_{s-2})
        bool dense = true;
        for (int j = 0; j < is; ++j) dense = dense && m->A[is][j] != 0.0;
        if (dense) {
          double cc[S_MAX + 1], hb[S_MAX], hd[S_MAX], out[2] = {0, 0};
          N_Vector XX[S_MAX + 1];
          cc[0] = 1.0; XX[0] = m->yn;
          for (int j = 0; j < is; ++j) { cc[j + 1] = m->h * m->A[is][j]; XX[j + 1] = m->Fp[j]; }
          for (int j = 0; j < m->s; ++j) { hb[j] = m->h * m->b[j]; hd[j] = m->h * (m->b[j] - m->b2[j]); }
          int fr = m->fused->rhs_lincomb_finish(m->tn + m->c[is] * m->h, m->s, cc, hb, hd, XX, m->ycur, m->rtol, m->atol, out, m->user_data);
          if (fr < 0) return ARK_RHSFUNC_FAIL;
          if (fr == 0) {
            m->nfe++;
            dsm_fused = std::sqrt(out[0] / (double)m->nglobal);
            m->ynorm_sq_next = out[1];
            finished = true;
            continue;
          }
          m->no_stage_finish = true;   // does not apply to this problem: stop asking
        }
      }
      if (is > 0 && m->fused && m->fused->rhs_lincomb) {
        // stage state yn + h sum_j A_ij F_j formed inside the evaluation (zero coefficients dropped)
        double cc[S_MAX + 1];
        N_Vector XX[S_MAX + 1];
        int n = 0;
        cc[n] = 1.0; XX[n++] = m->yn;
        for (int j = 0; j < is; ++j)
          if (m->A[is][j] != 0.0) { cc[n] = m->h * m->A[is][j]; XX[n++] = m->Fp[j]; }
        r = m->fused->rhs_lincomb(m->tn + m->c[is] * m->h, n, cc, XX, m->Fp[is], m->user_data);
        m->nfe++;
      } else {
        N_Vector ystage = m->yn;
        if (is > 0) {
          double coef[S_MAX];
          N_Vector Xs[S_MAX];
          int n = 0;
          for (int j = 0; j < is; ++j)
            if (m->A[is][j] != 0.0) { coef[n] = m->h * m->A[is][j]; Xs[n++] = m->Fp[j]; }
          if (assemble(m, m->yn, n, coef, Xs, m->ycur) != 0) return ARK_MEM_FAIL;
          ystage = m->ycur;
        }
        r = rhs(m, m->tn + m->c[is] * m->h, ystage, m->Fp[is]);
      }
      if (r < 0) return ARK_RHSFUNC_FAIL;
      if (r > 0) {  // recoverable: shrink and retry (etacf = 0.25)
        if (++ncf == m->maxncf || m->hfixed != 0.0) return ARK_REPTD_RHSFUNC_ERR;
        m->etamax = 1.0;
        m->h *= 0.25;
        retry_rhs = true;
        break;
      }
    }
    if (retry_rhs) continue;
    double dsm = dsm_fused;
    if (!finished) {
      int r = compute_solution(m, &dsm);
      if (r != 0) return r;
    }
    if (m->hfixed != 0.0) { m->eta = 1.0; m->ehist[1] = m->ehist[0]; m->ehist[0] = dsm * m->bias; return 0; }
    m->eta = adapt_eta(m, dsm);
    if (dsm <= 1.0) {
      m->ehist[1] = m->ehist[0];
      m->ehist[0] = dsm * m->bias;
      return 0;
    }
    // error test failed
    nef++;
    m->netf++;
    m->etamax = 1.0;
    if (nef == m->maxnef) return ARK_ERR_FAILURE;
    m->eta = std::min(adapt_eta(m, dsm), 1.0);
    if (nef >= m->small_nef) m->eta = std::min(m->eta, m->etamxf);
    m->h *= m->eta;
    if (std::fabs(m->h) <= 0.0 || m->tn + m->h == m->tn) return ARK_ERR_FAILURE;
  }
}

// cubic Hermite interpolant on [tn - hold, tn]
int dense_eval(ArkMem *m, double t, N_Vector yout) {
  if (m->nst == 0) { N_VScale(1.0, m->yn, yout); return 0; }
  double tau = (t - m->tn) / m->hold;  // in [-1, 0]
  double tfuzz = 100.0 * UROUND * (std::fabs(m->tn) + std::fabs(m->hold));
  if ((t - (m->tn - m->hold)) * m->hold < -tfuzz || (t - m->tn) * m->hold > tfuzz) return ARK_BAD_T;
  if (tau == 0.0) { N_VScale(1.0, m->yn, yout); return 0; }
  double s = 1.0 + tau;
  double h00 = (2.0 * s - 3.0) * s * s + 1.0, h01 = (3.0 - 2.0 * s) * s * s;
  double h10 = ((s - 2.0) * s + 1.0) * s * m->hold, h11 = (s - 1.0) * s * s * m->hold;
  if (m->fused && m->fused->lincomb) {
    double cc[4] = {h00, h01, h10, h11};
    N_Vector XX[4] = {m->yold, m->yn, m->fold, m->fnew};
    return m->fused->lincomb(4, cc, XX, yout) == 0 ? 0 : ARK_MEM_FAIL;
  }
  N_VLinearSum(h00, m->yold, h01, m->yn, yout);
  N_VLinearSum(h10, m->fold, 1.0, yout, yout);
  N_VLinearSum(h11, m->fnew, 1.0, yout, yout);
  return 0;
}

// Hand the loop to the vector implementation (crd_fused_ops.erk_evolve).  Returns EVOLVE_NA when it does not apply,
// otherwise what the host loop would return before its dense-output evaluation.
constexpr int EVOLVE_NA = 12345;
int evolve_resident(ArkMem *m, double tout, int itask) {
  if (m->s > CRD_ERK_MAX_STAGES) return EVOLVE_NA;
  // the "too much accuracy" test of the first step needs the norm the fused finish supplies from then on
  if (m->ynorm_sq_next < 0) {
    double nrm = N_VWrmsNorm(m->yn, m->ewt);
    m->ynorm_sq_next = nrm * nrm * (double)m->nglobal;
  }
  crd_erk_state st;
  std::memset(&st, 0, sizeof st);
  st.s = m->s; st.p = m->p;
  for (int i = 0; i < m->s; ++i) {
    for (int j = 0; j < m->s; ++j) st.A[i][j] = m->A[i][j];
    st.b[i] = m->b[i]; st.d[i] = m->b[i] - m->b2[i]; st.c[i] = m->c[i];
  }
  st.rtol = m->rtol; st.atol = m->atol;
  st.k1 = m->k1; st.k2 = m->k2; st.k3 = m->k3; st.bias = m->bias; st.safety = m->safety; st.growth = m->growth;
  st.etamxf = m->etamxf; st.etamin = m->etamin; st.lbound = m->lbound; st.ubound = m->ubound;
  st.small_nef = m->small_nef; st.maxnef = m->maxnef;
  st.nglobal = m->nglobal;
  st.tout = tout; st.itask = itask; st.max_steps = m->mxstep;
  st.tn = m->tn; st.next_h = m->next_h; st.hold = m->hold; st.eta = m->eta; st.etamax = m->etamax;
  st.ehist[0] = m->ehist[0]; st.ehist[1] = m->ehist[1];
  st.ynorm_sq = m->ynorm_sq_next;
  st.nst = m->nst; st.nst_attempts = m->nst_attempts; st.nfe = m->nfe; st.netf = m->netf;
  st.yn = m->yn; st.yold = m->yold; st.ycur = m->ycur; st.fnew = m->fnew; st.fold = m->fold;
  for (int i = 0; i < m->s; ++i) st.F[i] = m->F[i];
  int r = m->fused->erk_evolve(&st, m->user_data);
  if (r > 0) return EVOLVE_NA;
  if (r < 0) return ARK_MEM_FAIL;
  m->tn = st.tn; m->next_h = st.next_h; m->hold = st.hold; m->eta = st.eta; m->etamax = st.etamax;
  m->ehist[0] = st.ehist[0]; m->ehist[1] = st.ehist[1];
  m->ynorm_sq_next = st.ynorm_sq;
  m->nst = st.nst; m->nst_attempts = st.nst_attempts; m->nfe = st.nfe; m->netf = st.netf;
  m->yn = st.yn; m->yold = st.yold; m->ycur = st.ycur; m->fnew = st.fnew; m->fold = st.fold;
  m->h = (st.flag < 0 && st.h_failed != 0.0) ? st.h_failed : st.hold;
  return st.flag;
}

}  // namespace

extern "C" {

void *ARKodeCreate(void) {
  ArkMem *m = new (std::nothrow) ArkMem;
  if (m) set_zonneveld(m);
  return m;
}

int ARKodeInit(void *mem, ARKRhsFn fe, ARKRhsFn fi, realtype t0, N_Vector y0) {
  if (!mem) return ARK_MEM_NULL;
  ArkMem *m = (ArkMem *)mem;
  if (!y0 || !fe) return ARK_ILL_INPUT;
  if (fi != nullptr) {
    std::fprintf(stderr, "crd_ark: implicit / IMEX right-hand sides are not supported (explicit only)\n");
    return ARK_ILL_INPUT;
  }
  const struct _generic_N_Vector_Ops *o = y0->ops;
  if (!o->nvclone || !o->nvdestroy || !o->nvlinearsum || !o->nvconst || !o->nvdiv || !o->nvscale || !o->nvabs ||
      !o->nvinv || !o->nvaddconst || !o->nvmaxnorm || !o->nvwrmsnorm || !o->nvmin || !o->nvspace)
    return ARK_ILL_INPUT;
  free_vectors(m);
  m->fe = fe;
  m->tn = t0;
  N_Vector *vs[] = {&m->yn, &m->yold, &m->ycur, &m->fnew, &m->fold, &m->ewt, &m->tempv, &m->sdata};
  for (N_Vector *v : vs) { *v = N_VClone(y0); if (!*v) { free_vectors(m); return ARK_MEM_FAIL; } }
  for (int i = 0; i < m->s; ++i) { m->F[i] = N_VClone(y0); if (!m->F[i]) { free_vectors(m); return ARK_MEM_FAIL; } }
  N_VScale(1.0, y0, m->yn);
  m->nglobal = global_length(y0);
  m->nst = m->nst_attempts = m->nfe = m->netf = 0;
  m->ehist[0] = m->ehist[1] = 1.0;
  m->etamax = m->etamx1;
  m->h = m->next_h = m->hold = 0;
  m->inited = true;
  m->setup_done = false;
  return ARK_SUCCESS;
}

int ARKodeSStolerances(void *mem, realtype reltol, realtype abstol) {
  if (!mem) return ARK_MEM_NULL;
  ArkMem *m = (ArkMem *)mem;
  if (!m->inited) return ARK_NO_MALLOC;
  if (reltol < 0 || abstol < 0) return ARK_ILL_INPUT;
  m->rtol = reltol; m->atol = abstol; m->tol_set = true;
  return ARK_SUCCESS;
}

int ARKodeSetUserData(void *mem, void *user_data) {
  if (!mem) return ARK_MEM_NULL;
  ((ArkMem *)mem)->user_data = user_data;
  return ARK_SUCCESS;
}

int ARKodeSetMaxNumSteps(void *mem, long int mxsteps) {
  if (!mem) return ARK_MEM_NULL;
  ((ArkMem *)mem)->mxstep = mxsteps == 0 ? 500 : mxsteps;  // 0 -> default, <0 -> no limit
  return ARK_SUCCESS;
}

int crd_ARKodeSetFusedOps(void *mem, const crd_fused_ops *ops) {
  if (!mem) return ARK_MEM_NULL;
  ((ArkMem *)mem)->fused = ops;
  return ARK_SUCCESS;
}
int crd_ARKodeSetReuseFirstStage(void *mem, int on) {
  if (!mem) return ARK_MEM_NULL;
  ((ArkMem *)mem)->reuse_first = on != 0;
  return ARK_SUCCESS;
}
int crd_ARKodeSetResident(void *mem, int on) {
  if (!mem) return ARK_MEM_NULL;
  ((ArkMem *)mem)->resident = on != 0;
  ((ArkMem *)mem)->resident_na = false;
  return ARK_SUCCESS;
}
int crd_ARKodeSetStageFinish(void *mem, int on) {
  if (!mem) return ARK_MEM_NULL;
  ((ArkMem *)mem)->no_stage_finish = on == 0;
  return ARK_SUCCESS;
}
int crd_ARKodeSetStagePair(void *mem, int on) {
  if (!mem) return ARK_MEM_NULL;
  ((ArkMem *)mem)->no_stage_pair = on == 0;
  ((ArkMem *)mem)->stage2_ready = false;
  return ARK_SUCCESS;
}
int crd_ARKodeSetInitStep(void *mem, realtype hin) {
  if (!mem) return ARK_MEM_NULL;
  ((ArkMem *)mem)->hin = hin;
  return ARK_SUCCESS;
}
int crd_ARKodeSetFixedStep(void *mem, realtype hfixed) {
  if (!mem) return ARK_MEM_NULL;
  ((ArkMem *)mem)->hfixed = hfixed;
  return ARK_SUCCESS;
}

int crd_ARKodeGetButcherTable(void *mem, int *s, int *q, int *p, realtype *A, realtype *c, realtype *b, realtype *b2) {
  static_assert(S_MAX == CRD_ARK_TABLE_DIM, "table dimension");
  if (!mem) return ARK_MEM_NULL;
  if (!s || !q || !p || !A || !c || !b || !b2) return ARK_ILL_INPUT;
  const ArkMem *m = (const ArkMem *)mem;
  *s = m->s; *q = m->q; *p = m->p;
  for (int i = 0; i < S_MAX; ++i) {
    for (int j = 0; j < S_MAX; ++j) A[i * S_MAX + j] = m->A[i][j];
    c[i] = m->c[i]; b[i] = m->b[i]; b2[i] = m->b2[i];
  }
  return ARK_SUCCESS;
}

int ARKode(void *mem, realtype tout, N_Vector yout, realtype *tret, int itask) {
  if (!mem) return ARK_MEM_NULL;
  ArkMem *m = (ArkMem *)mem;
  if (!m->inited) return ARK_NO_MALLOC;
  if (!yout || !tret || (itask != ARK_NORMAL && itask != ARK_ONE_STEP)) return ARK_ILL_INPUT;
  const bool use_ewt_vec = !(m->fused && m->fused->erk_finish);

  if (!m->setup_done) {
    if (!m->tol_set) { std::fprintf(stderr, "crd_ark: no integration tolerances set\n"); return ARK_ILL_INPUT; }
    if (ewt_set(m, m->yn) != 0) return ARK_ILL_INPUT;
    int r = rhs(m, m->tn, m->yn, m->fnew);
    if (r != 0) return ARK_FIRST_RHSFUNC_ERR;
    if (m->hfixed != 0.0) m->h = m->hfixed;
    else if (m->hin != 0.0) m->h = m->hin;
    else { int hr = initial_step(m, tout); if (hr != 0) return hr; }
    if ((tout - m->tn) * m->h < 0) return ARK_ILL_INPUT;
    m->next_h = m->h;
    m->ynorm_sq_next = -1;
    m->setup_done = true;
  } else if (itask == ARK_NORMAL && (m->tn - tout) * m->h >= 0.0) {
    // already past tout: interpolate, no new step
    int r = dense_eval(m, tout, yout);
    if (r != 0) return r;
    *tret = tout;
    return ARK_SUCCESS;
  }

  // device-resident step loop: the vector implementation runs the whole loop below by itself
  if (m->resident && !m->resident_na && m->fused && m->fused->erk_evolve && m->hfixed == 0.0) {
    int r = evolve_resident(m, tout, itask);
    if (r == EVOLVE_NA) m->resident_na = true;
    else {
      if (r != ARK_SUCCESS) {
        if (r == ARK_TOO_MUCH_WORK) std::fprintf(stderr, "crd_ark: at t = %g, mxstep steps taken before reaching tout\n", m->tn);
        else if (r == ARK_TOO_MUCH_ACC) std::fprintf(stderr, "crd_ark: at t = %g, too much accuracy requested\n", m->tn);
        else std::fprintf(stderr, "crd_ark: at t = %g and h = %g, step failed with flag %d\n", m->tn, m->h, r);
        N_VScale(1.0, m->yn, yout); *tret = m->tn;
        return r;
      }
      if (itask == ARK_NORMAL) {
        int dr = dense_eval(m, tout, yout);
        if (dr != 0) return dr;
        *tret = tout;
      } else {
        N_VScale(1.0, m->yn, yout);
        *tret = m->tn;
      }
      return ARK_SUCCESS;
    }
  }

  long nstloc = 0;
  for (;;) {
    m->h = (m->hfixed != 0.0) ? m->hfixed : m->next_h;
    double nrm;
    if (m->nst > 0 && use_ewt_vec) { if (ewt_set(m, m->yn) != 0) { N_VScale(1.0, m->yn, yout); *tret = m->tn; return ARK_ILL_INPUT; } }
    if (m->mxstep > 0 && nstloc >= m->mxstep) {
      std::fprintf(stderr, "crd_ark: at t = %g, mxstep steps taken before reaching tout\n", m->tn);
      N_VScale(1.0, m->yn, yout); *tret = m->tn;
      return ARK_TOO_MUCH_WORK;
    }
    if (!use_ewt_vec && m->ynorm_sq_next >= 0) nrm = std::sqrt(m->ynorm_sq_next / (double)m->nglobal);
    else nrm = N_VWrmsNorm(m->yn, m->ewt);
    if (UROUND * nrm > 1.0) {
      std::fprintf(stderr, "crd_ark: at t = %g, too much accuracy requested\n", m->tn);
      N_VScale(1.0, m->yn, yout); *tret = m->tn;
      return ARK_TOO_MUCH_ACC;
    }
    int kflag = take_step(m);
    if (kflag != 0) {
      std::fprintf(stderr, "crd_ark: at t = %g and h = %g, step failed with flag %d\n", m->tn, m->h, kflag);
      N_VScale(1.0, m->yn, yout); *tret = m->tn;
      return kflag;
    }
    // complete the step: (told, yold, fold) <- (tn, yn, fnew); yn <- ycur; fnew <- f(tn, yn)
    std::swap(m->yold, m->yn);
    std::swap(m->yn, m->ycur);
    std::swap(m->fold, m->fnew);
    m->hold = m->h;
    m->tn += m->h;
    m->nst++;
    nstloc++;
    m->etamax = m->growth;
    m->next_h = m->h * m->eta;
    int r = 1;
    m->stage2_ready = false;
    if (!m->no_stage_pair && m->reuse_first && m->hfixed == 0.0 && m->fused && m->fused->rhs_pair && m->s >= 2 && m->A[1][0] != 0.0) {
      // f(tn, yn) and the next step's second stage f(tn + c2 h, yn + h a21 f(tn, yn)) in one pass over yn
      r = m->fused->rhs_pair(m->tn, m->tn + m->c[1] * m->next_h, m->next_h * m->A[1][0], m->yn, m->fnew, m->F[1], m->user_data);
      if (r < 0) { N_VScale(1.0, m->yn, yout); *tret = m->tn; return ARK_RHSFUNC_FAIL; }
      if (r == 0) { m->nfe += 2; m->stage2_ready = true; m->stage2_h = m->next_h; }
      else m->no_stage_pair = true;   // does not apply to this problem: stop asking
    }
    if (r != 0) {
      r = rhs(m, m->tn, m->yn, m->fnew);
      if (r != 0) { N_VScale(1.0, m->yn, yout); *tret = m->tn; return ARK_RHSFUNC_FAIL; }
    }

    if (itask == ARK_NORMAL && (m->tn - tout) * m->h >= 0.0) {
      int dr = dense_eval(m, tout, yout);
      if (dr != 0) return dr;
      *tret = tout;
      return ARK_SUCCESS;
    }
    if (itask == ARK_ONE_STEP) {
      N_VScale(1.0, m->yn, yout);
      *tret = m->tn;
      return ARK_SUCCESS;
    }
  }
}

void ARKodeFree(void **mem) {
  if (!mem || !*mem) return;
  ArkMem *m = (ArkMem *)*mem;
  // the reference never queries the counters; CRD_ARK_STATS=1 prints them when the integrator is freed
  if (const char *e = std::getenv("CRD_ARK_STATS"))
    if (e[0] == '1')
      std::fprintf(stderr, "crd_ark: nst = %ld, attempts = %ld, nfe = %ld, netf = %ld, t = %.17g\n", m->nst, m->nst_attempts, m->nfe,
                   m->netf, m->tn);
  free_vectors(m);
  delete m;
  *mem = nullptr;
}

int ARKodeGetNumSteps(void *mem, long int *n) { if (!mem) return ARK_MEM_NULL; *n = ((ArkMem *)mem)->nst; return 0; }
int ARKodeGetNumStepAttempts(void *mem, long int *n) { if (!mem) return ARK_MEM_NULL; *n = ((ArkMem *)mem)->nst_attempts; return 0; }
int ARKodeGetNumRhsEvals(void *mem, long int *nfe, long int *nfi) {
  if (!mem) return ARK_MEM_NULL;
  *nfe = ((ArkMem *)mem)->nfe;
  if (nfi) *nfi = 0;
  return 0;
}
int ARKodeGetNumErrTestFails(void *mem, long int *n) { if (!mem) return ARK_MEM_NULL; *n = ((ArkMem *)mem)->netf; return 0; }
int ARKodeGetCurrentStep(void *mem, realtype *h) { if (!mem) return ARK_MEM_NULL; *h = ((ArkMem *)mem)->next_h; return 0; }
int ARKodeGetLastStep(void *mem, realtype *h) { if (!mem) return ARK_MEM_NULL; *h = ((ArkMem *)mem)->hold; return 0; }
int ARKodeGetCurrentTime(void *mem, realtype *t) { if (!mem) return ARK_MEM_NULL; *t = ((ArkMem *)mem)->tn; return 0; }

}  // extern "C"
