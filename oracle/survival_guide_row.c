/* survival_guide_row.c -- plain-C restatement of what ONE guide contributes to the survival MixtureNormal SVI step
 * (reference: bean/model/survival_model.py:215-424 model, :651-739 guide), in the shape of the per-guide loop a fused CUDA
 * step runs: everything local to the guide in closed form, the library-wide quantities (abundance sums, pathwise Dirichlet
 * derivative factors) passed in.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/): built by oracle/build_c.py with gcc, loaded with ctypes by
 * tests/test_survival_guide_row_c.py and compared guide by guide with oracle/survival_closed_form.py (numpy), which is
 * pinned to the autograd oracle and to the reference's own loss.  Never linked into crispr_bean_b200/.
 */
#include <math.h>
#include <stddef.h>

#define MAXB 8
static const double EPS = 1e-5;

/* digamma: upward recurrence to z >= 10, then the asymptotic series (as crispr_bean_b200/csrc/bean_math.cuh:digamma_f64) */
static double digamma(double z) {
  double sub = 0.0;
  while (z < 10.0) {
    sub += 1.0 / z;
    z += 1.0;
  }
  const double r = 1.0 / z, r2 = r * r;
  const double s = r2 * (1.0 / 12 + r2 * (-1.0 / 120 + r2 * (1.0 / 252 + r2 * (-1.0 / 240 + r2 * (1.0 / 132 + r2 * (-691.0 / 32760 + r2 * (1.0 / 12)))))));
  return log(z) - 0.5 * r - s - sub;
}

/* One Dirichlet-Multinomial row: get_alpha (model/utils.py:10-25) + DirichletMultinomial.log_prob; adds d ll / d e to de[]. */
static double dm_row(int B, const double* e, const double* sf, const double* smask, double a0, const double* x, double* de) {
  double p[MAXB], frac[MAXB], a[MAXB], g[MAXB];
  int live[MAXB];
  double S = 0.0, A = 0.0, N = 0.0;
  for (int b = 0; b < B; ++b) {
    p[b] = e[b] * sf[b];
    S += p[b];
    N += x[b];
  }
  for (int b = 0; b < B; ++b) {
    frac[b] = (p[b] + EPS / B) / (S + EPS);
    const double raw = frac[b] * a0 * smask[b];
    live[b] = raw >= EPS;
    a[b] = live[b] ? raw : EPS;
    A += a[b];
  }
  double ll = lgamma(A) + lgamma(1.0 + N) - lgamma(N + A);
  const double dA = digamma(A) - digamma(N + A);
  double dot = 0.0;
  for (int b = 0; b < B; ++b) {
    ll -= lgamma(1.0 + x[b]) + lgamma(a[b]) - lgamma(x[b] + a[b]);
    g[b] = live[b] ? (dA + digamma(x[b] + a[b]) - digamma(a[b])) * smask[b] : 0.0;
    dot += g[b] * frac[b];
  }
  for (int b = 0; b < B; ++b) de[b] += sf[b] * a0 / (S + EPS) * (g[b] - dot);
  return ll;
}

/* Guide-local part of the step.
 *   in : R replicates, B timepoints (tb), C control conditions (tc), L count layers;
 *        x [L][R][B] masked counts, a0 [L], sf [L][R][B] size factors, smask [R][B] sample mask, rg [R] replicate mask;
 *        counts [R][C][2] control allele counts; pa0 = pi_a0[g]; al[2] = alpha_pi[g]; mu[2] = (u, u + mu_variant);
 *        pi [R][2] the guide's draws; dgrad [R][2] = torch._dirichlet_grad(pi, cg, sum cg) (the pathwise factor)
 *   out: returns the guide's ELBO terms (likelihood, pi sites; NOT the variant / abundance / negctrl sites);
 *        d_log_alpha[2] = d ELBO / d log alpha_pi[g]; *d_mu_edited = d ELBO / d (growth rate of the edited allele)
 */
double survival_mixture_guide(int R, int B, int C, int L, const double* x, const double* a0, const double* sf, const double* smask,
                              const unsigned char* rg, const double* tb, const double* tc, const double* counts, double pa0,
                              const double* al, const double* mu, const double* pi, const double* dgrad, double mask_thres,
                              double prob_eps, double* d_log_alpha, double* d_mu_edited) {
  const double asum = al[0] + al[1];
  double cm[2], cg[2], d_cm[2] = {0, 0}, d_cg[2] = {0, 0};
  for (int a = 0; a < 2; ++a) {
    cm[a] = al[a] / asum * pa0;
    cg[a] = cm[a] > 1e-5 ? cm[a] : 1e-5;
  }
  const double norm_m = lgamma(cm[0] + cm[1]) - lgamma(cm[0]) - lgamma(cm[1]);
  const double norm_g = lgamma(cg[0] + cg[1]) - lgamma(cg[0]) - lgamma(cg[1]);
  const double psi_m = digamma(cm[0] + cm[1]), psi_g = digamma(cg[0] + cg[1]);
  double P[MAXB][2];
  for (int b = 0; b < B; ++b)
    for (int a = 0; a < 2; ++a) P[b][a] = exp(mu[a] * tb[b]);
  double elbo = 0.0, dmu1 = 0.0;
  for (int r = 0; r < R; ++r) {
    const double* p = pi + 2 * r;
    const double lp[2] = {log(p[0]), log(p[1])};
    double dpi[2] = {0, 0};
    /* guide `pi` site: unmasked (survival_model.py:699-712) */
    elbo -= norm_g + (cg[0] - 1) * lp[0] + (cg[1] - 1) * lp[1];
    for (int a = 0; a < 2; ++a) {
      d_cg[a] -= psi_g - digamma(cg[a]) + lp[a];
      dpi[a] -= (cg[a] - 1) / p[a];
    }
    if (rg[r]) {
      /* model `pi` prior and the Multinomial on pi exp(mu t_c), under repguide_mask (:313-346) */
      elbo += norm_m + (cm[0] - 1) * lp[0] + (cm[1] - 1) * lp[1];
      for (int a = 0; a < 2; ++a) {
        d_cm[a] += psi_m - digamma(cm[a]) + lp[a];
        dpi[a] += (cm[a] - 1) / p[a];
      }
      for (int c = 0; c < C; ++c) {
        const double* xc = counts + ((size_t)r * C + c) * 2;
        double w[2], q[2], n[2], h[2];
        for (int a = 0; a < 2; ++a) {
          w[a] = exp(mu[a] * tc[c]);
          q[a] = p[a] * w[a];
        }
        const double Sq = q[0] + q[1];
        double hbar = 0.0;
        for (int a = 0; a < 2; ++a) {
          n[a] = q[a] / Sq;
          const double cl = n[a] < prob_eps ? prob_eps : (n[a] > 1 - prob_eps ? 1 - prob_eps : n[a]);
          if (xc[a] != 0.0) elbo += xc[a] * log(cl);
          h[a] = (n[a] >= prob_eps && n[a] <= 1 - prob_eps) ? xc[a] / n[a] : 0.0;
          hbar += h[a] * n[a];
        }
        elbo += lgamma(xc[0] + xc[1] + 1) - lgamma(xc[0] + 1) - lgamma(xc[1] + 1);
        for (int a = 0; a < 2; ++a) {
          const double dq = (h[a] - hbar) / Sq;
          dpi[a] += dq * w[a];
          if (a == 1) dmu1 += dq * q[a] * tc[c];
        }
      }
    }
    /* count likelihood of every layer: e[b] = pi0 exp(mu0 t_b) + pi1 exp(mu1 t_b) */
    double e[MAXB], de[MAXB];
    for (int b = 0; b < B; ++b) {
      e[b] = p[0] * P[b][0] + p[1] * P[b][1];
      de[b] = 0.0;
    }
    for (int l = 0; l < L; ++l) {
      const double* xr = x + ((size_t)l * R + r) * B;
      double N = 0.0;
      for (int b = 0; b < B; ++b) N += xr[b];
      if (!(rg[r] && N > mask_thres)) continue; /* poutine.mask */
      elbo += dm_row(B, e, sf + ((size_t)l * R + r) * B, smask + (size_t)r * B, a0[l], xr, de);
    }
    for (int b = 0; b < B; ++b) {
      dpi[0] += de[b] * P[b][0];
      dpi[1] += de[b] * P[b][1];
      dmu1 += de[b] * p[1] * P[b][1] * tb[b];
    }
    /* pathwise derivative of the draw w.r.t. the guide concentration (torch _Dirichlet_backward) */
    const double gbar = p[0] * dpi[0] + p[1] * dpi[1];
    for (int a = 0; a < 2; ++a) d_cg[a] += dgrad[2 * r + a] * (dpi[a] - gbar);
  }
  /* concentrations -> log alpha_pi; clamp(min = 1e-5) passes the gradient where its input >= 1e-5 */
  double dC[2];
  for (int a = 0; a < 2; ++a) dC[a] = d_cm[a] + (cm[a] >= 1e-5 ? d_cg[a] : 0.0);
  const double k = pa0 / (asum * asum), dotC = dC[0] * al[0] + dC[1] * al[1];
  for (int a = 0; a < 2; ++a) d_log_alpha[a] = k * (dC[a] * asum - dotC) * al[a];
  *d_mu_edited = dmu1;
  return elbo;
}
