#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "fr_roots.cuh"
int main() {
  std::mt19937_64 rng(1);
  std::uniform_real_distribution<double> U(0, 1);
  double worst = 0; long n = 0, nan_mismatch = 0, big = 0;
  auto check = [&](float p, float d1) {
    float a = psc::solve_cubic(p, d1), b = psc::solve_cubic_f64(p, d1);
    n++;
    if (std::isnan(b) || std::isnan(a)) { if (std::isnan(a) != std::isnan(b)) nan_mismatch++; return; }
    double scale = std::fabs((double)b) > 1e-30 ? std::fabs((double)b) : 1e-30;
    double e = std::fabs((double)a - b) / scale;
    // a root near 0 relative to the problem scale: compare against sqrt(|p|) too
    double s2 = std::sqrt(std::fabs((double)p)) + std::cbrt(std::fabs((double)d1 / 27));
    double e2 = std::fabs((double)a - b) / (s2 > 0 ? s2 : 1);
    double ee = e < e2 ? e : e2;
    if (ee > worst) { worst = ee; }
    if (ee > 2e-6) { big++; if (big < 10) printf("  p=%.9g d1=%.9g fast=%.9g f64=%.9g rel=%.3g\n", p, d1, a, b, ee); }
  };
  // wide log-uniform sweep over signs
  for (long i = 0; i < 3000000; i++) {
    float p = (float)((U(rng) < 0.5 ? -1 : 1) * std::pow(10.0, -8 + 9 * U(rng)));
    float d1 = (float)((U(rng) < 0.5 ? -1 : 1) * std::pow(10.0, -12 + 13 * U(rng)));
    check(p, d1);
  }
  printf("wide sweep: n=%ld worst rel diff %.3g, >2e-6: %ld, NaN mismatches %ld\n", n, worst, big, nan_mismatch);
  // f(R) regime: u ~ [0.01, 3], p = h2 b - avg(u^2) with h2 b << 1, d1 = 27 h2 q (q < 0) (+ small FAS rhs)
  worst = 0; n = 0; big = 0; nan_mismatch = 0;
  for (long i = 0; i < 3000000; i++) {
    double u = 0.01 + 3 * U(rng), h2 = std::pow(2.0, -2 * (2 + (int)(8 * U(rng))));
    float p = (float)(h2 * (2.0 + U(rng)) - u * u * (0.8 + 0.4 * U(rng)));
    float d1 = (float)(27.0 * h2 * (-3.0 * U(rng)) + 27.0 * 1e-4 * (U(rng) - 0.5));
    check(p, d1);
  }
  printf("f(R) regime: n=%ld worst rel diff %.3g, >2e-6: %ld, NaN mismatches %ld\n", n, worst, big, nan_mismatch);
  return 0;
}
