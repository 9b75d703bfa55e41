// Host-side LSH Ensemble parameter selection, evaluated ONCE per (query size, threshold) and handed
// to the kernels as small integers.
//
// Replaces, from github.com/ekzhu/lshensemble v1.1.0 (reference call sites src/lshe/lshe.go:157,165):
//   LshForest.OptimalKL + probFalsePositive/probFalseNegative (probability.go: midpoint-rule integration,
//   precision 0.01) and Containment(q, x, qSize, xSize) = (xSize/qSize + 1) * J / (1 + J), J = eq/S.
// The f64 expressions are kept in the published form and order so that thresholds fall on the same
// side of every comparison as the Go code.
#include <cmath>

#include "../flat_index.h"

namespace groot {
namespace {
const double kPrecision = 0.01;

inline double collision(double s, double xq, int k, int l) {
    return 1.0 - std::pow(1.0 - std::pow(s / (1.0 + xq - s), static_cast<double>(k)), static_cast<double>(l));
}
double integrate_fp(double xq, int k, int l, double a, double b) {
    double area = 0.0;
    for (double x = a; x < b; x += kPrecision) area += collision(x + 0.5 * kPrecision, xq, k, l) * kPrecision;
    return area;
}
double integrate_fn(double xq, int k, int l, double a, double b) {
    double area = 0.0;
    for (double x = a; x < b; x += kPrecision) area += (1.0 - collision(x + 0.5 * kPrecision, xq, k, l)) * kPrecision;
    return area;
}
}  // namespace

void optimal_kl(int max_k, int max_l, int x, int q, double t, int* K, int* L) {
    const double xq = static_cast<double>(x) / static_cast<double>(q);
    double best = 1.7976931348623157e308;
    *K = 0; *L = 0;
    for (int l = 1; l <= max_l; l++) {
        for (int k = 1; k <= max_k; k++) {
            double fp = (xq >= 1.0 || xq >= t) ? integrate_fp(xq, k, l, 0.0, t) : integrate_fp(xq, k, l, 0.0, xq);
            double fn = xq >= 1.0 ? integrate_fn(xq, k, l, t, 1.0) : (xq >= t ? integrate_fn(xq, k, l, t, xq) : 0.0);
            double err = fn + fp;
            if (best > err) { best = err; *K = k; *L = l; }
        }
    }
}

int eq_min_for(int S, int q_size, int x_size, double threshold) {
    if (q_size == 0 || x_size == 0) return S + 1;
    for (int eq = 1; eq <= S; eq++) {
        double j = static_cast<double>(eq) / static_cast<double>(S);
        double c = (static_cast<double>(x_size) / static_cast<double>(q_size) + 1.0) * j / (1.0 + j);
        if (c > threshold) return eq;
    }
    return S + 1;
}

}  // namespace groot
