/* CPU oracle, FP64, plain C + OpenMP. TEST INFRASTRUCTURE ONLY (see galaxify_oracle.py for the rules).
 *
 * Restates the formulas of the reference (bikuta6/nbody-deep-sim, src/galaxify/simulation.py) in double precision
 * from the FP32-rounded inputs, for problem sizes the reference's own (N,N,3) tensors cannot reach:
 *   oracle_accel_f64    : simulation.py:80-88  a_i = G * sum_{j != i} m_j (r_j - r_i) (|r_j - r_i|^2 + eps2)^-1.5
 *   oracle_energies_f64 : simulation.py:100-113 k = sum 0.5 m v^2 ; u = sum_{i<j} -G m_i m_j / (|r_ij| + eps)
 * Parity pin: checked against tests/golden/*.npz (outputs of the unmodified reference) by tests/test_oracle.py,
 * to the 1e-6 the reference's FP32 arithmetic allows.
 */
#include <math.h>
#include <stddef.h>

void oracle_accel_f64(const float* pos, const float* mass, int n, int lo, int hi, double g, double eps2, double* out) {
#pragma omp parallel for schedule(static)
    for (int i = lo; i < hi; ++i) {
        const double xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int j = 0; j < n; ++j) {
            if (j == i) continue; /* fill_diagonal_(0), simulation.py:85 */
            const double dx = pos[3 * j] - xi, dy = pos[3 * j + 1] - yi, dz = pos[3 * j + 2] - zi;
            const double d2 = dx * dx + dy * dy + dz * dz + eps2;
            const double w = mass[j] / (d2 * sqrt(d2));
            ax += w * dx;
            ay += w * dy;
            az += w * dz;
        }
        out[3 * (size_t)(i - lo)] = g * ax;
        out[3 * (size_t)(i - lo) + 1] = g * ay;
        out[3 * (size_t)(i - lo) + 2] = g * az;
    }
}

/* Same sum, plus its condition number kappa_i = sum_j |term_ij| / |a_i| (1-norm of the terms over the norm of the
 * result): the factor by which per-term rounding errors are amplified for body i. Used to state a principled
 * tolerance for bodies whose forces nearly cancel (a galaxy's centre), where NO FP32 summation reaches 1e-5. */
void oracle_accel_cond_f64(const float* pos, const float* mass, int n, int lo, int hi, double g, double eps2, double* out,
                           double* kappa) {
#pragma omp parallel for schedule(static)
    for (int i = lo; i < hi; ++i) {
        const double xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
        double ax = 0.0, ay = 0.0, az = 0.0, sabs = 0.0;
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            const double dx = pos[3 * j] - xi, dy = pos[3 * j + 1] - yi, dz = pos[3 * j + 2] - zi;
            const double d2 = dx * dx + dy * dy + dz * dz + eps2;
            const double w = mass[j] / (d2 * sqrt(d2));
            ax += w * dx;
            ay += w * dy;
            az += w * dz;
            sabs += w * sqrt(dx * dx + dy * dy + dz * dz);
        }
        out[3 * (size_t)(i - lo)] = g * ax;
        out[3 * (size_t)(i - lo) + 1] = g * ay;
        out[3 * (size_t)(i - lo) + 2] = g * az;
        const double norm = sqrt(ax * ax + ay * ay + az * az);
        kappa[i - lo] = norm > 0.0 ? sabs / norm : INFINITY;
    }
}

void oracle_energies_f64(const float* pos, const float* vel, const float* mass, int n, double g, double eps,
                         double* out_uk) {
    double u = 0.0, k = 0.0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : u, k)
    for (int i = 0; i < n; ++i) {
        const double xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
        double phi = 0.0;
        for (int j = i + 1; j < n; ++j) { /* upper triangle, simulation.py:113 */
            const double dx = pos[3 * j] - xi, dy = pos[3 * j + 1] - yi, dz = pos[3 * j + 2] - zi;
            phi += mass[j] / (sqrt(dx * dx + dy * dy + dz * dz) + eps);
        }
        u -= g * mass[i] * phi;
        k += 0.5 * mass[i] * ((double)vel[3 * i] * vel[3 * i] + (double)vel[3 * i + 1] * vel[3 * i + 1] +
                              (double)vel[3 * i + 2] * vel[3 * i + 2]);
    }
    out_uk[0] = u;
    out_uk[1] = k;
}
