// Host check of the rigid-body transfer arithmetic shared with the device kernels (csrc/multilevel_math.hpp).
#include <cmath>
#include <cstdio>
#include <random>

#include "multilevel_math.hpp"
using namespace mbfea::dev;

int main()
{
    std::mt19937_64 g(5);
    std::uniform_real_distribution<double> U(-2.0, 2.0);
    double worst = 0.0;
    for (int trial = 0; trial < 200; ++trial) {
        double da[3], db[3], A[36], c[6], r[6];
        for (double &x : da) x = U(g);
        for (double &x : db) x = U(g);
        for (double &x : A) x = U(g);
        for (double &x : c) x = U(g);
        for (double &x : r) x = U(g);
        // explicit B(d) = [[I, -[d]x], [0, I]]
        auto fill = [](const double *d, double B[6][6]) {
            for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) B[i][j] = i == j ? 1.0 : 0.0;
            B[0][4] = d[2]; B[0][5] = -d[1]; B[1][3] = -d[2]; B[1][5] = d[0]; B[2][3] = d[1]; B[2][4] = -d[0];
        };
        double Ba[6][6], Bb[6][6], out[6];
        fill(da, Ba); fill(db, Bb);
        for (int j = 0; j < 6; ++j) for (int i = 0; i < 6; ++i) worst = std::fmax(worst, std::fabs(rigid_entry(da, j, i) - Ba[j][i]));
        rigid_apply(da, c, out);
        for (int i = 0; i < 6; ++i) { double s = 0; for (int j = 0; j < 6; ++j) s += Ba[i][j] * c[j]; worst = std::fmax(worst, std::fabs(s - out[i])); }
        rigid_apply_t(da, r, out);
        for (int i = 0; i < 6; ++i) { double s = 0; for (int j = 0; j < 6; ++j) s += Ba[j][i] * r[j]; worst = std::fmax(worst, std::fabs(s - out[i])); }
        for (int i = 0; i < 6; ++i) for (int m = 0; m < 6; ++m) {
            double s = 0;
            for (int j = 0; j < 6; ++j) for (int l = 0; l < 6; ++l) s += Ba[j][i] * A[6 * j + l] * Bb[l][m];
            worst = std::fmax(worst, std::fabs(s - galerkin_entry(da, db, A, i, m)));
        }
    }
    std::printf("max deviation %.3e\n", worst);
    return worst < 1e-12 ? 0 : 1;
}
