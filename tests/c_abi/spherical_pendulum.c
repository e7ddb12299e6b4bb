/* Plain-C user of the C ABI (include/bionc_cuda.h): no Python, no torch.
 * A single segment hanging from the ground by a spherical joint (one LIN3 block: rp pinned to the origin),
 * two systems, 100 RK4 steps through bionc_cuda_rk4_host, then one forward-dynamics call.  The numbers are
 * printed with 17 significant digits; tests/test_gpu_c_abi_example.py compares them with the CPU oracle.
 *
 * build: gcc -std=c99 -I include tests/c_abi/spherical_pendulum.c -L bionc_b200/lib -lbionc_cuda -Wl,-rpath,$PWD/bionc_b200/lib */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "bionc_cuda.h"

int main(void) {
    /* segment: L = 1, alpha = beta = gamma = pi/2; m = 2, natural COM (0, -0.5, 0), natural pseudo-inertia diag(0.1, 0.6, 0.1) */
    const double half_pi = 1.5707963267948966;
    double phic[4] = {1.0 * cos(half_pi), cos(half_pi), 1.0, 1.0 * cos(half_pi)};
    double mass[1] = {2.0}, com[3] = {0.0, -0.5, 0.0};
    double J[9] = {0.1, 0, 0, 0, 0.6, 0, 0, 0, 0.1};
    int32_t rigid_row[1] = {3}; /* holonomic order: the joint's 3 rows first, then the 6 rigid-body rows */
    int32_t blk_type[1] = {BIONC_BLK_LIN3}, blk_parent[1] = {-1}, blk_child[1] = {0}, row_h[1] = {0}, row_j[1] = {0};
    double param[BIONC_BLK_NPARAM];
    memset(param, 0, sizeof(param));
    param[4 + 1] = 1.0; /* child point = rp: coefficients (0, 1, 0, 0) over (u, rp, rd, w); ground point = origin */

    BioncModelDesc d;
    memset(&d, 0, sizeof(d));
    d.nb_segments = 1, d.nb_blocks = 1;
    d.seg_phi_const = phic, d.seg_mass = mass, d.seg_com = com, d.seg_J = J, d.seg_rigid_row = rigid_row;
    d.blk_type = blk_type, d.blk_parent = blk_parent, d.blk_child = blk_child, d.blk_row_h = row_h, d.blk_row_j = row_j;
    d.blk_param = param;

    BioncModel* model = NULL;
    if (bionc_cuda_model_create(&d, 0, &model) != 0) {
        fprintf(stderr, "model_create: %s\n", bionc_cuda_last_error());
        return 1;
    }
    int32_t N, mj, m;
    bionc_cuda_model_counts(model, &N, &mj, &m);
    printf("counts %d %d %d\n", N, mj, m);

    /* two systems: u, rp, rd, w; the second one tilted by 0.3 rad about x and swinging */
    double Q[2][12] = {{1, 0, 0, 0, 0, 0, 0, -1, 0, 0, 0, 1}, {0}};
    double Qd[2][12] = {{0}, {0}};
    const double a = 0.3, om = 0.7;
    double v[3] = {0, cos(a), sin(a)};             /* v = rp - rd, rotated (0, 1, 0) */
    double w3[3] = {0, -sin(a), cos(a)};           /* rotated (0, 0, 1) */
    Q[1][0] = 1;                                    /* u = x */
    Q[1][6] = -v[0], Q[1][7] = -v[1], Q[1][8] = -v[2];
    Q[1][9] = w3[0], Q[1][10] = w3[1], Q[1][11] = w3[2];
    /* angular velocity om about x: d/dt of a vector p is om * (x cross p) = om * (0, -pz, py) */
    Qd[1][6] = 0, Qd[1][7] = -om * Q[1][8], Qd[1][8] = om * Q[1][7];
    Qd[1][9] = 0, Qd[1][10] = -om * Q[1][11], Qd[1][11] = om * Q[1][10];

    for (int s = 0; s < 2; ++s) {
        printf("Q0 %d", s);
        for (int i = 0; i < 12; ++i) printf(" %.17g", Q[s][i]);
        printf("\nQd0 %d", s);
        for (int i = 0; i < 12; ++i) printf(" %.17g", Qd[s][i]);
        printf("\n");
    }

    double Qdd[2][12], lam[2][9];
    int32_t status[2] = {0, 0};
    if (bionc_cuda_forward_dynamics_host(model, &Q[0][0], &Qd[0][0], NULL, &Qdd[0][0], &lam[0][0], status, 2) != 0) {
        fprintf(stderr, "forward_dynamics_host: %s\n", bionc_cuda_last_error());
        return 1;
    }
    for (int s = 0; s < 2; ++s) {
        printf("Qdd %d", s);
        for (int i = 0; i < 12; ++i) printf(" %.17g", Qdd[s][i]);
        printf("\nlam %d", s);
        for (int i = 0; i < 9; ++i) printf(" %.17g", lam[s][i]);
        printf("\nstatus %d %d\n", s, status[s]);
    }
    if (bionc_cuda_rk4_host(model, &Q[0][0], &Qd[0][0], 0.01, NULL, 100, 1, NULL, status, 2) != 0) {
        fprintf(stderr, "rk4_host: %s\n", bionc_cuda_last_error());
        return 1;
    }
    for (int s = 0; s < 2; ++s) {
        printf("Qf %d", s);
        for (int i = 0; i < 12; ++i) printf(" %.17g", Q[s][i]);
        printf("\nQdf %d", s);
        for (int i = 0; i < 12; ++i) printf(" %.17g", Qd[s][i]);
        printf("\nstatus %d %d\n", s, status[s]);
    }
    printf("launches %lld\n", (long long)bionc_cuda_launch_count());
    bionc_cuda_model_destroy(model);
    return 0;
}
