/* le_force.c -- CPU restatement of the Lagrangian force evaluation and position updates that sit either
 * side of the spread/interpolate path.  TEST INFRASTRUCTURE ONLY: used by tests/ as the
 * checker of libibk.so's device versions (ibk_compute_lagrangian_force, ibk_markers_lincomb); the product
 * never calls it.
 *
 * Follows, loop for loop (serial, same accumulation order):
 *   springs        src/IB/IBStandardForceGen.cpp:813-930   (T = force_fcn(R, params), default
 *                  include/ibamr/IBSpringForceFunctions.h:99-103: kappa * (R - rest); skipped when R < eps)
 *   beams          src/IB/IBStandardForceGen.cpp:1037-1147 (F = K (X_next + X_prev - 2 X_mastr - D2X0))
 *   target points  src/IB/IBStandardForceGen.cpp:1201-1299 (F += kappa (X0 - X) - eta U)
 *   order of the three groups and the zeroing of F: IBStandardForceGen.cpp:253-303, IBMethod.cpp:834-858
 * Index arrays hold node numbers (not pre-multiplied by NDIM as the PETSc-index arrays of the reference are).
 * Parity with the reference is pinned only through the reference's own sample structure files and the
 * analytic checks in tests/test_force_oracle.py (no golden force output exists in the reference's tests).
 */
#include <float.h>
#include <math.h>

void le_oracle_spring_force(int ndim, int num_springs, const int* mastr, const int* slave, const double* kappa,
                            const double* rest, const double* X_node, double* F_node)
{
    for (int k = 0; k < num_springs; ++k)
    {
        const int m = ndim * mastr[k], s = ndim * slave[k];
        double D[3] = { 0.0, 0.0, 0.0 }, R2 = 0.0;
        for (int d = 0; d < ndim; ++d)
        {
            D[d] = X_node[s + d] - X_node[m + d];
            R2 += D[d] * D[d];
        }
        const double R = sqrt(R2);
        if (R < DBL_EPSILON) continue;
        const double T_over_R = (kappa[k] * (R - rest[k])) / R;
        for (int d = 0; d < ndim; ++d)
        {
            const double F = T_over_R * D[d];
            F_node[m + d] += F;
            F_node[s + d] -= F;
        }
    }
}

void le_oracle_beam_force(int ndim, int num_beams, const int* mastr, const int* next, const int* prev,
                          const double* rigidity, const double* curvature /* [num_beams][ndim] */, const double* X_node,
                          double* F_node)
{
    for (int k = 0; k < num_beams; ++k)
    {
        const int m = ndim * mastr[k], n = ndim * next[k], p = ndim * prev[k];
        const double K = rigidity[k];
        for (int d = 0; d < ndim; ++d)
        {
            const double F = K * (X_node[n + d] + X_node[p + d] - 2.0 * X_node[m + d] - curvature[ndim * k + d]);
            F_node[m + d] += 2.0 * F;
            F_node[n + d] -= F;
            F_node[p + d] -= F;
        }
    }
}

void le_oracle_target_force(int ndim, int num_targets, const int* idx, const double* kappa, const double* eta,
                            const double* X0 /* [num_targets][ndim] */, const double* X_node, const double* U_node,
                            double* F_node)
{
    for (int k = 0; k < num_targets; ++k)
    {
        const int i = ndim * idx[k];
        for (int d = 0; d < ndim; ++d) F_node[i + d] += kappa[k] * (X0[ndim * k + d] - X_node[i + d]) - eta[k] * U_node[i + d];
    }
}
