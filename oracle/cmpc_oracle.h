/*
 * cmpc_oracle.h -- CPU oracle of the centroidal-MPC condensed-QP path.
 * TEST INFRASTRUCTURE ONLY (see cmpc_oracle.c header). PARITY UNPINNED by the reference.
 * Shares cmpc_config / status codes with include/cmpc.h so both sides take the same struct.
 */
#ifndef CMPC_ORACLE_H_
#define CMPC_ORACLE_H_
#include "../include/cmpc.h"
#ifdef __cplusplus
extern "C" {
#endif
/* H [p*p] row-major, g [p], p = 3*L*N, pinned rows/cols = identity. */
int cmpc_oracle_build(const cmpc_config* c, const double* state, const double* des_state,
                      const double* des_inputs, double* H, double* g, int32_t* status);
/* One instance. forces [L][N][3]; lam [2][N][L][5] (may be NULL); active [N][L] (may be NULL). */
int cmpc_oracle_solve(const cmpc_config* c, const double* state, const double* des_state,
                      const double* des_inputs, double* forces, int32_t* status, int32_t* iters,
                      double* kkt, double* lam, uint16_t* active);
/* B instances, one per thread round-robin over nthreads pthreads. */
int cmpc_oracle_solve_batch(const cmpc_config* c, int B, const double* state, const double* des_state,
                            const double* des_inputs, double* forces, int32_t* status, int32_t* iters,
                            double* kkt, double* lam, uint16_t* active, int nthreads);
/* The CPU BASELINE port (cmpc_cpu_fast.c): same algorithm, compact closed-form build, no allocation per solve,
 * persistent thread pool (created on the first call, re-created when nthreads changes).  Checked against
 * cmpc_oracle_solve_batch in tests/test_oracle.py; timed by bench.py's cpu_baseline / --impl reference legs. */
int cmpc_fast_solve_batch(const cmpc_config* c, int B, const double* state, const double* des_state,
                          const double* des_inputs, double* forces, int32_t* status, int32_t* iters,
                          double* kkt, uint16_t* active, int nthreads);
/* Reference plant step (CentroidalMPC.cpp:85-92). forces [L][3], contact [L]. */
void cmpc_oracle_plant_step(const cmpc_config* c, const double* x, const double* feet,
                            const double* contact, const double* forces, double* xn);
#ifdef __cplusplus
}
#endif
#endif
