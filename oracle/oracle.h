/* ORACLE -- test infrastructure only.  CPU restatement (fp64, scalar, one problem at a time, OpenMP over
 * problems) of the reference's hot path, used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the checker and the timed CPU baseline.  The product library
 * (safe_mpc_b200/csrc) never includes, links or calls anything in this directory.
 *
 * PARITY PARTLY PINNED: the reference's numerics live in acados/HPIPM, CasADi, adam and
 * L4CasADi, none of which is vendored under /root/reference or installed here, and the reference's tests hold no golden
 * vectors.  Pinned on reference code executed in the build container (tests/golden/make_ref_*.py,
 * tests/test_ref_golden.py): the network (a5, the reference's NeuralNetwork class), the controller state machines with the
 * warm-start shift (a8, a9, the reference's controller classes driven with scripted solves), the closed loop (a12, the simulation
 * statements of the reference's scripts/mpc.py around those classes), the capsule distance (a4),
 * the plant step and the feasibility predicates (a10, a11), randomize_model (a13) and the configuration layer (the reference's
 * Parameters on its own config.yaml).  The acados / HPIPM / CasADi / adam numerics (a2, a3, a7) are PARITY UNPINNED: that part of the restatement is pinned only by independent cross-checks
 * (tests/): finite differences, a numpy re-implementation of the chain algorithms, an Euler-Lagrange (kinetic-energy) derivation of the
 * mass matrix and the Coriolis vector, direct verification of the KKT
 * conditions of every QP solution, and agreement of the QP solution with an independent dense solver (scipy SLSQP).
 *
 * The API mirrors include/safe_mpc_b200.h function by function (orc_* <-> smpc_*), host memory only; the
 * problem description is the very same struct (the boundary header is shared, the implementation is not).
 */
#ifndef ORC_ORACLE_H
#define ORC_ORACLE_H
#include <stdint.h>
#include "../include/safe_mpc_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NQ SMPC_NQ
#define ORC_NX SMPC_NX
#define ORC_NU SMPC_NU
#define ORC_NPAIR SMPC_NPAIR
#define ORC_MAX_POINTS SMPC_MAX_POINTS
#define ORC_HID SMPC_HID

typedef smpc_problem_t orc_problem_t;
typedef struct orc_handle orc_handle_t;
typedef struct orc_sim orc_sim_t;

/* threads <= 0: all OpenMP threads */
int orc_create(const orc_problem_t* prob, int32_t batch, int32_t threads, orc_handle_t** out);
void orc_destroy(orc_handle_t* h);
int orc_num_threads(const orc_handle_t* h);
/* 0: viability MLP accumulates in fp64 (default, deterministic); 1: plain fp32 like libtorch on CPU */
int orc_set_mlp_fp32(orc_handle_t* h, int32_t on);

/* tests only: hand the QP of every solve to an external solver (tests/emu: the engine's kernel sources compiled for the host), so that
 * whole closed loops of "kernel arithmetic" can be compared with the oracle's own solver without a GPU.  rec: the stage records of the
 * problem [N+1][SMPC_REC]; outputs as smpc_rti_solve produces them (x_temp, u_temp, acados status, IPM iterations, QP status, res[5]) */
typedef int (*orc_qp_hook_t)(const orc_problem_t* P, const double* rec, const double* x0, int32_t r, double* xt, double* ut, int32_t* status,
                             int32_t* qp_iter, int32_t* qp_status, double* qp_res);
int orc_set_qp_hook(orc_handle_t* h, orc_qp_hook_t fn);
/* tests only: margin probe.  eps > 0 switches it on: every accepted solve is run three iterations past its stop test (a robust solve stays
 * converged; at the boundary of feasibility the multipliers diverge and it does not), every failed solve is checked for an iterate that came
 * within qp_maxiter_accept x the tolerances.  orc_get_probe_flips returns, per problem, how many of its solves were such knife-edge solves:
 * their accept / fail status is decided by rounding, two correct implementations may disagree on it */
int orc_set_probe(orc_handle_t* h, double eps);
int orc_get_probe_flips(orc_handle_t* h, int32_t* out);

int orc_set_plant_inertial(orc_handle_t* h, const double* inertial);
int orc_set_torque_noise(orc_handle_t* h, const double* tau_noise);
/* cost.traj of the reference (controller.py:153-156, cost_definition.py:29-31,102-288): see smpc_set_ee_trajectory */
int orc_set_ee_trajectory(orc_handle_t* h, const double* traj, int32_t n);
int orc_set_guess(orc_handle_t* h, const double* xg, const double* ug);
int orc_get_guess(orc_handle_t* h, double* xg, double* ug);
int orc_get_temp(orc_handle_t* h, double* x_temp, double* u_temp);
int orc_reset_controller(orc_handle_t* h);
int orc_rti_solve(orc_handle_t* h, const double* x0, const uint8_t* active, int32_t* status);
int orc_controller_step(orc_handle_t* h, const double* x, const uint8_t* active, double* u, uint8_t* abort_flag);
/* the same with the solve replaced by a scripted outcome (tests only) */
int orc_controller_step_scripted(orc_handle_t* h, const double* x, const int32_t* status, const double* xt, const double* ut, double* u,
                                 uint8_t* abort_flag);
int orc_plant_step(orc_handle_t* h, const double* x, const double* u, double* x_next, double* a_applied);
int orc_tau(orc_handle_t* h, int32_t n, const double* x, const double* u, double* tau);
/* torque-input RK4 step with sensitivities (forward-mode AD); counterpart of smpc_rk4_sens */
int orc_rk4_sens(orc_handle_t* h, int32_t n, const double* x, const double* tau, double dt, double* x_next, double* A, double* B);
int orc_kinematics(orc_handle_t* h, int32_t n, const double* x, double* ee, double* dist);
int orc_nn_constraint(orc_handle_t* h, int32_t n, const double* x, double* c, double* grad);
int orc_get_lin(orc_handle_t* h, double* lin);
int orc_get_qp(orc_handle_t* h, double* dz, double* pi, double* lam, double* t);
int orc_get_state_i32(orc_handle_t* h, int32_t field, int32_t* out);
int orc_set_state_i32(orc_handle_t* h, int32_t field, const int32_t* in);
int orc_get_x_viable(orc_handle_t* h, double* x_viable);
int orc_get_qp_residuals(orc_handle_t* h, double* res5);

int orc_sim_create(orc_handle_t* main_ctrl, orc_handle_t* backup, int32_t n_steps, orc_sim_t** out);
void orc_sim_destroy(orc_sim_t* s);
int orc_sim_reset(orc_sim_t* s, const double* x_init);
int orc_sim_step(orc_sim_t* s);
/* tests only: replace the solves of the following steps by scripted outcomes */
int orc_sim_set_script(orc_sim_t* s, const int32_t* status, const double* xt, const double* ut, const int32_t* bk_status, const double* bk_xt,
                       const double* bk_ut);
int orc_sim_run(orc_sim_t* s, int32_t n_steps);
int orc_sim_get_outcome(orc_sim_t* s, int32_t* outcome);
int orc_sim_get_log(orc_sim_t* s, double* x, double* u);
int orc_sim_get_x_viable(orc_sim_t* s, double* xv);
int orc_sim_get_counters(orc_sim_t* s, int64_t* out4);

/* extras for pinning the oracle itself (tests only) */
/* mass matrix M(q) [5][5] and bias h(q,v) [5] of problem b; nominal != 0: controller model, else the perturbed plant */
int orc_mass_bias(orc_handle_t* h, int32_t b, int32_t nominal, const double* x, double* M, double* bias);
/* residual norms of the last QP of problem b: res[4], mu, and the iteration count */
int orc_qp_info(orc_handle_t* h, int32_t b, double* res4, double* mu, int32_t* iter, int32_t* status);

#ifdef __cplusplus
}
#endif
#endif
