/* ORACLE -- test infrastructure only.  CPU restatement (fp64, scalar, one problem at a time, OpenMP over
 * problems) of the reference's hot path, used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the checker and the timed CPU baseline.  The product library
 * (safe_mpc_b200/csrc) never includes, links or calls anything in this directory.
 *
 * PARITY UNPINNED: the reference's numerics live in acados/HPIPM, CasADi, adam and L4CasADi, none of which
 * is vendored under /root/reference or installed here, and the reference's tests hold no golden vectors
 * (SURVEY.md section 4, 8c).  This restatement is pinned only by independent cross-checks (tests/):
 * finite differences, a numpy re-implementation of the chain algorithms, torch fp64 autograd for the MLP,
 * and direct verification of the KKT conditions of every QP solution.
 *
 * The API mirrors include/safe_mpc_b200.h function by function (orc_* <-> smpc_*), host memory only.
 * struct orc_problem has the same layout as struct smpc_problem so the Python host layer fills one ctypes
 * structure for both (safe_mpc_b200/abi.py).
 */
#ifndef ORC_ORACLE_H
#define ORC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NQ 5
#define ORC_NX 10
#define ORC_NU 5
#define ORC_NPAIR 6
#define ORC_MAX_POINTS 8
#define ORC_HID 256
#define ORC_LIN_FIELDS 162

typedef struct orc_problem {
  int32_t nq, N, n_pairs, n_points, controller, nn_rows, nn_terminal_soft, stage0_collision_rows,
      cost_type, abort_flag, qp_iter_max, reserved_i[5];
  double dt, q_weight, r_weight, lm, alpha, eps, slack_penalty_e;
  double tol_x, tol_tau, tol_obs, tol_safe, tol_conv;
  double qp_mu0, qp_tol_stat, qp_tol_eq, qp_tol_ineq, qp_tol_comp, qp_alpha_min, qp_reg_prim;
  double gravity[3];
  double reserved_d[8];
  double joint_R[ORC_NQ][9], joint_p[ORC_NQ][3], joint_axis[ORC_NQ][3], inertial[ORC_NQ][10];
  double x_min[ORC_NX], x_max[ORC_NX], lbx[ORC_NX], ubx[ORC_NX], lbx_e[ORC_NX], ubx_e[ORC_NX];
  double tau_min[ORC_NU], tau_max[ORC_NU];
  double ee_ref[3];
  int32_t point_body[ORC_MAX_POINTS];
  double point_local[ORC_MAX_POINTS][3];
  int32_t pair_pa[ORC_NPAIR], pair_pb[ORC_NPAIR];
  double pair_C[ORC_NPAIR][3], pair_D[ORC_NPAIR][3];
  double pair_lo_ocp[ORC_NPAIR], pair_lo_chk[ORC_NPAIR], pair_hi;
  double nn_mean[ORC_NQ], nn_std[ORC_NQ];
  const float* nn_weights;
} orc_problem_t;

typedef struct orc_handle orc_handle_t;

int orc_create(const orc_problem_t* prob, int32_t batch, int32_t threads, orc_handle_t** out);
void orc_destroy(orc_handle_t* h);
int orc_num_threads(const orc_handle_t* h);

int orc_set_plant_inertial(orc_handle_t* h, const double* inertial);
int orc_set_torque_noise(orc_handle_t* h, const double* tau_noise);
int orc_set_guess(orc_handle_t* h, const double* xg, const double* ug);
int orc_get_guess(orc_handle_t* h, double* xg, double* ug);
int orc_get_temp(orc_handle_t* h, double* x_temp, double* u_temp);
int orc_reset_controller(orc_handle_t* h);
int orc_rti_solve(orc_handle_t* h, const double* x0, int32_t* status);
int orc_controller_step(orc_handle_t* h, const double* x, const uint8_t* active, double* u, uint8_t* abort_flag);
int orc_plant_step(orc_handle_t* h, const double* x, const double* u, double* x_next, double* a_applied);
int orc_tau(orc_handle_t* h, int32_t n, const double* x, const double* u, double* tau);
int orc_kinematics(orc_handle_t* h, int32_t n, const double* x, double* ee, double* dist);
int orc_nn_constraint(orc_handle_t* h, int32_t n, const double* x, double* c, double* grad);
int orc_linearize(orc_handle_t* h, double* lin);
int orc_get_state_i32(orc_handle_t* h, int32_t field, int32_t* out);
int orc_set_state_i32(orc_handle_t* h, int32_t field, const int32_t* in);
int orc_get_x_viable(orc_handle_t* h, double* x_viable);

/* extras for pinning the oracle itself (tests only) */
/* QP of the last orc_rti_solve of problem `b` in dense form + its primal/dual solution:
 * see oracle/oracle.py:dump_qp for the buffer layout */
int orc_dump_qp(orc_handle_t* h, int32_t b, double* buf, int64_t buf_len, int64_t* used);
/* mass matrix M(q) and bias h(q,v) of the plant model of problem b (env_model.py:200-201) */
int orc_mass_bias(orc_handle_t* h, int32_t b, int32_t nominal, const double* x, double* M, double* bias);

#ifdef __cplusplus
}
#endif
#endif
