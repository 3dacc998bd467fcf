/*
 * raceline_oracle.h -- TEST INFRASTRUCTURE (see raceline_oracle.c).
 * CPU restatement of the reference hot path; shares only the POD structs of the
 * public ABI header so that oracle and CUDA outputs can be compared field by field.
 */
#ifndef RACELINE_ORACLE_H
#define RACELINE_ORACLE_H

#include "../include/raceline_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

void orc_default_params(rl_params* p);

double orc_ray_to_ring(double px, double py, double dx, double dy, const double* seg, int m);
double orc_min_dist_to_segments(double px, double py, const double* seg, int m);
void orc_corridor(const double* pxy, const double* nxy, int n,
                  const double* inner_seg, int m_in, const double* outer_seg, int m_out,
                  double guard, double* lo, double* hi);
void orc_distances_to_rings(double px, double py, double nx, double ny,
                            const double* inner_seg, int m_in, const double* outer_seg, int m_out,
                            double* d_inner, double* d_outer);

void orc_normals(const double* pxy, int n, int closed, double* nxy);
void orc_heading_curv(const double* pxy, int n, double h, int closed, double* heading, double* kappa);
void orc_lin_geom(const double* pxy, const double* nxy, int n, double h, int closed,
                  double* A1, double* A2, double* N0, double* W);

/* gamma2 == NULL: eval_cost_grad_frozen; else eval_cost_grad_timeweighted. work: 8*n doubles. */
double orc_eval_cost_grad(const double* A1, const double* A2, const double* N0, const double* W,
                          const double* gamma2, double h, double lambda_smooth, const double* alpha,
                          int n, int closed, double* grad, double* work);

void orc_ax_max_at(const rl_params* C, double vi, double ki, double* a_acc, double* a_brk);
double orc_velocity_profile(const rl_params* C, const double* kappa, int n, double h, int closed,
                            double* v, double* ax);
void orc_time_weights(const rl_params* C, const double* kappa, const double* v, int n, double* gamma2);

/* stage = RL_STAGE_MINCURV | RL_STAGE_MINTIME; out_v/out_ax may be NULL for MINCURV */
int orc_solve(int stage, const double* center_xy, int n,
              const double* inner_seg, int m_in, const double* outer_seg, int m_out,
              double L, int closed, const rl_params* C,
              double* out_xy, double* out_heading, double* out_curv,
              double* out_alpha_total, double* out_alpha_last,
              double* out_v, double* out_ax, rl_job_stats* st);

/* geom_oracle.c: the centre-line + width/geometry stage (main.cpp:404-474, 1270-1335) */
int orc_geom_rows(int samples, int closed, int emit_dup);
int orc_centerline_geom(const double* mids_xy, int n_mid, int samples, int closed, int emit_dup,
                        const double* inner_seg, int m_in, const double* outer_seg, int m_out, const rl_params* C,
                        double* out_xy, double* s_rel, double* heading, double* curvature,
                        double* d_inner, double* d_outer, double* width, double* v_kappa, double* L_out, double* s0_out);

#ifdef __cplusplus
}
#endif
#endif
