/* ref_dubins.c -- TEST ORACLE (see oracle.h).  Restatement of
 * Edge::NewEdge (src/dubinsedge.cpp:12-21), DubinsEdge::CalculateTrajectory
 * (:172-830, space_has_time_ == false branch :807-827) and the LineCheck
 * polyline (src/obstacle.cpp:1092-1133) of cosa0481/DRRT.
 *
 * Eigen is only a container there: `diff/D` is an element-wise division,
 * `(diff*diff).sum()` is d0*d0 + d1*d1, `c + r*phis.cos()` is c + r*cos(phi)
 * with libm cos.  Candidate order, strict `bestDist > len` updates, the R,L,R
 * arc lengths of the lrl candidate, the "0s0" override (:445-449)
 * and the repeated `val -= 0.1` arc sampling (:488-502 and siblings) are kept.
 */
#include "oracle.h"

#include <math.h>
#include <string.h>

static double len2(double dx, double dy) { return sqrt(dx * dx + dy * dy); }

/* arc sampling: dubinsedge.cpp:484-505 (right, dir=-1) / :522-543 (left, dir=+1).
 * phis = Zero(ceil(|phi_end-phi_start|/0.1)+1); writes past the end of phis are
 * undefined behaviour in the reference (no Eigen bounds check in release
 * builds): the oracle drops them. Trailing entries the loop does not reach
 * stay 0, exactly as the zero-initialised Eigen vector would. */
static int sample_arc(double cx, double cy, double r, double phi_start,
                      double phi_end, int dir, double *x, double *y, int at) {
  const double delta_phi = 0.1;
  if (dir < 0) {
    if (phi_end > phi_start) phi_end -= 2.0 * REF_PI;
  } else {
    if (phi_end < phi_start) phi_end += 2.0 * REF_PI;
  }
  int n = (int)(ceil(fabs(phi_end - phi_start) / delta_phi) + 1);
  double phis[REF_TRAJ_MAX];
  if (n > REF_TRAJ_MAX - at) n = REF_TRAJ_MAX - at;
  for (int i = 0; i < n; i++) phis[i] = 0.0;
  if (phi_end == phi_start) {
    phis[0] = phi_start;
  } else {
    double val = phi_start;
    int i = 0;
    if (dir < 0) {
      while (val >= phi_end) {
        if (i < n) phis[i] = val;
        val -= delta_phi;
        i += 1;
      }
    } else {
      while (val <= phi_end) {
        if (i < n) phis[i] = val;
        val += delta_phi;
        i += 1;
      }
    }
    if (i < n) phis[i] = phi_end;
  }
  for (int i = 0; i < n; i++) {
    x[at + i] = cx + r * cos(phis[i]);
    y[at + i] = cy + r * sin(phis[i]);
  }
  return n;
}

void ref_dubins_trajectory(const double *start, const double *end, double r_min,
                           int metric, ref_traj *out) {
  const double il[2] = {start[0], start[1]};
  const double gl[2] = {end[0], end[1]};
  const double initial_theta = start[2], goal_theta = end[2];
  const double edge_dist = ref_metric(metric, start, end); /* NewEdge :19 */

  /* :183-202 */
  double irc[2], ilc[2], grc[2], glc[2];
  irc[0] = il[0] + r_min * cos(initial_theta - (REF_PI / 2.0));
  irc[1] = il[1] + r_min * sin(initial_theta - (REF_PI / 2.0));
  ilc[0] = il[0] + r_min * cos(initial_theta + (REF_PI / 2.0));
  ilc[1] = il[1] + r_min * sin(initial_theta + (REF_PI / 2.0));
  grc[0] = gl[0] + r_min * cos(goal_theta - (REF_PI / 2.0));
  grc[1] = gl[1] + r_min * sin(goal_theta - (REF_PI / 2.0));
  glc[0] = gl[0] + r_min * cos(goal_theta + (REF_PI / 2.0));
  glc[1] = gl[1] + r_min * sin(goal_theta + (REF_PI / 2.0));

  double best = REF_INF;
  const char *type = "xxx";
  double D, R, sq, a, b, d0, d1, v0, v1, first, second, third, len;
  double t0[2], t1[2];

  /* rsl :216-254 */
  double rsl_x[2] = {0, 0}, rsl_y[2] = {0, 0};
  d0 = glc[0] - irc[0]; d1 = glc[1] - irc[1];
  D = sqrt(d0 * d0 + d1 * d1);
  v0 = d0 / D; v1 = d1 / D;
  R = -2.0 * r_min / D;
  if (fabs(R) > 1.0) {
  } else {
    sq = sqrt(1.0 - R * R);
    a = r_min * (R * v0 + v1 * sq);
    b = r_min * (R * v1 - v0 * sq);
    rsl_x[0] = irc[0] - a; rsl_x[1] = glc[0] + a;
    rsl_y[0] = irc[1] - b; rsl_y[1] = glc[1] + b;
    t0[0] = rsl_x[0]; t0[1] = rsl_y[0];
    t1[0] = rsl_x[1]; t1[1] = rsl_y[1];
    first = ref_right_turn_dist(il, t0, irc, r_min);
    second = len2(t1[0] - t0[0], t1[1] - t0[1]);
    third = ref_left_turn_dist(t1, gl, glc, r_min);
    len = first + second + third;
    if (best > len) { best = len; type = "rsl"; }
  }

  /* rsr :257-290 */
  double rsr_x[2], rsr_y[2];
  d0 = grc[0] - irc[0]; d1 = grc[1] - irc[1];
  D = sqrt(d0 * d0 + d1 * d1);
  v0 = d0 / D; v1 = d1 / D;
  rsr_x[0] = -r_min * v1 + irc[0]; rsr_x[1] = -r_min * v1 + grc[0];
  rsr_y[0] = r_min * v0 + irc[1];  rsr_y[1] = r_min * v0 + grc[1];
  t0[0] = rsr_x[0]; t0[1] = rsr_y[0];
  t1[0] = rsr_x[1]; t1[1] = rsr_y[1];
  first = ref_right_turn_dist(il, t0, irc, r_min);
  second = len2(t1[0] - t0[0], t1[1] - t0[1]);
  third = ref_right_turn_dist(t1, gl, grc, r_min);
  len = first + second + third;
  if (best > len) { best = len; type = "rsr"; }

  /* rlr :293-328 (D, v from rsr) */
  double rlr_rl[2] = {0, 0}, rlr_lr[2] = {0, 0}, rlr_c[2] = {0, 0};
  if (D < 4.0 * r_min) {
    double theta = -acos(D / (4 * r_min)) + atan2(v1, v0);
    rlr_c[0] = irc[0] + 2 * r_min * cos(theta);
    rlr_c[1] = irc[1] + 2 * r_min * sin(theta);
    rlr_rl[0] = (rlr_c[0] + irc[0]) / 2.0; rlr_rl[1] = (rlr_c[1] + irc[1]) / 2.0;
    rlr_lr[0] = (rlr_c[0] + grc[0]) / 2.0; rlr_lr[1] = (rlr_c[1] + grc[1]) / 2.0;
    first = ref_right_turn_dist(il, rlr_rl, irc, r_min);
    second = ref_left_turn_dist(rlr_rl, rlr_lr, rlr_c, r_min);
    third = ref_right_turn_dist(rlr_lr, gl, grc, r_min);
    len = first + second + third;
    if (best > len) { best = len; type = "rlr"; }
  }

  /* lsr :331-369 */
  double lsr_x[2] = {0, 0}, lsr_y[2] = {0, 0};
  d0 = grc[0] - ilc[0]; d1 = grc[1] - ilc[1];
  D = sqrt(d0 * d0 + d1 * d1);
  v0 = d0 / D; v1 = d1 / D;
  R = 2.0 * r_min / D;
  if (fabs(R) > 1.0) {
  } else {
    sq = sqrt(1.0 - R * R);
    a = R * v0 + v1 * sq;
    b = R * v1 - v0 * sq;
    lsr_x[0] = ilc[0] + a * r_min; lsr_x[1] = grc[0] - a * r_min;
    lsr_y[0] = ilc[1] + b * r_min; lsr_y[1] = grc[1] - b * r_min;
    t0[0] = lsr_x[0]; t0[1] = lsr_y[0];
    t1[0] = lsr_x[1]; t1[1] = lsr_y[1];
    first = ref_left_turn_dist(il, t0, ilc, r_min);
    second = len2(t1[0] - t0[0], t1[1] - t0[1]);
    third = ref_right_turn_dist(t1, gl, grc, r_min);
    len = first + second + third;
    if (best > len) { best = len; type = "lsr"; }
  }

  /* lsl :372-405 */
  double lsl_x[2], lsl_y[2];
  d0 = glc[0] - ilc[0]; d1 = glc[1] - ilc[1];
  D = sqrt(d0 * d0 + d1 * d1);
  v0 = d0 / D; v1 = d1 / D;
  lsl_x[0] = r_min * v1 + ilc[0];  lsl_x[1] = r_min * v1 + glc[0];
  lsl_y[0] = -r_min * v0 + ilc[1]; lsl_y[1] = -r_min * v0 + glc[1];
  t0[0] = lsl_x[0]; t0[1] = lsl_y[0];
  t1[0] = lsl_x[1]; t1[1] = lsl_y[1];
  first = ref_left_turn_dist(il, t0, ilc, r_min);
  second = len2(t1[0] - t0[0], t1[1] - t0[1]);
  third = ref_left_turn_dist(t1, gl, glc, r_min);
  len = first + second + third;
  if (best > len) { best = len; type = "lsl"; }

  /* lrl :408-443 (D, v from lsl; arc lengths R,L,R as written) */
  double lrl_rl[2] = {0, 0}, lrl_lr[2] = {0, 0}, lrl_c[2] = {0, 0};
  if (D < 4.0 * r_min) {
    double theta = -acos(D / (4 * r_min)) + atan2(v1, v0);
    lrl_c[0] = ilc[0] + 2 * r_min * cos(theta);
    lrl_c[1] = ilc[1] + 2 * r_min * sin(theta);
    lrl_lr[0] = (lrl_c[0] + ilc[0]) / 2.0; lrl_lr[1] = (lrl_c[1] + ilc[1]) / 2.0;
    lrl_rl[0] = (lrl_c[0] + glc[0]) / 2.0; lrl_rl[1] = (lrl_c[1] + glc[1]) / 2.0;
    first = ref_right_turn_dist(il, lrl_lr, ilc, r_min);
    second = ref_left_turn_dist(lrl_lr, lrl_rl, lrl_c, r_min);
    third = ref_right_turn_dist(lrl_rl, gl, glc, r_min);
    len = first + second + third;
    if (best > len) { best = len; type = "lrl"; }
  }

  /* :445-449 */
  if (edge_dist <= 2.0 && start[2] == end[2]) type = "0s0";

  int n = 0;
  double *X = out->x, *Y = out->y;
  double p[2];

  /* first part :467-544 */
  if (type[0] == 'r') {
    if (!strcmp(type, "rsl")) { p[0] = rsl_x[0]; p[1] = rsl_y[0]; }
    else if (!strcmp(type, "rsr")) { p[0] = rsr_x[0]; p[1] = rsr_y[0]; }
    else { p[0] = rlr_rl[0]; p[1] = rlr_rl[1]; }
    double ps = atan2(il[1] - irc[1], il[0] - irc[0]);
    double pe = atan2(p[1] - irc[1], p[0] - irc[0]);
    n += sample_arc(irc[0], irc[1], r_min, ps, pe, -1, X, Y, n);
  } else if (type[0] == 'l') {
    if (!strcmp(type, "lsl")) { p[0] = lsl_x[0]; p[1] = lsl_y[0]; }
    else if (!strcmp(type, "lsr")) { p[0] = lsr_x[0]; p[1] = lsr_y[0]; }
    else { p[0] = lrl_lr[0]; p[1] = lrl_lr[1]; }
    double ps = atan2(il[1] - ilc[1], il[0] - ilc[0]);
    double pe = atan2(p[1] - ilc[1], p[0] - ilc[0]);
    n += sample_arc(ilc[0], ilc[1], r_min, ps, pe, +1, X, Y, n);
  }

  /* second part :547-643 */
  if (type[1] == 's') {
    double p1[2], p2[2];
    if (!strcmp(type, "lsr")) { p1[0] = lsr_x[0]; p1[1] = lsr_y[0]; p2[0] = lsr_x[1]; p2[1] = lsr_y[1]; }
    else if (!strcmp(type, "lsl")) { p1[0] = lsl_x[0]; p1[1] = lsl_y[0]; p2[0] = lsl_x[1]; p2[1] = lsl_y[1]; }
    else if (!strcmp(type, "rsr")) { p1[0] = rsr_x[0]; p1[1] = rsr_y[0]; p2[0] = rsr_x[1]; p2[1] = rsr_y[1]; }
    else if (!strcmp(type, "rsl")) { p1[0] = rsl_x[0]; p1[1] = rsl_y[0]; p2[0] = rsl_x[1]; p2[1] = rsl_y[1]; }
    else { p1[0] = start[0]; p1[1] = start[1]; p2[0] = end[0]; p2[1] = end[1]; } /* "0s0" */
    X[n] = p1[0]; Y[n] = p1[1]; n++;
    X[n] = p2[0]; Y[n] = p2[1]; n++;
  } else if (type[1] == 'r') { /* lrl middle :585-613 */
    double ps = atan2(lrl_lr[1] - lrl_c[1], lrl_lr[0] - lrl_c[0]);
    double pe = atan2(lrl_rl[1] - lrl_c[1], lrl_rl[0] - lrl_c[0]);
    n += sample_arc(lrl_c[0], lrl_c[1], r_min, ps, pe, -1, X, Y, n);
  } else if (type[1] == 'l') { /* rlr middle :614-643 */
    double ps = atan2(rlr_rl[1] - rlr_c[1], rlr_rl[0] - rlr_c[0]);
    double pe = atan2(rlr_lr[1] - rlr_c[1], rlr_lr[0] - rlr_c[0]);
    n += sample_arc(rlr_c[0], rlr_c[1], r_min, ps, pe, +1, X, Y, n);
  }

  /* third part :646-723 */
  if (type[2] == 'r') {
    if (!strcmp(type, "rsr")) { p[0] = rsr_x[1]; p[1] = rsr_y[1]; }
    else if (!strcmp(type, "lsr")) { p[0] = lsr_x[1]; p[1] = lsr_y[1]; }
    else { p[0] = rlr_lr[0]; p[1] = rlr_lr[1]; }
    double ps = atan2(p[1] - grc[1], p[0] - grc[0]);
    double pe = atan2(gl[1] - grc[1], gl[0] - grc[0]);
    n += sample_arc(grc[0], grc[1], r_min, ps, pe, -1, X, Y, n);
  } else if (type[2] == 'l') {
    if (!strcmp(type, "lsl")) { p[0] = lsl_x[1]; p[1] = lsl_y[1]; }
    else if (!strcmp(type, "rsl")) { p[0] = rsl_x[1]; p[1] = rsl_y[1]; }
    else { p[0] = lrl_rl[0]; p[1] = lrl_rl[1]; }
    double ps = atan2(p[1] - glc[1], p[0] - glc[0]);
    double pe = atan2(gl[1] - glc[1], gl[0] - glc[0]);
    n += sample_arc(glc[0], glc[1], r_min, ps, pe, +1, X, Y, n);
  }

  strcpy(out->type, type);
  out->npts = n;
  out->dist = best; /* :726, :734-735, :808 (w_dist_ == INF -> dist_ = INF) */
}

/* LineCheck obstacle.cpp:1092-1133: 51 points by repeated +-dist/50 */
void ref_line_trajectory(const double *n1, const double *n2, ref_traj *out) {
  double traj_length = 50;
  double x_val = n1[0], y_val = n1[1];
  double x_dist = fabs(n2[0] - x_val);
  double y_dist = fabs(n2[1] - y_val);
  int i = 0;
  while (i < traj_length) {
    out->x[i] = x_val;
    if (n2[0] > n1[0]) x_val += x_dist / traj_length;
    else x_val -= x_dist / traj_length;
    out->y[i] = y_val;
    if (n2[1] > n1[1]) y_val += y_dist / traj_length;
    else y_val -= y_dist / traj_length;
    i++;
  }
  out->x[i] = x_val;
  out->y[i] = y_val;
  out->npts = 51;
  strcpy(out->type, "lin");
  out->dist = 0.0;
}
