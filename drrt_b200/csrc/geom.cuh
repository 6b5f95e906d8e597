// geom.cuh -- device restatements of the reference's 2-D geometry helpers
// (src/distancefunctions.cpp) and of the Dubins candidate solve
// (DubinsEdge::CalculateTrajectory, src/dubinsedge.cpp:172-449), in the
// reference's floating-point operation order.  No FMA contraction
// (-fmad=false); sin/cos/atan2/acos are CUDA's double-precision routines,
// which differ from glibc by at most a few ulp (DESIGN.md "Trig parity").
#pragma once

#include "common.cuh"

namespace drrt {

// DistanceSqrdPointToSegment  distancefunctions.cpp:50-74
__device__ __forceinline__ double dist_sqrd_point_segment(double px, double py, double sx,
                                                          double sy, double ex, double ey) {
  double vx = px - sx, vy = py - sy;
  double ux = ex - sx, uy = ey - sy;
  double determinate = vx * ux + vy * uy;
  if (determinate <= 0) return vx * vx + vy * vy;
  double len = ux * ux + uy * uy;
  if (determinate >= len) return (ex - px) * (ex - px) + (ey - py) * (ey - py);
  double c = ux * vy - uy * vx;
  return c * c / len;
}

// The same three values and the same choice among them, computed side by side and selected: lanes of
// a warp take different branches of the function above nearly every time it is called for a handful
// of (segment, polygon edge) pairs, so the warp runs all three paths anyway -- one after the other,
// behind branches that wait to be resolved.  Straight-line, the four calls of SegmentDistSqrd
// overlap their dependent chains (the division above all).  Bit-identical results: every value is
// computed by the same operations, a quotient that is not selected may be anything.  Used where a
// warp tests few pairs at a time (the cooperative test of the walk: the cull and SegmentDistSqrd's
// four calls, 7.96 -> 7.61 ms per 10 M edges); everywhere else it measured the same as the branches.
// (The cull as c*c > far2*len, without the quotient, measured SLOWER again: 7.92.)
__device__ __forceinline__ double dist_sqrd_point_segment_sel(double px, double py, double sx,
                                                              double sy, double ex, double ey) {
  const double vx = px - sx, vy = py - sy;
  const double ux = ex - sx, uy = ey - sy;
  const double determinate = vx * ux + vy * uy;
  const double r0 = vx * vx + vy * vy;
  const double len = ux * ux + uy * uy;
  const double r1 = (ex - px) * (ex - px) + (ey - py) * (ey - py);
  const double c = ux * vy - uy * vx;
  const double r2 = c * c / len;
  return (determinate <= 0) ? r0 : ((determinate >= len) ? r1 : r2);
}
// SegmentDistSqrd  distancefunctions.cpp:94-164
__device__ __forceinline__ double segment_dist_sqrd(double pax, double pay, double pbx, double pby,
                                                    double qax, double qay, double qbx, double qby) {
  bool possible = true;
  double m, diffA, diffB;
  if (fabs(pbx - pax) < 0.000001) {
    if ((qax >= pax && qbx >= pax) || (qax <= pax && qbx <= pax)) possible = false;
  } else {
    m = (pby - pay) / (pbx - pax);
    diffA = (m * (qax - pax) + pay) - qay;
    diffB = (m * (qbx - pax) + pay) - qby;
    if ((diffA > 0.0 && diffB > 0.0) || (diffA < 0.0 && diffB < 0.0)) possible = false;
  }
  if (possible) {
    if (fabs(qbx - qax) < 0.000001) {
      if ((pax >= qax && pbx >= qax) || (pax <= qax && pbx <= qax)) possible = false;
    } else {
      m = (qby - qay) / (qbx - qax);
      diffA = (m * (pax - qax) + qay) - pay;
      diffB = (m * (pbx - qax) + qay) - pby;
      if ((diffA > 0.0 && diffB > 0.0) || (diffA < 0.0 && diffB < 0.0)) possible = false;
    }
  }
  if (possible) return 0.0;
  double d0 = dist_sqrd_point_segment_sel(pax, pay, qax, qay, qbx, qby);
  double d1 = dist_sqrd_point_segment_sel(pbx, pby, qax, qay, qbx, qby);
  double d2 = dist_sqrd_point_segment_sel(qax, qay, pax, pay, pbx, pby);
  double d3 = dist_sqrd_point_segment_sel(qbx, qby, pax, pay, pbx, pby);
  double mn = d0;
  if (d1 < mn) mn = d1;
  if (d2 < mn) mn = d2;
  if (d3 < mn) mn = d3;
  return mn;
}

// RightTurnDist / LeftTurnDist  distancefunctions.cpp:24-46
__device__ __forceinline__ double right_turn_dist(double ax, double ay, double bx, double by,
                                                  double cx, double cy, double r) {
  double theta = atan2(ay - cy, ax - cx) - atan2(by - cy, bx - cx);
  if (theta < 0) theta += 2 * 3.1415926536;
  return theta * r;
}
__device__ __forceinline__ double left_turn_dist(double ax, double ay, double bx, double by,
                                                 double cx, double cy, double r) {
  double theta = atan2(by - cy, bx - cx) - atan2(ay - cy, ax - cx);
  if (theta < 0) theta += 2 * 3.1415926536;
  return theta * r;
}

// PointInPolygon  distancefunctions.cpp:166-205
__device__ __forceinline__ bool point_in_polygon(double px, double py, const double *vx,
                                                 const double *vy, int nv) {
  if (nv < 2) return false;
  int crossings = 0;
  double sx = vx[nv - 1], sy = vy[nv - 1];
  for (int i = 0; i < nv; i++) {
    double ex = vx[i], ey = vy[i];
    if ((sy > py && ey < py) || (sy < py && ey > py)) {
      if (sx > px && ex > px) {
        crossings++;
      } else if (sx < px && ex < px) {
      } else {
        double mx = (sx < ex) ? ex : sx;
        double t = 2 * mx;
        double x = (-((sx * ey - sy * ex) * (px - t)) + ((sx - ex) * (px * py - py * t))) /
                   ((sy - ey) * (px - t));
        if (x > px) crossings++;
      }
    }
    sx = ex;
    sy = ey;
  }
  return (crossings % 2) == 1;
}

// DistToPolygonSqrd  distancefunctions.cpp:77-92
__device__ __forceinline__ double dist_to_polygon_sqrd(double px, double py, const double *vx,
                                                       const double *vy, int nv) {
  double mn = DRRT_INF;
  double sx = vx[nv - 1], sy = vy[nv - 1];
  for (int i = 0; i < nv; i++) {
    double ex = vx[i], ey = vy[i];
    double d = dist_sqrd_point_segment(px, py, sx, sy, ex, ey);
    if (d < mn) mn = d;
    sx = ex;
    sy = ey;
  }
  return mn;
}

// ---------------------------------------------------------------------------
// Dubins word solved by CalculateTrajectory.  Each of the (up to) three parts
// is an arc (centre, start/end angle as atan2 gives them, turn direction) or
// the straight piece (two points).
enum { TRAJ_XXX = 0, TRAJ_RSL, TRAJ_RSR, TRAJ_RLR, TRAJ_LSR, TRAJ_LSL, TRAJ_LRL, TRAJ_0S0 };
enum { PART_NONE = 0, PART_ARC_R = 1, PART_ARC_L = 2, PART_LINE = 3 };

struct DubinsPart {
  double cx, cy;  // arc centre            | line: first point
  double a, b;    // arc phi_start/phi_end | line: second point
};

// 112 bytes; also the record the solve kernel hands to the walk kernel
struct __align__(16) DubinsPath {
  double dist;  // edge->dist_ (dubinsedge.cpp:808; INF when no candidate exists)
  int type;
  int kinds;    // part kinds, 2 bits each
  DubinsPart part[3];
  __host__ __device__ __forceinline__ int kind(int k) const { return (kinds >> (2 * k)) & 3; }
  // bit 8+k set: no obstacle comes near part k (edge check only); its segments need no test
  __host__ __device__ __forceinline__ bool idle(int k) const { return (kinds >> (8 + k)) & 1; }
};

__device__ __forceinline__ double len2(double dx, double dy) { return sqrt(dx * dx + dy * dy); }

__device__ __forceinline__ void arc_part(DubinsPath &out, int k, int kind, double cx, double cy,
                                         double fx, double fy, double tx, double ty) {
  DubinsPart &P = out.part[k];
  out.kinds |= kind << (2 * k);
  P.cx = cx;
  P.cy = cy;
  P.a = atan2(fy - cy, fx - cx);
  P.b = atan2(ty - cy, tx - cx);
}

// Candidate bits, in the order CalculateTrajectory tries them.
enum { CAND_RSL = 1, CAND_RSR = 2, CAND_RLR = 4, CAND_LSR = 8, CAND_LSL = 16, CAND_LRL = 32, CAND_ALL = 63 };

// edge_dist = Edge::NewEdge's dist_ (dubinsedge.cpp:19), read by the "0s0" test.
// `mask` lists the candidates to evaluate (CAND_ALL: all six, as the reference does).  The
// winner is the first candidate, in the reference's order, of smallest length under the strict
// `bestDist > len` test; leaving out candidates that are provably longer than the shortest
// one changes neither the winner nor bestDist (dubins_rank below decides that, conservatively).
__device__ void dubins_solve(double sx, double sy, double st, double ex, double ey, double et,
                             double r_min, double edge_dist, DubinsPath &out, int mask = CAND_ALL) {
  // :183-202 turn-circle centres (each only if a listed candidate uses it)
  double sn, cs;
  double ircx = 0, ircy = 0, ilcx = 0, ilcy = 0, grcx = 0, grcy = 0, glcx = 0, glcy = 0;
  if (mask & (CAND_RSL | CAND_RSR | CAND_RLR)) {
    sincos(st - (DRRT_PI / 2.0), &sn, &cs);
    ircx = sx + r_min * cs; ircy = sy + r_min * sn;
  }
  if (mask & (CAND_LSR | CAND_LSL | CAND_LRL)) {
    sincos(st + (DRRT_PI / 2.0), &sn, &cs);
    ilcx = sx + r_min * cs; ilcy = sy + r_min * sn;
  }
  if (mask & (CAND_RSR | CAND_RLR | CAND_LSR)) {
    sincos(et - (DRRT_PI / 2.0), &sn, &cs);
    grcx = ex + r_min * cs; grcy = ey + r_min * sn;
  }
  if (mask & (CAND_RSL | CAND_LSL | CAND_LRL)) {
    sincos(et + (DRRT_PI / 2.0), &sn, &cs);
    glcx = ex + r_min * cs; glcy = ey + r_min * sn;
  }

  double best = DRRT_INF;
  int type = TRAJ_XXX;
  double D = 0, R, sq, a, b, d0, d1, v0 = 0, v1 = 0, first, second, third, len;

  // rsl :216-254
  double rsl_x0 = 0, rsl_x1 = 0, rsl_y0 = 0, rsl_y1 = 0;
  if (mask & CAND_RSL) {
    d0 = glcx - ircx; d1 = glcy - ircy;
    D = sqrt(d0 * d0 + d1 * d1);
    v0 = d0 / D; v1 = d1 / D;
    R = -2.0 * r_min / D;
    if (!(fabs(R) > 1.0)) {
      sq = sqrt(1.0 - R * R);
      a = r_min * (R * v0 + v1 * sq);
      b = r_min * (R * v1 - v0 * sq);
      rsl_x0 = ircx - a; rsl_x1 = glcx + a;
      rsl_y0 = ircy - b; rsl_y1 = glcy + b;
      first = right_turn_dist(sx, sy, rsl_x0, rsl_y0, ircx, ircy, r_min);
      second = len2(rsl_x1 - rsl_x0, rsl_y1 - rsl_y0);
      third = left_turn_dist(rsl_x1, rsl_y1, ex, ey, glcx, glcy, r_min);
      len = first + second + third;
      if (best > len) { best = len; type = TRAJ_RSL; }
    }
  }

  // rsr :257-290
  double rsr_x0 = 0, rsr_x1 = 0, rsr_y0 = 0, rsr_y1 = 0;
  if (mask & (CAND_RSR | CAND_RLR)) {
    d0 = grcx - ircx; d1 = grcy - ircy;
    D = sqrt(d0 * d0 + d1 * d1);
    v0 = d0 / D; v1 = d1 / D;
  }
  if (mask & CAND_RSR) {
    rsr_x0 = -r_min * v1 + ircx; rsr_x1 = -r_min * v1 + grcx;
    rsr_y0 = r_min * v0 + ircy; rsr_y1 = r_min * v0 + grcy;
    first = right_turn_dist(sx, sy, rsr_x0, rsr_y0, ircx, ircy, r_min);
    second = len2(rsr_x1 - rsr_x0, rsr_y1 - rsr_y0);
    third = right_turn_dist(rsr_x1, rsr_y1, ex, ey, grcx, grcy, r_min);
    len = first + second + third;
    if (best > len) { best = len; type = TRAJ_RSR; }
  }

  // rlr :293-328 (D, v of rsr)
  double rlr_rlx = 0, rlr_rly = 0, rlr_lrx = 0, rlr_lry = 0, rlr_cx = 0, rlr_cy = 0;
  if ((mask & CAND_RLR) && D < 4.0 * r_min) {
    double theta = -acos(D / (4 * r_min)) + atan2(v1, v0);
    sincos(theta, &sn, &cs);
    rlr_cx = ircx + 2 * r_min * cs;
    rlr_cy = ircy + 2 * r_min * sn;
    rlr_rlx = (rlr_cx + ircx) / 2.0; rlr_rly = (rlr_cy + ircy) / 2.0;
    rlr_lrx = (rlr_cx + grcx) / 2.0; rlr_lry = (rlr_cy + grcy) / 2.0;
    first = right_turn_dist(sx, sy, rlr_rlx, rlr_rly, ircx, ircy, r_min);
    second = left_turn_dist(rlr_rlx, rlr_rly, rlr_lrx, rlr_lry, rlr_cx, rlr_cy, r_min);
    third = right_turn_dist(rlr_lrx, rlr_lry, ex, ey, grcx, grcy, r_min);
    len = first + second + third;
    if (best > len) { best = len; type = TRAJ_RLR; }
  }

  // lsr :331-369
  double lsr_x0 = 0, lsr_x1 = 0, lsr_y0 = 0, lsr_y1 = 0;
  if (mask & CAND_LSR) {
    d0 = grcx - ilcx; d1 = grcy - ilcy;
    D = sqrt(d0 * d0 + d1 * d1);
    v0 = d0 / D; v1 = d1 / D;
    R = 2.0 * r_min / D;
    if (!(fabs(R) > 1.0)) {
      sq = sqrt(1.0 - R * R);
      a = R * v0 + v1 * sq;
      b = R * v1 - v0 * sq;
      lsr_x0 = ilcx + a * r_min; lsr_x1 = grcx - a * r_min;
      lsr_y0 = ilcy + b * r_min; lsr_y1 = grcy - b * r_min;
      first = left_turn_dist(sx, sy, lsr_x0, lsr_y0, ilcx, ilcy, r_min);
      second = len2(lsr_x1 - lsr_x0, lsr_y1 - lsr_y0);
      third = right_turn_dist(lsr_x1, lsr_y1, ex, ey, grcx, grcy, r_min);
      len = first + second + third;
      if (best > len) { best = len; type = TRAJ_LSR; }
    }
  }

  // lsl :372-405
  double lsl_x0 = 0, lsl_x1 = 0, lsl_y0 = 0, lsl_y1 = 0;
  if (mask & (CAND_LSL | CAND_LRL)) {
    d0 = glcx - ilcx; d1 = glcy - ilcy;
    D = sqrt(d0 * d0 + d1 * d1);
    v0 = d0 / D; v1 = d1 / D;
  }
  if (mask & CAND_LSL) {
    lsl_x0 = r_min * v1 + ilcx; lsl_x1 = r_min * v1 + glcx;
    lsl_y0 = -r_min * v0 + ilcy; lsl_y1 = -r_min * v0 + glcy;
    first = left_turn_dist(sx, sy, lsl_x0, lsl_y0, ilcx, ilcy, r_min);
    second = len2(lsl_x1 - lsl_x0, lsl_y1 - lsl_y0);
    third = left_turn_dist(lsl_x1, lsl_y1, ex, ey, glcx, glcy, r_min);
    len = first + second + third;
    if (best > len) { best = len; type = TRAJ_LSL; }
  }

  // lrl :408-443 (D, v of lsl; arc lengths R,L,R as the reference writes them)
  double lrl_rlx = 0, lrl_rly = 0, lrl_lrx = 0, lrl_lry = 0, lrl_cx = 0, lrl_cy = 0;
  if ((mask & CAND_LRL) && D < 4.0 * r_min) {
    double theta = -acos(D / (4 * r_min)) + atan2(v1, v0);
    sincos(theta, &sn, &cs);
    lrl_cx = ilcx + 2 * r_min * cs;
    lrl_cy = ilcy + 2 * r_min * sn;
    lrl_lrx = (lrl_cx + ilcx) / 2.0; lrl_lry = (lrl_cy + ilcy) / 2.0;
    lrl_rlx = (lrl_cx + glcx) / 2.0; lrl_rly = (lrl_cy + glcy) / 2.0;
    first = right_turn_dist(sx, sy, lrl_lrx, lrl_lry, ilcx, ilcy, r_min);
    second = left_turn_dist(lrl_lrx, lrl_lry, lrl_rlx, lrl_rly, lrl_cx, lrl_cy, r_min);
    third = right_turn_dist(lrl_rlx, lrl_rly, ex, ey, glcx, glcy, r_min);
    len = first + second + third;
    if (best > len) { best = len; type = TRAJ_LRL; }
  }

  if (edge_dist <= 2.0 && st == et) type = TRAJ_0S0;  // :445-449

  out.type = type;
  out.dist = best;
  out.kinds = 0;
  DubinsPart &P1 = out.part[1];
  switch (type) {  // :467-723
    case TRAJ_RSL:
      arc_part(out, 0, PART_ARC_R, ircx, ircy, sx, sy, rsl_x0, rsl_y0);
      out.kinds |= PART_LINE << 2; P1.cx = rsl_x0; P1.cy = rsl_y0; P1.a = rsl_x1; P1.b = rsl_y1;
      arc_part(out, 2, PART_ARC_L, glcx, glcy, rsl_x1, rsl_y1, ex, ey);
      break;
    case TRAJ_RSR:
      arc_part(out, 0, PART_ARC_R, ircx, ircy, sx, sy, rsr_x0, rsr_y0);
      out.kinds |= PART_LINE << 2; P1.cx = rsr_x0; P1.cy = rsr_y0; P1.a = rsr_x1; P1.b = rsr_y1;
      arc_part(out, 2, PART_ARC_R, grcx, grcy, rsr_x1, rsr_y1, ex, ey);
      break;
    case TRAJ_RLR:
      arc_part(out, 0, PART_ARC_R, ircx, ircy, sx, sy, rlr_rlx, rlr_rly);
      arc_part(out, 1, PART_ARC_L, rlr_cx, rlr_cy, rlr_rlx, rlr_rly, rlr_lrx, rlr_lry);
      arc_part(out, 2, PART_ARC_R, grcx, grcy, rlr_lrx, rlr_lry, ex, ey);
      break;
    case TRAJ_LSR:
      arc_part(out, 0, PART_ARC_L, ilcx, ilcy, sx, sy, lsr_x0, lsr_y0);
      out.kinds |= PART_LINE << 2; P1.cx = lsr_x0; P1.cy = lsr_y0; P1.a = lsr_x1; P1.b = lsr_y1;
      arc_part(out, 2, PART_ARC_R, grcx, grcy, lsr_x1, lsr_y1, ex, ey);
      break;
    case TRAJ_LSL:
      arc_part(out, 0, PART_ARC_L, ilcx, ilcy, sx, sy, lsl_x0, lsl_y0);
      out.kinds |= PART_LINE << 2; P1.cx = lsl_x0; P1.cy = lsl_y0; P1.a = lsl_x1; P1.b = lsl_y1;
      arc_part(out, 2, PART_ARC_L, glcx, glcy, lsl_x1, lsl_y1, ex, ey);
      break;
    case TRAJ_LRL:
      arc_part(out, 0, PART_ARC_L, ilcx, ilcy, sx, sy, lrl_lrx, lrl_lry);
      arc_part(out, 1, PART_ARC_R, lrl_cx, lrl_cy, lrl_lrx, lrl_lry, lrl_rlx, lrl_rly);
      arc_part(out, 2, PART_ARC_L, glcx, glcy, lrl_rlx, lrl_rly, ex, ey);
      break;
    case TRAJ_0S0:
      out.kinds |= PART_LINE << 2; P1.cx = sx; P1.cy = sy; P1.a = ex; P1.b = ey;
      break;
    default:
      break;
  }
}

// Which candidates can be the shortest?  A float32 pass over all six (start pose moved to the
// origin first, in double, so that magnitudes stay small) ranks them; a candidate is listed
// when it may exist and its float length is within `margin` of the shortest candidate that
// surely exists, or when the float pass cannot tell: it sits on an existence boundary (D near
// 2r / 4r), has a turn angle near the 0 / 2pi jump of Right/LeftTurnDist, a degenerate centre
// distance, or is not finite.  The bands and the margin are several times the float error budget
// spelled out in the function.  The listed candidates are then
// solved in double exactly as the reference does; the others are provably longer than the
// winner, so the winner and bestDist are the reference's (dubins_solve).  The float pass uses
// fast approximations (rank_atan2 1.2e-5 rad, __sincosf, __fdividef): their errors are part of
// the same budget.
// atan2 to ~1.2e-5 rad (Hastings' five-term odd polynomial on [0,1] plus octant folding): the
// ranking only needs angles to well under its 1e-3 bands
__device__ __forceinline__ float rank_atan2(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float a = __fdividef(mn, mx);  // NaN for 0/0: the candidate then counts as "cannot tell"
  const float s = a * a;
  float r = a * (0.9998660f + s * (-0.3302995f + s * (0.1801410f + s * (-0.0851330f + s * 0.0208351f))));
  if (ay > ax) r = 1.57079632679f - r;
  if (x < 0.f) r = 3.14159265359f - r;
  return (y < 0.f) ? -r : r;
}

__device__ __forceinline__ void rank_sincos(float x, float *sn, float *cs) {
  x -= 6.28318530718f * rintf(x * 0.15915494309f);  // into [-pi, pi], where __sincosf is accurate to ~1e-6
  __sincosf(x, sn, cs);
}

__device__ __forceinline__ int dubins_rank(double sx, double sy, double st, double ex, double ey,
                                           double et, double r_min) {
  const float TWO_PI_F = 6.28318530718f;
  const float r = (float)r_min;
  const float gx = (float)(ex - sx), gy = (float)(ey - sy);
  const float scale = fmaxf(fmaxf(fabsf(gx), fabsf(gy)), fabsf(r));
  // float32 is only trusted for ordinary geometry
  if (!(r > 0.0f) || !(scale <= 64.0f * r) || !(fabs(st) <= 64.0) || !(fabs(et) <= 64.0)) return CAND_ALL;
  // error budget of the float pass: rank_atan2 1.2e-5 rad per angle, a turn is the difference of
  // two (2.5e-5 with the rounding of its points), a length three of those times r plus ~1e-6 of
  // segment: < 1e-4 * (r + scale), and two lengths are compared
  const float band = 2e-4f * (r + scale);    // existence boundaries, degenerate D
  const float margin = 5e-4f * (r + scale);  // length comparison
  const float eps_th = 2e-4f;                // turn angles near the 0 / 2pi jump
  float ss, cs0, se, ce;
  rank_sincos((float)st, &ss, &cs0);
  rank_sincos((float)et, &se, &ce);
  // turn-circle centres relative to the start point (dubinsedge.cpp:183-202)
  const float ircx = r * ss, ircy = -r * cs0, ilcx = -r * ss, ilcy = r * cs0;
  const float grcx = gx + r * se, grcy = gy - r * ce, glcx = gx - r * se, glcy = gy + r * ce;
  bool unsure[6] = {false, false, false, false, false, false};
  bool exists[6] = {false, true, false, false, true, false};
  float len[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto turn = [&](float ax, float ay, float bx, float by, float cx, float cy, bool right, bool &uns) {
    const float ta = rank_atan2(ay - cy, ax - cx), tb = rank_atan2(by - cy, bx - cx);
    float th = right ? ta - tb : tb - ta;
    if (fabsf(th) < eps_th || fabsf(fabsf(th) - TWO_PI_F) < eps_th) uns = true;
    if (th < 0.f) th += TWO_PI_F;
    return th * r;
  };
  auto norm = [](float x, float y) { return sqrtf(x * x + y * y); };
  float d0, d1, D, v0, v1;
  // rsl
  d0 = glcx - ircx; d1 = glcy - ircy; D = norm(d0, d1);
  if (D > 2.f * r - band) {
    exists[0] = true;
    if (D < 2.f * r + band) unsure[0] = true;
    v0 = d0 / D; v1 = d1 / D;
    const float R = fminf(fmaxf(-2.f * r / D, -1.f), 1.f), sq = sqrtf(fmaxf(1.f - R * R, 0.f));
    const float a = r * (R * v0 + v1 * sq), b = r * (R * v1 - v0 * sq);
    const float x0 = ircx - a, y0 = ircy - b, x1 = glcx + a, y1 = glcy + b;
    len[0] = turn(0.f, 0.f, x0, y0, ircx, ircy, true, unsure[0]) + norm(x1 - x0, y1 - y0) +
             turn(x1, y1, gx, gy, glcx, glcy, false, unsure[0]);
  }
  // rsr, rlr
  d0 = grcx - ircx; d1 = grcy - ircy; D = norm(d0, d1);
  v0 = d0 / D; v1 = d1 / D;
  if (D < band) unsure[1] = true;
  {
    const float x0 = -r * v1 + ircx, y0 = r * v0 + ircy, x1 = -r * v1 + grcx, y1 = r * v0 + grcy;
    len[1] = turn(0.f, 0.f, x0, y0, ircx, ircy, true, unsure[1]) + norm(x1 - x0, y1 - y0) +
             turn(x1, y1, gx, gy, grcx, grcy, true, unsure[1]);
  }
  if (D < 4.f * r + band) {
    exists[2] = true;
    if (D > 4.f * r - band || D < band) unsure[2] = true;
    const float theta = -acosf(fminf(D / (4.f * r), 1.f)) + rank_atan2(v1, v0);
    float sn, cs;
    rank_sincos(theta, &sn, &cs);
    const float cx = ircx + 2.f * r * cs, cy = ircy + 2.f * r * sn;
    const float rlx = (cx + ircx) * 0.5f, rly = (cy + ircy) * 0.5f, lrx = (cx + grcx) * 0.5f, lry = (cy + grcy) * 0.5f;
    len[2] = turn(0.f, 0.f, rlx, rly, ircx, ircy, true, unsure[2]) +
             turn(rlx, rly, lrx, lry, cx, cy, false, unsure[2]) +
             turn(lrx, lry, gx, gy, grcx, grcy, true, unsure[2]);
  }
  // lsr
  d0 = grcx - ilcx; d1 = grcy - ilcy; D = norm(d0, d1);
  if (D > 2.f * r - band) {
    exists[3] = true;
    if (D < 2.f * r + band) unsure[3] = true;
    v0 = d0 / D; v1 = d1 / D;
    const float R = fminf(fmaxf(2.f * r / D, -1.f), 1.f), sq = sqrtf(fmaxf(1.f - R * R, 0.f));
    const float a = R * v0 + v1 * sq, b = R * v1 - v0 * sq;
    const float x0 = ilcx + a * r, y0 = ilcy + b * r, x1 = grcx - a * r, y1 = grcy - b * r;
    len[3] = turn(0.f, 0.f, x0, y0, ilcx, ilcy, false, unsure[3]) + norm(x1 - x0, y1 - y0) +
             turn(x1, y1, gx, gy, grcx, grcy, true, unsure[3]);
  }
  // lsl, lrl
  d0 = glcx - ilcx; d1 = glcy - ilcy; D = norm(d0, d1);
  v0 = d0 / D; v1 = d1 / D;
  if (D < band) unsure[4] = true;
  {
    const float x0 = r * v1 + ilcx, y0 = -r * v0 + ilcy, x1 = r * v1 + glcx, y1 = -r * v0 + glcy;
    len[4] = turn(0.f, 0.f, x0, y0, ilcx, ilcy, false, unsure[4]) + norm(x1 - x0, y1 - y0) +
             turn(x1, y1, gx, gy, glcx, glcy, false, unsure[4]);
  }
  if (D < 4.f * r + band) {
    exists[5] = true;
    if (D > 4.f * r - band || D < band) unsure[5] = true;
    const float theta = -acosf(fminf(D / (4.f * r), 1.f)) + rank_atan2(v1, v0);
    float sn, cs;
    rank_sincos(theta, &sn, &cs);
    const float cx = ilcx + 2.f * r * cs, cy = ilcy + 2.f * r * sn;
    const float lrx = (cx + ilcx) * 0.5f, lry = (cy + ilcy) * 0.5f, rlx = (cx + glcx) * 0.5f, rly = (cy + glcy) * 0.5f;
    len[5] = turn(0.f, 0.f, lrx, lry, ilcx, ilcy, true, unsure[5]) +   // R, L, R as the reference writes them
             turn(lrx, lry, rlx, rly, cx, cy, false, unsure[5]) +
             turn(rlx, rly, gx, gy, glcx, glcy, true, unsure[5]);
  }
  float best = 3.0e38f;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    if (exists[k] && !(len[k] == len[k] && fabsf(len[k]) < 1.0e30f)) unsure[k] = true;  // not finite
    if (exists[k] && !unsure[k]) best = fminf(best, len[k]);
  }
  int mask = 0;
#pragma unroll
  for (int k = 0; k < 6; k++)
    if (exists[k] && (unsure[k] || len[k] <= best + margin)) mask |= 1 << k;
  return mask ? mask : CAND_ALL;
}

// Walks the polyline of a solved path point by point, in the order
// CalculateTrajectory stores it (:484-502 and siblings; :577-583).
struct PolylineWalker {
  const DubinsPath *path;  // local or global
  double r_min;
  int part;       // current part, 3 = done
  int i, n;       // index / count within the part
  int kind;       // current part's kind and fields, cached in registers
  double cx, cy, a, b;
  double val, phi_end;
  bool ended, equal;
  bool skip_idle;  // fast-forward parts flagged idle to their last point
  bool prepared;   // the record went through prepare(): end angles adjusted, entry counts in `type`

  // end angle after the +-2pi adjustment and number of entries of an arc (:484-502 and siblings)
  __device__ __forceinline__ static void arc_extent(int kind, double a, double &pe, int &n) {
    if (kind == PART_ARC_R) { if (pe > a) pe -= 2.0 * DRRT_PI; }
    else { if (pe < a) pe += 2.0 * DRRT_PI; }
    n = (int)(ceil(fabs(pe - a) / 0.1) + 1);
    if (n > DRRT_TRAJ_MAX) n = DRRT_TRAJ_MAX;
  }
  // Does begin_part's arithmetic once, where every lane has a path (the solve kernels), instead of
  // in the walk, where a lane entering a new part runs it alone: every arc's `b` becomes its
  // adjusted end angle and `type` (which the walk does not read) its entry counts, 8 bits a part.
  __device__ __forceinline__ static void prepare(DubinsPath &P) {
    int packed = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int kind = P.kind(k);
      if (kind == PART_ARC_R || kind == PART_ARC_L) {
        int cnt;
        arc_extent(kind, P.part[k].a, P.part[k].b, cnt);
        packed |= cnt << (8 * k);
      }
    }
    P.type = packed;
  }
  bool fresh;      // the next point does not continue a tested polyline segment

  __device__ __forceinline__ void begin_part() {
    while (part < 3) {
      kind = path->kind(part);
      if (kind != PART_NONE) {
        const DubinsPart &P = path->part[part];
        cx = P.cx; cy = P.cy; a = P.a; b = P.b;
        if (kind == PART_LINE) { n = 2; i = 0; return; }
        double pe = b;
        if (prepared) {
          n = (path->type >> (8 * part)) & 0xff;
        } else {
          arc_extent(kind, a, pe, n);
        }
        phi_end = pe;
        equal = (pe == a);
        val = a;
        ended = false;
        i = 0;
        if (n > 0) return;
      }
      part++;
    }
  }
  __device__ __forceinline__ void start(const DubinsPath *p, double r, bool skip = false, bool prep = false) {
    path = p; r_min = r; part = 0; i = 0; n = 0; kind = PART_NONE; fresh = true; skip_idle = skip;
    prepared = prep;
    begin_part();
  }
  __device__ __forceinline__ bool done() const { return part >= 3; }
  // *connects: the segment from the previous point to this one is part of the polyline to test
  __device__ __forceinline__ bool next(double *x, double *y, bool *connects = nullptr) {
    if (part >= 3) return false;
    if (connects) *connects = !fresh;
    fresh = false;
    if (kind == PART_LINE) {
      if (i == 0) { *x = cx; *y = cy; } else { *x = a; *y = b; }
    } else {
      double phi;
      if (equal) {
        phi = (i == 0) ? a : 0.0;
      } else if (!ended) {
        bool in = (kind == PART_ARC_R) ? (val >= phi_end) : (val <= phi_end);
        if (in) { phi = val; if (kind == PART_ARC_R) val -= 0.1; else val += 0.1; }
        else { phi = phi_end; ended = true; }
      } else {
        phi = 0.0;  // entries the reference's loop never reaches stay Zero()
      }
      double sn, cs;
      sincos(phi, &sn, &cs);
      *x = cx + r_min * cs;
      *y = cy + r_min * sn;
    }
    i++;
    if (i == 1 && skip_idle && n > 2 && kind != PART_LINE && path->idle(part)) {
      // nothing can collide inside this arc's circle: after its first point (whose
      // segment from the previous part is still tested) run the angle recurrence
      // without sin/cos up to the last point, which starts the next junction segment
      if (equal) {
        i = n - 1;
      } else {
        // Unless |dphi|/0.1 sits on an integer (to within the rounding the recurrence can
        // accumulate) the loop provably runs n-1 times and the last entry is phi_end.
        double steps = fabs(phi_end - a) / 0.1;
        double frac = steps - floor(steps);
        if (frac > 1e-9 && frac < 1.0 - 1e-9) {
          i = n - 1;
          ended = false;
          val = (kind == PART_ARC_R) ? phi_end - 1.0 : phi_end + 1.0;  // next() emits phi_end
        } else {
          for (; i < n - 1; i++) {
            if (ended) continue;
            bool in = (kind == PART_ARC_R) ? (val >= phi_end) : (val <= phi_end);
            if (in) { if (kind == PART_ARC_R) val -= 0.1; else val += 0.1; }
            else ended = true;
          }
        }
      }
      fresh = true;
    }
    if (i >= n) { part++; begin_part(); }
    return true;
  }
};

}  // namespace drrt
