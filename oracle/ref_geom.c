/* ref_geom.c -- TEST ORACLE (see oracle.h).  Restatement of the 2-D geometry
 * helpers in src/distancefunctions.cpp of cosa0481/DRRT, operation order kept.
 */
#include "oracle.h"

#include <math.h>

/* RightTurnDist distancefunctions.cpp:24-34 */
double ref_right_turn_dist(const double *a, const double *b, const double *c,
                           double r) {
  double theta = atan2(a[1] - c[1], a[0] - c[0]) - atan2(b[1] - c[1], b[0] - c[0]);
  if (theta < 0) theta += 2 * 3.1415926536;
  return theta * r;
}

/* LeftTurnDist distancefunctions.cpp:36-46 */
double ref_left_turn_dist(const double *a, const double *b, const double *c,
                          double r) {
  double theta = atan2(b[1] - c[1], b[0] - c[0]) - atan2(a[1] - c[1], a[0] - c[0]);
  if (theta < 0) theta += 2 * 3.1415926536;
  return theta * r;
}

/* DistanceSqrdPointToSegment distancefunctions.cpp:50-74 */
double ref_dist_sqrd_point_segment(const double *pt, const double *s,
                                   const double *e) {
  double vx = pt[0] - s[0];
  double vy = pt[1] - s[1];
  double ux = e[0] - s[0];
  double uy = e[1] - s[1];
  double determinate = vx * ux + vy * uy;
  if (determinate <= 0) {
    return vx * vx + vy * vy;
  } else {
    double len = ux * ux + uy * uy;
    if (determinate >= len) {
      return (e[0] - pt[0]) * (e[0] - pt[0]) + (e[1] - pt[1]) * (e[1] - pt[1]);
    } else {
      return (ux * vy - uy * vx) * (ux * vy - uy * vx) / len;
    }
  }
}

/* DistToPolygonSqrd distancefunctions.cpp:77-92 */
double ref_dist_to_polygon_sqrd(const double *pt, const double *poly, int nv) {
  double min_dist_sqrd = REF_INF;
  const double *start = poly + 2 * (nv - 1);
  for (int i = 0; i < nv; i++) {
    const double *end = poly + 2 * i;
    double d = ref_dist_sqrd_point_segment(pt, start, end);
    if (d < min_dist_sqrd) min_dist_sqrd = d;
    start = end;
  }
  return min_dist_sqrd;
}

/* SegmentDistSqrd distancefunctions.cpp:94-164 */
double ref_segment_dist_sqrd(const double *PA, const double *PB,
                             const double *QA, const double *QB) {
  int possible = 1;
  double m, diffA, diffB;
  if (fabs(PB[0] - PA[0]) < 0.000001) { /* :104-110 */
    if ((QA[0] >= PA[0] && QB[0] >= PA[0]) || (QA[0] <= PA[0] && QB[0] <= PA[0]))
      possible = 0;
  } else { /* :111-122 */
    m = (PB[1] - PA[1]) / (PB[0] - PA[0]);
    diffA = (m * (QA[0] - PA[0]) + PA[1]) - QA[1];
    diffB = (m * (QB[0] - PA[0]) + PA[1]) - QB[1];
    if ((diffA > 0.0 && diffB > 0.0) || (diffA < 0.0 && diffB < 0.0)) possible = 0;
  }
  if (possible) { /* :124-143 */
    if (fabs(QB[0] - QA[0]) < 0.000001) {
      if ((PA[0] >= QA[0] && PB[0] >= QA[0]) || (PA[0] <= QA[0] && PB[0] <= QA[0]))
        possible = 0;
    } else {
      m = (QB[1] - QA[1]) / (QB[0] - QA[0]);
      diffA = (m * (PA[0] - QA[0]) + QA[1]) - PA[1];
      diffB = (m * (PB[0] - QA[0]) + QA[1]) - PB[1];
      if ((diffA > 0.0 && diffB > 0.0) || (diffA < 0.0 && diffB < 0.0)) possible = 0;
    }
  }
  if (possible) return 0.0; /* :145-148 */
  double d0 = ref_dist_sqrd_point_segment(PA, QA, QB); /* :153-163 */
  double d1 = ref_dist_sqrd_point_segment(PB, QA, QB);
  double d2 = ref_dist_sqrd_point_segment(QA, PA, PB);
  double d3 = ref_dist_sqrd_point_segment(QB, PA, PB);
  double mn = d0;
  if (d1 < mn) mn = d1;
  if (d2 < mn) mn = d2;
  if (d3 < mn) mn = d3;
  return mn;
}

/* PointInPolygon distancefunctions.cpp:166-205 */
int ref_point_in_polygon(const double *pt, const double *poly, int nv) {
  if (nv < 2) return 0;
  int num_crossings = 0;
  const double *sp = poly + 2 * (nv - 1);
  for (int i = 0; i < nv; i++) {
    const double *ep = poly + 2 * i;
    if ((sp[1] > pt[1] && ep[1] < pt[1]) || (sp[1] < pt[1] && ep[1] > pt[1])) {
      if (sp[0] > pt[0] && ep[0] > pt[0]) {
        num_crossings++;
      } else if (sp[0] < pt[0] && ep[0] < pt[0]) {
      } else {
        double mx = (sp[0] < ep[0]) ? ep[0] : sp[0];
        double t = 2 * mx;
        double x = (-((sp[0] * ep[1] - sp[1] * ep[0]) * (pt[0] - t)) +
                    ((sp[0] - ep[0]) * (pt[0] * pt[1] - pt[1] * t))) /
                   ((sp[1] - ep[1]) * (pt[0] - t));
        if (x > pt[0]) num_crossings++;
      }
    }
    sp = ep;
  }
  return (num_crossings % 2 == 1) ? 1 : 0;
}
