// Device helpers shared by several front-end translation units.
#pragma once
#include "lvb_internal.h"

// cv::undistortPoints (radtan: 5 fixed iterations in double) and
// cv::fisheye::undistortPoints (equidistant), P = K when to_pixels else identity; float output.
// Call sites: image_processor.cpp:1040-1072.
__device__ __forceinline__ float2 lvb_undistort_point(const LvbCamera& c, float2 p, int to_pixels) {
  double x = ((double)p.x - c.cx) / c.fx, y = ((double)p.y - c.cy) / c.fy;   // cv: (x - cx)*ifx with ifx = 1/fx
  x = ((double)p.x - c.cx) * (1.0 / c.fx);
  y = ((double)p.y - c.cy) * (1.0 / c.fy);
  if (c.model == 0) {
    const double k1 = c.dist[0], k2 = c.dist[1], p1 = c.dist[2], p2 = c.dist[3];
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {
      const double r2 = x * x + y * y;
      const double icdist = 1.0 / (1 + ((0.0 * r2 + k2) * r2 + k1) * r2);
      const double dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
      const double dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
      x = (x0 - dx) * icdist;
      y = (y0 - dy) * icdist;
    }
  } else {
    const double k0 = c.dist[0], k1 = c.dist[1], k2 = c.dist[2], k3 = c.dist[3];
    double theta_d = sqrt(x * x + y * y);
    theta_d = fmin(fmax(-1.5707963267948966, theta_d), 1.5707963267948966);
    double scale = 0.0;
    if (theta_d > 1e-8) {
      double theta = theta_d;
      for (int j = 0; j < 10; ++j) {
        const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
        const double k0t2 = k0 * t2, k1t4 = k1 * t4, k2t6 = k2 * t6, k3t8 = k3 * t8;
        const double fix = (theta * (1 + k0t2 + k1t4 + k2t6 + k3t8) - theta_d) /
                           (1 + 3 * k0t2 + 5 * k1t4 + 7 * k2t6 + 9 * k3t8);
        theta = theta - fix;
        if (fabs(fix) < 1e-12) break;
      }
      scale = tan(theta) / theta_d;
    }
    x *= scale; y *= scale;
  }
  if (to_pixels) { x = x * c.fx + c.cx; y = y * c.fy + c.cy; }
  return make_float2((float)x, (float)y);
}

