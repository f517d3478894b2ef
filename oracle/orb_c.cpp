// Compiled steered-BRIEF descriptor of the CPU oracle (TEST INFRASTRUCTURE - see oracle/__init__.py; used by bench.py's CPU
// legs so that they time compiled code like the reference's ORBdescriptor class, and pinned bit for bit to oracle/orb.py by
// tests/test_cpu.py::test_compiled_orb_equals_the_numpy_restatement).
// Follows /root/reference/src/ORBDescriptor.cpp: IC_Angle :486-513 (intensity centroid over the radius-15 disc, cv::fastAtan2),
// computeOrbDescriptor :334-382 (256 rotated pair tests on the blurred image), Hamming distance ORBDescriptor.h:44-60.
// Build with -ffp-contract=off: every float product and sum is rounded separately, like the reference's non-FMA build.
#include <cmath>
#include <cstddef>
#include <cstdint>

namespace {
constexpr int HALF_PATCH = 15;

// cv::fastAtan2 (degrees): OpenCV's polynomial, scalar path
inline float fast_atan2_deg(float y, float x) {
  const float p1 = 57.283626556396484f, p3 = -18.66744613647461f, p5 = 8.914000511169434f, p7 = -2.539724588394165f;
  const float eps = 2.220446049250313e-16f;
  const float ax = std::fabs(x), ay = std::fabs(y);
  float a, c, c2;
  if (ax >= ay) { c = ay / (ax + eps); c2 = c * c; a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
  else { c = ax / (ay + eps); c2 = c * c; a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
  if (x < 0.f) a = 180.f - a;
  if (y < 0.f) a = 360.f - a;
  return a;
}
}  // namespace

extern "C" {

// raw / blur: the 32-px bordered images of oracle/orb.py (pitch bytes per row); pts [n][2] float32; pattern [256][4] int32
// (x1, y1, x2, y2); umax [16]; out [n][32]; angles_out optional [n]
void orbc_compute(const uint8_t* raw, const uint8_t* blur, int pitch, int border, const float* pts, int n, const int32_t* pattern,
                  const int32_t* umax, uint8_t* out, float* angles_out) {
  const float factor_pi = 0.017453292519943295f;           // (float)(CV_PI / 180.f)
  for (int i = 0; i < n; ++i) {
    const int cx = (int)std::nearbyintf(pts[2 * i]) + border, cy = (int)std::nearbyintf(pts[2 * i + 1]) + border;
    const uint8_t* c = raw + (size_t)cy * pitch + cx;
    long long m10 = 0, m01 = 0;
    for (int u = -HALF_PATCH; u <= HALF_PATCH; ++u) m10 += (long long)u * c[u];
    for (int v = 1; v <= HALF_PATCH; ++v) {
      const int d = umax[v];
      long long vsum = 0;
      for (int u = -d; u <= d; ++u) {
        const int plus = c[u + (std::ptrdiff_t)v * pitch], minus = c[u - (std::ptrdiff_t)v * pitch];
        vsum += plus - minus;
        m10 += (long long)u * (plus + minus);
      }
      m01 += (long long)v * vsum;
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);
    if (angles_out) angles_out[i] = angle;
    const float ang = angle * factor_pi;
    const float a = (float)std::cos((double)ang), b = (float)std::sin((double)ang);
    const uint8_t* cb = blur + (size_t)cy * pitch + cx;
    for (int byte = 0; byte < 32; ++byte) {
      unsigned val = 0;
      for (int k = 0; k < 8; ++k) {
        const int32_t* p = pattern + (size_t)(byte * 8 + k) * 4;
        int t[2];
        for (int e = 0; e < 2; ++e) {
          const float px = (float)p[2 * e], py = (float)p[2 * e + 1];
          const float x = px * a - py * b;
          const float y = px * b + py * a;
          const int ix = (int)std::nearbyintf(x), iy = (int)std::nearbyintf(y);
          t[e] = cb[(std::ptrdiff_t)iy * pitch + ix];
        }
        val |= (unsigned)(t[0] < t[1]) << k;
      }
      out[(size_t)i * 32 + byte] = (uint8_t)val;
    }
  }
}

// Hamming distances of descriptor rows a[i], b[i]
void orbc_hamming_rows(const uint8_t* a, const uint8_t* b, int n, int32_t* out) {
  for (int i = 0; i < n; ++i) {
    const uint64_t* pa = (const uint64_t*)(a + (size_t)i * 32);
    const uint64_t* pb = (const uint64_t*)(b + (size_t)i * 32);
    int d = 0;
    for (int k = 0; k < 4; ++k) d += __builtin_popcountll(pa[k] ^ pb[k]);
    out[i] = d;
  }
}

}  // extern "C"
