// Host-side on-disk formats of the replay path, C ABI, no OpenCV:
//   * 8-bit grey PNG decode (what cv::imread(path, 0) does for EuRoC's cam0 images, app/larvioMain.cpp:92-96):
//     zlib inflate + the five PNG scan-line filters; anything else (16-bit, colour, palette, interlaced) is refused;
//   * EuRoC ASL csv readers with the reference's parsing rules (include/utils/DataReader.hpp:31-120): first line is a
//     header, stamp = 1e-9 * integer ns, IMU columns w_x w_y w_z a_x a_y a_z.
// Built into liblarvio_io.so (links zlib) so that the CUDA library itself keeps no extra dependency.
#include <zlib.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/larvio_b200.h"

static std::string g_io_err;
extern "C" const char* lvbio_last_error(void) { return g_io_err.c_str(); }
static int io_fail(const std::string& m) { g_io_err = m; return -1; }

static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// Decodes `path` into out[h][w] (tightly packed).  Returns 0, or -1 with lvbio_last_error().
extern "C" int lvbio_png_read_gray8(const char* path, uint8_t* out, int cap_bytes, int* w_out, int* h_out) {
  if (!path || !w_out || !h_out) return io_fail("lvbio_png_read_gray8: null argument");
  FILE* f = std::fopen(path, "rb");
  if (!f) return io_fail(std::string("cannot open ") + path);
  std::vector<uint8_t> buf;
  uint8_t tmp[65536];
  size_t n;
  while ((n = std::fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
  std::fclose(f);
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (buf.size() < 33 || std::memcmp(buf.data(), sig, 8) != 0) return io_fail(std::string(path) + ": not a PNG file");
  size_t pos = 8;
  int w = 0, h = 0;
  bool have_hdr = false;
  std::vector<uint8_t> idat;
  while (pos + 12 <= buf.size()) {
    const uint32_t len = be32(&buf[pos]);
    const char* type = reinterpret_cast<const char*>(&buf[pos + 4]);
    if (pos + 12 + (size_t)len > buf.size()) return io_fail(std::string(path) + ": truncated chunk");
    const uint8_t* data = &buf[pos + 8];
    if (!std::memcmp(type, "IHDR", 4)) {
      if (len != 13) return io_fail(std::string(path) + ": bad IHDR");
      w = (int)be32(data); h = (int)be32(data + 4);
      if (data[8] != 8 || data[9] != 0 || data[10] != 0 || data[11] != 0 || data[12] != 0)
        return io_fail(std::string(path) + ": only 8-bit greyscale, non-interlaced PNG is supported");
      have_hdr = true;
    } else if (!std::memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), data, data + len);
    } else if (!std::memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + (size_t)len;
  }
  if (!have_hdr || w <= 0 || h <= 0 || idat.empty()) return io_fail(std::string(path) + ": missing IHDR/IDAT");
  *w_out = w; *h_out = h;
  if (!out) return 0;                                   // size query
  if ((long long)w * h > cap_bytes) return io_fail(std::string(path) + ": output buffer too small");
  std::vector<uint8_t> raw((size_t)(w + 1) * h);
  uLongf rawlen = (uLongf)raw.size();
  if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size())
    return io_fail(std::string(path) + ": zlib inflate failed");
  for (int y = 0; y < h; ++y) {
    const uint8_t* src = &raw[(size_t)y * (w + 1)];
    uint8_t* dst = out + (size_t)y * w;
    const uint8_t* up = y ? out + (size_t)(y - 1) * w : nullptr;
    const int ft = src[0];
    for (int x = 0; x < w; ++x) {
      const int a = x ? dst[x - 1] : 0, b = up ? up[x] : 0, c = (x && up) ? up[x - 1] : 0;
      int v = src[1 + x];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: {
          const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
          v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
          break;
        }
        default: return io_fail(std::string(path) + ": bad scan-line filter");
      }
      dst[x] = (uint8_t)v;
    }
  }
  return 0;
}

// loadImuFile: rows -> out[cap]; *n = rows found (may exceed cap: call again with a larger buffer).
extern "C" int lvbio_euroc_read_imu(const char* csv_path, LvbImu* out, int cap, int* n) {
  if (!csv_path || !n) return io_fail("lvbio_euroc_read_imu: null argument");
  FILE* f = std::fopen(csv_path, "r");
  if (!f) return io_fail(std::string("cannot open ") + csv_path);
  char line[1024];
  int cnt = 0;
  if (!std::fgets(line, sizeof(line), f)) { std::fclose(f); *n = 0; return 0; }          // header
  while (std::fgets(line, sizeof(line), f)) {
    if (line[0] == '\n' || line[0] == '\r' || line[0] == 0) continue;
    char* p = line;
    const long long ns = std::atoll(p);                                                   // atol(substr(0, comma))
    double v[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 0; j < 6; ++j) {
      p = std::strchr(p, ',');
      if (!p) break;
      ++p;
      v[j] = std::atof(p);
    }
    if (out && cnt < cap) {
      out[cnt].t = 1e-9 * (double)ns;
      for (int k = 0; k < 3; ++k) { out[cnt].gyro[k] = v[k]; out[cnt].acc[k] = v[3 + k]; }
    }
    ++cnt;
  }
  std::fclose(f);
  *n = cnt;
  return 0;
}

// loadImageList: stamps -> t[cap], file names -> names[cap][name_len] (NUL terminated, '\r' stripped).
extern "C" int lvbio_euroc_read_image_list(const char* csv_path, double* t, char* names, int name_len, int cap, int* n) {
  if (!csv_path || !n) return io_fail("lvbio_euroc_read_image_list: null argument");
  FILE* f = std::fopen(csv_path, "r");
  if (!f) return io_fail(std::string("cannot open ") + csv_path);
  char line[1024];
  int cnt = 0;
  if (!std::fgets(line, sizeof(line), f)) { std::fclose(f); *n = 0; return 0; }
  while (std::fgets(line, sizeof(line), f)) {
    if (line[0] == '\n' || line[0] == '\r' || line[0] == 0) continue;
    const long long ns = std::atoll(line);
    const char* c = std::strchr(line, ',');
    std::string name = c ? std::string(c + 1) : std::string();
    while (!name.empty() && (name.back() == '\n' || name.back() == '\r' || name.back() == ' ')) name.pop_back();
    const size_t c2 = name.find(',');
    if (c2 != std::string::npos) name.resize(c2);
    if (t && names && cnt < cap) {
      t[cnt] = 1e-9 * (double)ns;
      std::snprintf(names + (size_t)cnt * name_len, name_len, "%s", name.c_str());
    }
    ++cnt;
  }
  std::fclose(f);
  *n = cnt;
  return 0;
}

// findFirstAlign (include/utils/DataReader.hpp:123-165): first image / IMU sample pair with equal stamps.  Returns 0 and the
// two start indices, -1 if the streams never align.
extern "C" int lvbio_first_align(const double* t_img, int n_img, const LvbImu* imu, int n_imu, int* img0, int* imu0) {
  if (!t_img || !imu || n_img <= 0 || n_imu <= 0 || !img0 || !imu0) { g_io_err = "lvbio_first_align: bad argument"; return -1; }
  const double imu_t0 = imu[0].t, img_t0 = t_img[0];
  if (imu_t0 > img_t0) {
    for (int i = 1; i < n_img; ++i)
      if (imu_t0 <= t_img[i]) {
        for (int j = 0; j < n_imu; ++j)
          if (imu[j].t == t_img[i]) { *img0 = i; *imu0 = j; return 0; }
        break;
      }
  } else if (imu_t0 < img_t0) {
    for (int j = 1; j < n_imu; ++j)
      if (imu[j].t == img_t0) { *img0 = 0; *imu0 = j; return 0; }
  } else {
    *img0 = 0; *imu0 = 0;
    return 0;
  }
  g_io_err = "no image/IMU pair with equal stamps";
  return -1;
}
