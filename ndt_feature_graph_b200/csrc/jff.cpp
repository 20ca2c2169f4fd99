// jff.cpp — the on-disk LazyGrid map format of perception_oru ("JFF"), host code only.
//
// Replaces NDTMap::writeToJFF / NDTMap::loadFromJFF [upstream], reached from the reference at
// ndt_feature/src/ndt_feature_src/ndt_feature_fuser_hmt.cpp:15 (save), :24 and :39 (load).  The byte layout was decoded
// from the maps the reference ships (ndt_feature/data/FULL GRAPH/mapping{0..7}.jff):
//   "#JFF V0.50" (10 bytes) | int32 3 (= LazyGrid) | 9 x f64 {size_m[3], cell[3], center[3]} |
//   480 bytes: raw image of upstream's prototype NDTCell (pointers and all; meaningless outside the writing process) |
//   size_x*size_y*size_z records of 181 bytes, x-major / z-minor:
//     +0 4 x f32 cell centre (w = 1)   +16 3 x f64 cell size   +40 6 x f64 covariance xx,xy,xz,yy,yz,zz
//     +88 3 x f64 mean   +112 2 x f64 (0)   +128 i32 N   +132 i32 0   +136 i32 hasGaussian_   +140 12 bytes 0
//     +152 f32 log-odds occupancy   +156 u8 127   +157 4 x f32 1.0 (R,G,B + 1)   +173 8 bytes 0
// Records of cells without a Gaussian hold uninitialised memory in bytes +40..+127 upstream; the reader ignores them
// and the writer emits zeros.  Every byte this file models is reproduced exactly (tests/test_jff.py compares with the
// shipped maps); the prototype block is written as zeros.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/ndtb.h"

namespace {
constexpr size_t REC = 181, HDR = 10 + 4 + 72, PROTO = 480;

template <class T>
void put(unsigned char *p, T v) { std::memcpy(p, &v, sizeof v); }
template <class T>
T get(const unsigned char *p) { T v; std::memcpy(&v, p, sizeof v); return v; }
}  // namespace

extern "C" {

int ndtb_jff_write_cells(const char *path, const ndtb_grid *g, const ndtb_cell *cells, int64_t n) {
  if (!path || !g || n < 0 || (n > 0 && !cells)) return NDTB_ERR_ARG;
  const int64_t total = (int64_t)g->size[0] * g->size[1] * g->size[2];
  if (total <= 0) return NDTB_ERR_GRID;
  std::vector<int64_t> where((size_t)total, -1);
  for (int64_t i = 0; i < n; i++) {
    const int32_t *x = cells[i].idx;
    if (x[0] < 0 || x[1] < 0 || x[2] < 0 || x[0] >= g->size[0] || x[1] >= g->size[1] || x[2] >= g->size[2]) return NDTB_ERR_ARG;
    where[(size_t)(((int64_t)x[0] * g->size[1] + x[1]) * g->size[2] + x[2])] = i;
  }
  FILE *f = std::fopen(path, "wb");
  if (!f) return NDTB_ERR_ARG;
  unsigned char hdr[HDR + PROTO];
  std::memset(hdr, 0, sizeof hdr);
  std::memcpy(hdr, "#JFF V0.50", 10);
  put<int32_t>(hdr + 10, 3);
  for (int a = 0; a < 3; a++) {
    put<double>(hdr + 14 + 8 * a, g->size[a] * g->cell[a]);  // sizeXmeters: LazyGrid::setSize keeps ceil(size_m / cell)
    put<double>(hdr + 38 + 8 * a, g->cell[a]);
    put<double>(hdr + 62 + 8 * a, g->center[a]);
  }
  bool ok = std::fwrite(hdr, 1, sizeof hdr, f) == sizeof hdr;
  std::vector<unsigned char> buf(REC * (size_t)g->size[2] * (size_t)g->size[1]);
  for (int ix = 0; ix < g->size[0] && ok; ix++) {
    std::memset(buf.data(), 0, buf.size());
    for (int iy = 0; iy < g->size[1]; iy++)
      for (int iz = 0; iz < g->size[2]; iz++) {
        unsigned char *r = buf.data() + REC * ((size_t)iy * g->size[2] + iz);
        const int idx[3] = {ix, iy, iz};
        for (int a = 0; a < 3; a++) {
          put<float>(r + 4 * a, (float)(g->center[a] + (idx[a] - (int)(g->size[a] / 2.0)) * g->cell[a]));
          put<double>(r + 16 + 8 * a, g->cell[a]);
        }
        put<float>(r + 12, 1.f);
        r[156] = 127;
        for (int q = 0; q < 4; q++) put<float>(r + 157 + 4 * q, 1.f);
        const int64_t w = where[(size_t)(((int64_t)ix * g->size[1] + iy) * g->size[2] + iz)];
        if (w < 0) continue;
        const ndtb_cell &c = cells[w];
        if (c.has_gaussian) {
          for (int q = 0; q < 6; q++) put<double>(r + 40 + 8 * q, c.cov[q]);
          for (int q = 0; q < 3; q++) put<double>(r + 88 + 8 * q, c.mean[q]);
        }
        put<int32_t>(r + 128, c.n);
        put<int32_t>(r + 136, c.has_gaussian ? 1 : 0);
        put<float>(r + 152, c.occ);
      }
    ok = std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
  }
  ok = (std::fclose(f) == 0) && ok;
  return ok ? NDTB_OK : NDTB_ERR_ARG;
}

// Reads the grid and every cell that carries information (hasGaussian_, N > 0 or a non-zero occupancy).
// cells == NULL or cap too small: only *n (and *g) are returned, so a caller can size its buffer with a first call.
int ndtb_jff_read_cells(const char *path, ndtb_grid *g, ndtb_cell *cells, int64_t cap, int64_t *n) {
  if (!path || !g || !n) return NDTB_ERR_ARG;
  FILE *f = std::fopen(path, "rb");
  if (!f) return NDTB_ERR_ARG;
  unsigned char hdr[HDR + PROTO];
  if (std::fread(hdr, 1, sizeof hdr, f) != sizeof hdr || std::memcmp(hdr, "#JFF V0.50", 10) != 0 || get<int32_t>(hdr + 10) != 3) {
    std::fclose(f);
    return NDTB_ERR_ARG;
  }
  for (int a = 0; a < 3; a++) {
    const double size_m = get<double>(hdr + 14 + 8 * a);
    g->cell[a] = get<double>(hdr + 38 + 8 * a);
    g->center[a] = get<double>(hdr + 62 + 8 * a);
    if (!(g->cell[a] > 0)) {
      std::fclose(f);
      return NDTB_ERR_GRID;
    }
    const double cells_axis = std::fabs(std::ceil(size_m / g->cell[a]));  // LazyGrid::setSize
    if (!(cells_axis >= 1.0 && cells_axis <= 1048576.0) || !std::isfinite(g->center[a])) {  // also rejects NaN / inf headers
      std::fclose(f);
      return NDTB_ERR_GRID;
    }
    g->size[a] = (int32_t)cells_axis;
  }
  const int64_t total = (int64_t)g->size[0] * g->size[1] * g->size[2];  // <= 2^60: no overflow
  {  // a truncated file is refused before anything is read
    const long here = std::ftell(f);
    if (here < 0 || std::fseek(f, 0, SEEK_END) != 0) {
      std::fclose(f);
      return NDTB_ERR_ARG;
    }
    const long end = std::ftell(f);
    if (total > ((int64_t)1 << 31) || (int64_t)(end - here) < total * (int64_t)REC || std::fseek(f, here, SEEK_SET) != 0) {
      std::fclose(f);
      return NDTB_ERR_GRID;
    }
  }
  std::vector<unsigned char> buf(REC * 4096);
  int64_t found = 0, done = 0;
  bool ok = total > 0;
  while (ok && done < total) {
    const size_t want = (size_t)std::min<int64_t>(4096, total - done);
    if (std::fread(buf.data(), REC, want, f) != want) {
      ok = false;
      break;
    }
    for (size_t k = 0; k < want; k++) {
      const unsigned char *r = buf.data() + REC * k;
      const int32_t N = get<int32_t>(r + 128), has = get<int32_t>(r + 136);
      const float occ = get<float>(r + 152);
      if (!(has == 1 || N > 0 || occ != 0.f)) continue;
      if (cells && found < cap) {
        ndtb_cell &c = cells[found];
        std::memset(&c, 0, sizeof c);
        const int64_t lin = done + (int64_t)k;
        c.idx[2] = (int32_t)(lin % g->size[2]);
        c.idx[1] = (int32_t)((lin / g->size[2]) % g->size[1]);
        c.idx[0] = (int32_t)(lin / ((int64_t)g->size[2] * g->size[1]));
        c.n = N, c.has_gaussian = has == 1, c.occ = occ;
        if (has == 1) {
          for (int q = 0; q < 6; q++) c.cov[q] = get<double>(r + 40 + 8 * q);
          for (int q = 0; q < 3; q++) c.mean[q] = get<double>(r + 88 + 8 * q);
        }
      }
      found++;
    }
    done += (int64_t)want;
  }
  std::fclose(f);
  *n = found;
  if (!ok) return NDTB_ERR_ARG;
  return NDTB_OK;
}

}  // extern "C"
