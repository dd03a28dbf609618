// Native LAMMPS text-dump tokenizer (host code): the step *before* the hot path.
//
// Replaces the per-line Python parsing of the reference
//   mdsuite/file_io/lammps_trajectory_files.py:100-243   (header, box, columns, sample rate)
//   mdsuite/file_io/tabular_text_files.py:122-220        (readline().split() per atom, np.stack
//                                                         of strings, argsort by id)
// with one buffered pass per frame batch.  Numbers are converted with strtod (correctly rounded,
// identical to Python's float()), non-numeric tokens (element symbols) become NaN, rows are
// stably sorted by the `id` column per frame -- so the arrays handed to the store are bit-identical
// to what the Python reader produces.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/mdk.h"

namespace mdk {
void set_error(const char* fmt, ...);
}

namespace {

struct LineReader {
  FILE* fp;
  std::vector<char> buf;
  explicit LineReader(FILE* f) : fp(f), buf(1 << 16) {}
  // returns pointer to a NUL-terminated line (without '\n') or nullptr at EOF
  char* next() {
    size_t len = 0;
    for (;;) {
      if (!fgets(buf.data() + len, (int)(buf.size() - len), fp)) return len ? buf.data() : nullptr;
      len += strlen(buf.data() + len);
      if (len && buf[len - 1] == '\n') {
        buf[--len] = 0;
        if (len && buf[len - 1] == '\r') buf[--len] = 0;
        return buf.data();
      }
      if (feof(fp)) return buf.data();
      buf.resize(buf.size() * 2);
    }
  }
};

constexpr int HEADER_LINES = 9;

}  // namespace

// Scans the first header and counts the lines of the file.
//   n_atoms, n_frames        : out
//   steps[2]                 : TIMESTEP of the first two frames (steps[1] = steps[0] if only one)
//   box[6]                   : xlo xhi ylo yhi zlo zhi of the first frame
//   columns / columns_cap    : out, the space separated column names after "ITEM: ATOMS"
extern "C" int mdk_lammps_scan(const char* path, long long* n_atoms, long long* n_frames,
                               long long* steps, double* box, char* columns, int columns_cap) {
  if (!path || !n_atoms || !n_frames || !steps || !box || !columns || columns_cap < 2) {
    mdk::set_error("lammps_scan: bad argument");
    return MDK_EINVAL;
  }
  FILE* fp = fopen(path, "rb");
  if (!fp) {
    mdk::set_error("lammps_scan: cannot open %s", path);
    return MDK_EINVAL;
  }
  LineReader rd(fp);
  std::vector<std::string> hdr;
  for (int i = 0; i < HEADER_LINES; ++i) {
    char* l = rd.next();
    if (!l) {
      fclose(fp);
      mdk::set_error("lammps_scan: %s is shorter than one header", path);
      return MDK_EINVAL;
    }
    hdr.emplace_back(l);
  }
  *n_atoms = atoll(hdr[3].c_str());
  steps[0] = steps[1] = atoll(hdr[1].c_str());
  for (int d = 0; d < 3; ++d) {
    char* end = nullptr;
    box[2 * d] = strtod(hdr[5 + d].c_str(), &end);
    box[2 * d + 1] = strtod(end, nullptr);
  }
  const char* cols = strstr(hdr[8].c_str(), "ATOMS");
  cols = cols ? cols + 5 : hdr[8].c_str();
  while (*cols == ' ') ++cols;
  snprintf(columns, columns_cap, "%s", cols);
  if (*n_atoms < 0) {
    fclose(fp);
    mdk::set_error("lammps_scan: bad atom count");
    return MDK_EINVAL;
  }
  // second frame's TIMESTEP (sample rate, lammps_trajectory_files.py:228-242)
  long long line_no = HEADER_LINES;
  for (long long i = 0; i < *n_atoms; ++i, ++line_no)
    if (!rd.next()) break;
  if (rd.next()) {
    ++line_no;
    if (char* l = rd.next()) {
      ++line_no;
      steps[1] = atoll(l);
    }
  }
  // count the remaining lines with large reads
  long long lines = line_no;
  {
    std::vector<char> chunk(1 << 22);
    size_t got;
    char last = '\n';
    while ((got = fread(chunk.data(), 1, chunk.size(), fp)) > 0) {
      lines += std::count(chunk.begin(), chunk.begin() + got, '\n');
      last = chunk[got - 1];
    }
    if (last != '\n') ++lines;  // unterminated last line
  }
  fclose(fp);
  const long long per = *n_atoms + HEADER_LINES;
  if (lines % per != 0) {
    mdk::set_error("lammps_scan: %lld lines is not a multiple of (n_atoms + 9) = %lld", lines, per);
    return MDK_EINVAL;
  }
  *n_frames = lines / per;
  return MDK_OK;
}

// Reads `n` frames starting at byte *offset (0 for the first call; updated on return) into
// out[n][n_atoms][n_cols] (float64), rows stably sorted by column `id_col` unless `sorted`.
extern "C" int mdk_lammps_read(const char* path, long long n_atoms, int n_cols, int id_col,
                               int sorted, long long n, long long* offset, double* out) {
  if (!path || !offset || !out || n_atoms < 0 || n_cols < 1 || id_col < 0 || id_col >= n_cols) {
    mdk::set_error("lammps_read: bad argument");
    return MDK_EINVAL;
  }
  FILE* fp = fopen(path, "rb");
  if (!fp) {
    mdk::set_error("lammps_read: cannot open %s", path);
    return MDK_EINVAL;
  }
  if (fseeko(fp, (off_t)*offset, SEEK_SET) != 0) {
    fclose(fp);
    mdk::set_error("lammps_read: seek failed");
    return MDK_EINVAL;
  }
  setvbuf(fp, nullptr, _IOFBF, 1 << 22);
  LineReader rd(fp);
  std::vector<double> tab((size_t)n_atoms * n_cols);
  std::vector<long long> order((size_t)n_atoms);
  const double nan = std::nan("");
  for (long long f = 0; f < n; ++f) {
    for (int i = 0; i < HEADER_LINES; ++i)
      if (!rd.next()) {
        fclose(fp);
        mdk::set_error("lammps_read: unexpected end of file in frame header");
        return MDK_EINVAL;
      }
    for (long long a = 0; a < n_atoms; ++a) {
      char* p = rd.next();
      if (!p) {
        fclose(fp);
        mdk::set_error("lammps_read: unexpected end of file in frame body");
        return MDK_EINVAL;
      }
      double* row = tab.data() + (size_t)a * n_cols;
      for (int c = 0; c < n_cols; ++c) {
        while (*p == ' ' || *p == '\t') ++p;
        if (!*p) {
          fclose(fp);
          mdk::set_error("lammps_read: row with fewer than %d columns", n_cols);
          return MDK_EINVAL;
        }
        char* end = nullptr;
        const double v = strtod(p, &end);
        if (end == p || (*end && *end != ' ' && *end != '\t')) {
          row[c] = nan;  // not a number (e.g. an element symbol)
          while (*p && *p != ' ' && *p != '\t') ++p;
        } else {
          row[c] = v;
          p = end;
        }
      }
    }
    double* dst = out + (size_t)f * n_atoms * n_cols;
    if (sorted) {
      memcpy(dst, tab.data(), tab.size() * sizeof(double));
    } else {
      std::iota(order.begin(), order.end(), 0ll);
      std::stable_sort(order.begin(), order.end(), [&](long long x, long long y) {
        return tab[(size_t)x * n_cols + id_col] < tab[(size_t)y * n_cols + id_col];
      });
      for (long long a = 0; a < n_atoms; ++a)
        memcpy(dst + (size_t)a * n_cols, tab.data() + (size_t)order[a] * n_cols,
               n_cols * sizeof(double));
    }
  }
  *offset = (long long)ftello(fp);
  fclose(fp);
  return MDK_OK;
}
