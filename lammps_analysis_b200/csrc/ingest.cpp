// Native LAMMPS text-dump tokenizer (host code): the step *before* the hot path.
//
// Replaces the per-line Python parsing of the reference
//   mdsuite/file_io/lammps_trajectory_files.py:100-243   (header, box, columns, sample rate)
//   mdsuite/file_io/tabular_text_files.py:122-220        (readline().split() per atom, np.stack
//                                                         of strings, argsort by id)
// with one block read per frame batch and a pool of host threads that tokenise whole frames.
// Numbers are converted exactly as strtod / Python's float() do (exact fast path, strtod
// otherwise), non-numeric tokens (element symbols) become NaN, rows are stably sorted by the
// `id` column per frame -- so the arrays handed to the store are bit-identical to what the
// Python reader produces.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
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

// ---- frame-parallel reader ------------------------------------------------------------------
// One block read of the byte range that holds the n frames (newline count by memchr), then the
// frames are tokenised by a pool of host threads.  Numbers take Clinger's exact fast path
// (<= 15 significant digits and |decimal exponent| <= 22: the integer mantissa and the power of
// ten are both exact doubles, so one multiplication or division is the correctly rounded result,
// i.e. exactly what strtod / Python's float() return); anything else -- longer mantissas, "nan",
// "inf", hex floats -- falls back to strtod.  Non-numeric tokens (element symbols) become NaN.
namespace {

const double P10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                        1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

inline bool is_sep(char c) { return c == ' ' || c == '\t' || c == '\r'; }

// token [p, e) -> value; returns false when the token is not a number
bool parse_token(const char* p, const char* e, double* out) {
  const char* s = p;
  bool neg = false;
  if (s < e && (*s == '-' || *s == '+')) neg = (*s++ == '-');
  unsigned long long m = 0;
  int sig = 0, e10 = 0;
  bool digits = false, fast = true;
  for (; s < e && *s >= '0' && *s <= '9'; ++s) {
    digits = true;
    if (sig < 18) {
      m = m * 10 + (unsigned)(*s - '0');
      if (m) ++sig;
    } else {
      fast = false;
    }
  }
  if (s < e && *s == '.') {
    for (++s; s < e && *s >= '0' && *s <= '9'; ++s) {
      digits = true;
      if (sig < 18) {
        m = m * 10 + (unsigned)(*s - '0');
        if (m) ++sig;
        --e10;
      } else {
        fast = false;
      }
    }
  }
  if (digits && s < e && (*s == 'e' || *s == 'E')) {
    const char* t = s + 1;
    bool eneg = false;
    if (t < e && (*t == '-' || *t == '+')) eneg = (*t++ == '-');
    if (t < e && *t >= '0' && *t <= '9') {
      int ex = 0;
      for (; t < e && *t >= '0' && *t <= '9'; ++t)
        if (ex < 100000) ex = ex * 10 + (*t - '0');
      e10 += eneg ? -ex : ex;
      s = t;
    }
  }
  if (digits && s == e && fast && sig <= 15 && e10 >= -22 && e10 <= 22) {
    double v = (double)m;
    v = e10 < 0 ? v / P10[-e10] : v * P10[e10];
    *out = neg ? -v : v;
    return true;
  }
  // general case: strtod on a NUL-terminated copy (must consume the whole token)
  char tmp[64];
  const size_t len = (size_t)(e - p);
  std::string big;
  const char* z;
  if (len < sizeof(tmp)) {
    memcpy(tmp, p, len);
    tmp[len] = 0;
    z = tmp;
  } else {
    big.assign(p, len);
    z = big.c_str();
  }
  char* end = nullptr;
  const double v = strtod(z, &end);
  if (end == z || *end) return false;
  *out = v;
  return true;
}

struct FrameJob {
  const char* begin;  // first header line of the frame
  const char* end;    // one past the frame's last byte
};

// tokenises one frame into dst[n_atoms][n_cols]; returns nullptr or an error text
const char* parse_frame(const FrameJob& job, long long n_atoms, int n_cols, int id_col, int sorted,
                        double* dst, std::vector<double>& tab, std::vector<long long>& order) {
  const double nan = std::nan("");
  const char* p = job.begin;
  for (int i = 0; i < HEADER_LINES; ++i) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(job.end - p));
    if (!nl) return "unexpected end of file in frame header";
    p = nl + 1;
  }
  double* rows = sorted ? dst : tab.data();
  for (long long a = 0; a < n_atoms; ++a) {
    if (p >= job.end) return "unexpected end of file in frame body";
    const char* nl = (const char*)memchr(p, '\n', (size_t)(job.end - p));
    const char* le = nl ? nl : job.end;
    double* row = rows + (size_t)a * n_cols;
    for (int c = 0; c < n_cols; ++c) {
      while (p < le && is_sep(*p)) ++p;
      if (p >= le) return "row with too few columns";
      const char* t = p;
      while (t < le && !is_sep(*t)) ++t;
      if (!parse_token(p, t, &row[c])) row[c] = nan;  // e.g. an element symbol
      p = t;
    }
    p = nl ? nl + 1 : job.end;
  }
  if (!sorted) {
    std::iota(order.begin(), order.end(), 0ll);
    std::stable_sort(order.begin(), order.end(), [&](long long x, long long y) {
      return tab[(size_t)x * n_cols + id_col] < tab[(size_t)y * n_cols + id_col];
    });
    for (long long a = 0; a < n_atoms; ++a)
      memcpy(dst + (size_t)a * n_cols, tab.data() + (size_t)order[a] * n_cols,
             n_cols * sizeof(double));
  }
  return nullptr;
}

}  // namespace

// Reads `n` frames starting at byte *offset (0 for the first call; updated on return) into
// out[n][n_atoms][n_cols] (float64), rows stably sorted by column `id_col` unless `sorted`.
extern "C" int mdk_lammps_read(const char* path, long long n_atoms, int n_cols, int id_col,
                               int sorted, long long n, long long* offset, double* out) {
  if (!path || !offset || !out || n_atoms < 0 || n_cols < 1 || id_col < 0 || id_col >= n_cols ||
      n < 0) {
    mdk::set_error("lammps_read: bad argument");
    return MDK_EINVAL;
  }
  if (n == 0) return MDK_OK;
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
  // ---- block reads until the buffer holds n * (n_atoms + 9) lines -----------------------------
  const long long lines_per_frame = n_atoms + HEADER_LINES;
  std::vector<char> buf;
  std::vector<size_t> frame_start(1, 0);   // byte offsets of the frames inside buf
  size_t scanned = 0;
  long long lines_in_frame = 0;
  bool eof = false;
  const size_t block = (size_t)8 << 20;
  while ((long long)frame_start.size() <= n && !eof) {
    const size_t old = buf.size();
    buf.resize(old + block);
    const size_t got = fread(buf.data() + old, 1, block, fp);
    buf.resize(old + got);
    if (got < block) eof = true;
    while (scanned < buf.size() && (long long)frame_start.size() <= n) {
      const char* nl = (const char*)memchr(buf.data() + scanned, '\n', buf.size() - scanned);
      if (!nl) break;
      scanned = (size_t)(nl - buf.data()) + 1;
      if (++lines_in_frame == lines_per_frame) {
        lines_in_frame = 0;
        frame_start.push_back(scanned);
      }
    }
  }
  fclose(fp);
  if ((long long)frame_start.size() <= n) {
    // the file may end without a trailing newline: the last line then closes the last frame
    if (eof && (long long)frame_start.size() == n && lines_in_frame == lines_per_frame - 1 &&
        scanned < buf.size()) {
      frame_start.push_back(buf.size());
    } else {
      mdk::set_error("lammps_read: unexpected end of file (%lld of %lld frames)",
                     (long long)frame_start.size() - 1, n);
      return MDK_EINVAL;
    }
  }
  // ---- tokenise the frames on a pool of host threads ----------------------------------------------
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 1;
  if (const char* env = getenv("MDK_INGEST_THREADS")) hw = (unsigned)std::max(1, atoi(env));
  const unsigned n_threads = (unsigned)std::min<long long>(std::min<unsigned>(hw, 32u), n);
  std::vector<const char*> errors(n_threads, nullptr);
  auto worker = [&](unsigned tid) {
    std::vector<double> tab(sorted ? 0 : (size_t)n_atoms * n_cols);
    std::vector<long long> order(sorted ? 0 : (size_t)n_atoms);
    for (long long f = tid; f < n; f += n_threads) {
      const FrameJob job{buf.data() + frame_start[(size_t)f], buf.data() + frame_start[(size_t)f + 1]};
      const char* err = parse_frame(job, n_atoms, n_cols, id_col, sorted,
                                    out + (size_t)f * n_atoms * n_cols, tab, order);
      if (err) {
        errors[tid] = err;
        return;
      }
    }
  };
  if (n_threads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < n_threads; ++t) pool.emplace_back(worker, t);
    for (auto& th : pool) th.join();
  }
  for (const char* err : errors)
    if (err) {
      mdk::set_error("lammps_read: %s", err);
      return MDK_EINVAL;
    }
  *offset += (long long)frame_start[(size_t)n];
  return MDK_OK;
}
