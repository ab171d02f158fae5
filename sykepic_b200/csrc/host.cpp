// Host-side entry points of the C ABI: context-free utilities (.adc parsing, geometry
// validation, .prob.csv formatting, threshold quantisation).  No CUDA here.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <exception>

#include "spk_internal.h"

namespace spk {

constexpr int kStampCapHost = 8192;  // == tc::kStampCap (tc_common.cuh)

std::string& tls_error() {
  static thread_local std::string e;
  return e;
}

int fail(spk_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx)
    ctx->error = buf;
  else
    tls_error() = buf;
  return code;
}

ProfScope::ProfScope(spk_ctx* c, int category, double flops, double bytes, const char* fmt, ...) : ctx(c) {
  if (!c || !c->profiling) return;
  ProfRec r;
  r.category = category;
  r.flops = flops;
  r.bytes = bytes;
  char buf[256];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  r.what = buf;
  r.slot = -1;
  r.start = r.stop = nullptr;
  if (c->prof_mode == SPK_PROFILE_STAMPS) {
    if (!c->d_stamps || (int)c->prof.size() >= kStampCapHost) return;
    r.slot = (int)c->prof.size();
    c->cur_stamp = c->d_stamps + r.slot;
    idx = r.slot;
    c->prof.push_back(r);
    return;
  }
  if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) return;
  cudaEventRecord(r.start, c->stream);
  idx = (int)c->prof.size();
  c->prof.push_back(r);
}

ProfScope::~ProfScope() {
  if (idx < 0) return;
  if (ctx->prof[(size_t)idx].slot >= 0)
    ctx->cur_stamp = nullptr;
  else
    cudaEventRecord(ctx->prof[(size_t)idx].stop, ctx->stream);
}

namespace {

inline bool py_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r') || (c >= 0x1c && c <= 0x1f); }

// Python int(str) for the ASCII subset: [ws] [+-] digit (_? digit)* [ws].  false on anything else.
bool py_int(const char* b, const char* e, int64_t* out) {
  while (b < e && py_space(*b)) ++b;
  while (e > b && py_space(e[-1])) --e;
  if (b >= e) return false;
  bool neg = false;
  if (*b == '+' || *b == '-') {
    neg = (*b == '-');
    ++b;
  }
  if (b >= e || *b < '0' || *b > '9') return false;
  unsigned long long v = 0;
  bool prev_us = false;
  for (; b < e; ++b) {
    if (*b == '_') {
      if (prev_us) return false;
      prev_us = true;
      continue;
    }
    if (*b < '0' || *b > '9') return false;
    prev_us = false;
    if (v > (ULLONG_MAX - 9) / 10) return false;
    v = v * 10 + (unsigned)(*b - '0');
  }
  if (prev_us) return false;
  if (v > (unsigned long long)LLONG_MAX) return false;
  *out = neg ? -(int64_t)v : (int64_t)v;
  return true;
}

}  // namespace
}  // namespace spk

using namespace spk;

extern "C" {

int spk_abi_version(void) { return SPK_ABI_VERSION; }

static void prof_clear(spk_ctx* ctx) {
  for (auto& r : ctx->prof) {
    if (r.start) cudaEventDestroy(r.start);
    if (r.stop) cudaEventDestroy(r.stop);
  }
  ctx->prof.clear();
  ctx->cur_stamp = nullptr;
}

static int stamps_reset(spk_ctx* ctx) {
  if (!ctx->d_stamps) SPK_CUDA_OK(ctx, cudaMalloc(&ctx->d_stamps, 2 * (size_t)kStampCapHost * sizeof(unsigned long long)));
  SPK_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_stamps, 0xFF, (size_t)kStampCapHost * sizeof(unsigned long long), ctx->stream));
  SPK_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_stamps + kStampCapHost, 0, (size_t)kStampCapHost * sizeof(unsigned long long), ctx->stream));
  return SPK_OK;
}

int spk_profile_mode(spk_ctx* ctx, int mode) {
  if (!ctx || (mode != SPK_PROFILE_EVENTS && mode != SPK_PROFILE_STAMPS)) return fail(ctx, SPK_ERR_INVALID, "spk_profile_mode: bad arguments");
  if (ctx->profiling) return fail(ctx, SPK_ERR_STATE, "spk_profile_mode: a profiling pass is running");
  ctx->prof_mode = mode;
  return SPK_OK;
}

int spk_profile_begin(spk_ctx* ctx) {
  if (!ctx) return fail(nullptr, SPK_ERR_INVALID, "spk_profile_begin: null context");
  prof_clear(ctx);
  if (ctx->prof_mode == SPK_PROFILE_STAMPS) {
    SPK_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    int rc = stamps_reset(ctx);
    if (rc) return rc;
  }
  ctx->profiling = true;
  return SPK_OK;
}

int spk_profile_end(spk_ctx* ctx) {
  if (!ctx) return fail(nullptr, SPK_ERR_INVALID, "spk_profile_end: null context");
  cudaStreamSynchronize(ctx->stream);
  prof_clear(ctx);
  ctx->profiling = false;
  return SPK_OK;
}

int spk_profile_read(spk_ctx* ctx, double ms[SPK_PROF_CATEGORIES], double flops[SPK_PROF_CATEGORIES],
                     double bytes[SPK_PROF_CATEGORIES], int64_t launches[SPK_PROF_CATEGORIES], char* detail,
                     int64_t detail_cap) {
  if (!ctx || !ms || !flops || !bytes || !launches) return fail(ctx, SPK_ERR_INVALID, "spk_profile_read: bad arguments");
  SPK_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < SPK_PROF_CATEGORIES; ++i) {
    ms[i] = flops[i] = bytes[i] = 0.0;
    launches[i] = 0;
  }
  int64_t pos = 0;
  if (detail && detail_cap > 0) detail[0] = 0;
  std::vector<unsigned long long> st;
  if (ctx->prof_mode == SPK_PROFILE_STAMPS && ctx->d_stamps && !ctx->prof.empty()) {
    st.resize(2 * (size_t)kStampCapHost);
    SPK_CUDA_OK(ctx, cudaMemcpy(st.data(), ctx->d_stamps, st.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  }
  unsigned long long prev_end = 0;
  for (auto& r : ctx->prof) {
    float t = 0.f;
    char span[48] = "";
    if (r.slot >= 0) {
      const unsigned long long t0 = st[(size_t)r.slot], t1 = st[(size_t)kStampCapHost + (size_t)r.slot];
      if (t1 == 0 || t0 == ~0ull) continue;  // a kernel without stamps: its time goes to the next stamped launch
      const unsigned long long from = prev_end > t0 ? prev_end : t0;
      t = t1 > from ? (float)((double)(t1 - from) * 1e-6) : 0.f;
      snprintf(span, sizeof span, " |span_ms=%.6f", (double)(t1 - t0) * 1e-6);
      prev_end = t1 > prev_end ? t1 : prev_end;
    } else {
      SPK_CUDA_OK(ctx, cudaEventElapsedTime(&t, r.start, r.stop));
    }
    const int c = (r.category >= 0 && r.category < SPK_PROF_CATEGORIES) ? r.category : SPK_PROF_CATEGORIES - 1;
    ms[c] += t;
    flops[c] += r.flops;
    bytes[c] += r.bytes;
    launches[c] += 1;
    if (detail && pos < detail_cap - 1) {
      int m = snprintf(detail + pos, (size_t)(detail_cap - pos), "%d %.6f %.6g %.6g %s%s\n", c, (double)t, r.flops, r.bytes,
                       r.what.c_str(), span);
      if (m > 0) pos += (m < detail_cap - pos) ? m : (detail_cap - pos - 1);
    }
  }
  prof_clear(ctx);
  if (ctx->prof_mode == SPK_PROFILE_STAMPS && ctx->profiling) {
    int rc = stamps_reset(ctx);
    if (rc) return rc;
  }
  return SPK_OK;
}

const char* spk_last_error(const spk_ctx* ctx) { return ctx ? ctx->error.c_str() : tls_error().c_str(); }

void spk_new_dims(int h, int w, int target_h, int target_w, int* new_h, int* new_w) {
  new_dims(h, w, target_h, target_w, new_h, new_w);
}

int spk_adc_parse(const char* text, int64_t len, int64_t cap, int32_t* roi_id, int32_t* width, int32_t* height,
                  int64_t* start, int64_t* n_out, int64_t* n_lines) {
  if (!text || len < 0 || !n_out) return fail(nullptr, SPK_ERR_INVALID, "spk_adc_parse: bad arguments");
  int64_t n = 0, line_no = 0;
  const char* p = text;
  const char* end = text + len;
  while (p < end) {
    const char* eol = p;
    while (eol < end && *eol != '\n' && *eol != '\r') ++eol;
    ++line_no;
    // fields 15, 16, 17 of the comma-split line
    const char* fb[3] = {nullptr, nullptr, nullptr};
    const char* fe[3] = {nullptr, nullptr, nullptr};
    int field = 0;
    const char* q = p;
    const char* fstart = p;
    for (;; ++q) {
      if (q == eol || *q == ',') {
        if (field >= 15 && field <= 17) {
          fb[field - 15] = fstart;
          fe[field - 15] = q;
        }
        ++field;
        fstart = q + 1;
        if (q == eol || field > 17) break;
      }
    }
    if (field < 18)
      return fail(nullptr, SPK_ERR_PARSE, ".adc line %lld has %d fields, need at least 18", (long long)line_no, field);
    int64_t v[3];
    for (int i = 0; i < 3; ++i)
      if (!py_int(fb[i], fe[i], &v[i]))
        return fail(nullptr, SPK_ERR_PARSE, ".adc line %lld: field %d is not an integer", (long long)line_no, 15 + i);
    if (v[0] >= 1 && v[1] >= 1) {
      if (v[0] > INT32_MAX || v[1] > INT32_MAX)
        return fail(nullptr, SPK_ERR_PARSE, ".adc line %lld: ROI size out of range", (long long)line_no);
      if (n >= cap) return fail(nullptr, SPK_ERR_CAPACITY, "spk_adc_parse: more than %lld ROIs", (long long)cap);
      if (roi_id) roi_id[n] = (int32_t)line_no;
      if (width) width[n] = (int32_t)v[0];
      if (height) height[n] = (int32_t)v[1];
      if (start) start[n] = v[2];
      ++n;
    }
    // universal newlines: \r\n counts once
    p = eol;
    if (p < end) {
      if (*p == '\r' && p + 1 < end && p[1] == '\n')
        p += 2;
      else
        p += 1;
    }
  }
  *n_out = n;
  if (n_lines) *n_lines = line_no;
  return SPK_OK;
}

int spk_rois_validate(const int32_t* width, const int32_t* height, const int64_t* start, int64_t n, int64_t roi_len,
                      int target_h, int target_w, int64_t* first_bad) {
  if (first_bad) *first_bad = -1;
  for (int64_t i = 0; i < n; ++i) {
    int64_t w = width[i], h = height[i], s = start[i];
    // numpy slicing: a negative start counts from the end; the slice is never longer than asked
    int64_t lo = s < 0 ? (s + roi_len < 0 ? 0 : s + roi_len) : (s > roi_len ? roi_len : s);
    int64_t stop = s + w * h;
    int64_t hi = stop < 0 ? (stop + roi_len < 0 ? 0 : stop + roi_len) : (stop > roi_len ? roi_len : stop);
    int64_t got = hi > lo ? hi - lo : 0;
    if (w < 1 || h < 1 || got != w * h || s < 0) {
      if (first_bad) *first_bad = i;
      return fail(nullptr, SPK_ERR_FAULTY_BIN, "ROI %lld (%lldx%lld @%lld) runs past the %lld .roi bytes",
                  (long long)i, (long long)w, (long long)h, (long long)s, (long long)roi_len);
    }
  }
  for (int64_t i = 0; i < n; ++i) {
    int nh, nw;
    new_dims(height[i], width[i], target_h, target_w, &nh, &nw);
    if (nh < 1 || nw < 1) {
      if (first_bad) *first_bad = i;
      return fail(nullptr, SPK_ERR_EMPTY_RESIZE, "ROI %lld (%dx%d) resizes to %dx%d", (long long)i, width[i], height[i],
                  nw, nh);
    }
  }
  return SPK_OK;
}

int32_t spk_threshold_quantize(double thr, int strict) {
  // value(q) = the double nearest to q/1e5 (what float("%.5f") parses to); monotone in q.
  if (std::isnan(thr)) return INT32_MAX;
  auto value = [](int64_t q) { return (double)q / 100000.0; };  // correctly rounded division == strtod of the decimal
  double guess = std::ceil(thr * 100000.0);
  if (guess > 2.0e9) return INT32_MAX;
  if (guess < -2.0e9) return INT32_MIN;
  int64_t q = (int64_t)guess;
  auto ok = [&](int64_t k) { return strict ? value(k) > thr : value(k) >= thr; };
  while (ok(q - 1)) --q;
  while (!ok(q)) ++q;
  if (q > INT32_MAX) return INT32_MAX;
  if (q < INT32_MIN) return INT32_MIN;
  return (int32_t)q;
}

int spk_format_prob_csv(const char* header_line, const int32_t* roi_id, const float* probs, int64_t n, int k, char* out,
                        int64_t cap, int64_t* len) {
  if (!header_line || !len || n < 0 || k < 0) return fail(nullptr, SPK_ERR_INVALID, "spk_format_prob_csv: bad arguments");
  int64_t pos = 0;
  auto put = [&](const char* s, int64_t m) {
    if (out && pos + m <= cap) memcpy(out + pos, s, (size_t)m);
    pos += m;
  };
  put(header_line, (int64_t)strlen(header_line));
  char tmp[64];
  for (int64_t i = 0; i < n; ++i) {
    int m = snprintf(tmp, sizeof tmp, "%d", roi_id[i]);
    put(tmp, m);
    const float* row = probs + i * (int64_t)k;
    for (int j = 0; j < k; ++j) {
      float p = row[j];
      if (p >= 0.0f && p <= 1.0f) {
        // p * 1e5 is exact in double (24-bit x 17-bit); rint = round-half-even of the exact value,
        // which is what a correctly rounded "%.5f" prints.
        unsigned q = (unsigned)std::nearbyint((double)p * 100000.0);
        char* t = tmp;
        *t++ = ',';
        *t++ = (char)('0' + q / 100000u);
        unsigned f = q % 100000u;
        *t++ = '.';
        t[4] = (char)('0' + f % 10u); f /= 10u;
        t[3] = (char)('0' + f % 10u); f /= 10u;
        t[2] = (char)('0' + f % 10u); f /= 10u;
        t[1] = (char)('0' + f % 10u); f /= 10u;
        t[0] = (char)('0' + f % 10u);
        put(tmp, 8);
      } else {
        m = snprintf(tmp, sizeof tmp, ",%.5f", (double)p);  // nan / out of range: defer to libc
        put(tmp, m);
      }
    }
    put("\n", 1);
  }
  *len = pos;
  if (out && pos > cap) return fail(nullptr, SPK_ERR_CAPACITY, "spk_format_prob_csv: need %lld bytes", (long long)pos);
  return SPK_OK;
}

// ---- whole-file helpers for the host pipeline (no interpreter lock is held while these run)
namespace {
struct Fd {
  int fd = -1;
  ~Fd() {
    if (fd >= 0) close(fd);
  }
};
// reads exactly `want` bytes (the caller sized the buffer from fstat); false on error / short file
bool read_all(int fd, uint8_t* dst, int64_t want) {
  int64_t got = 0;
  while (got < want) {
    const ssize_t r = read(fd, dst + got, (size_t)std::min<int64_t>(want - got, (int64_t)1 << 30));
    if (r < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    if (r == 0) return false;
    got += r;
  }
  return true;
}
}  // namespace

int spk_bin_load(const char* adc_path, const char* roi_path, uint8_t* roi_buf, int64_t roi_cap, int64_t desc_cap,
                 int32_t* roi_id, int32_t* width, int32_t* height, int64_t* start, int64_t* n_out, int64_t* roi_len,
                 int target_h, int target_w) try {
  if (!adc_path || !roi_path || !n_out || !roi_len || roi_cap < 0 || desc_cap < 0)
    return fail(nullptr, SPK_ERR_INVALID, "spk_bin_load: bad arguments");
  *n_out = 0;
  *roi_len = 0;
  // ---- .adc: read whole, parse
  std::vector<char> adc;
  {
    Fd f;
    f.fd = open(adc_path, O_RDONLY | O_CLOEXEC);
    if (f.fd < 0) return fail(nullptr, SPK_ERR_IO, "%s: %s", adc_path, strerror(errno));
    struct stat st;
    if (fstat(f.fd, &st) != 0) return fail(nullptr, SPK_ERR_IO, "%s: %s", adc_path, strerror(errno));
    adc.resize((size_t)st.st_size);
    if (st.st_size > 0 && !read_all(f.fd, reinterpret_cast<uint8_t*>(adc.data()), (int64_t)st.st_size))
      return fail(nullptr, SPK_ERR_IO, "%s: short read", adc_path);
  }
  int64_t n = 0, lines = 0;
  int rc = spk_adc_parse(adc.empty() ? "" : adc.data(), (int64_t)adc.size(), desc_cap, roi_id, width, height, start, &n, &lines);
  if (rc != SPK_OK) return rc;
  // ---- .roi: straight into the caller's (pinned) buffer
  Fd f;
  f.fd = open(roi_path, O_RDONLY | O_CLOEXEC);
  if (f.fd < 0) return fail(nullptr, SPK_ERR_IO, "%s: %s", roi_path, strerror(errno));
  struct stat st;
  if (fstat(f.fd, &st) != 0) return fail(nullptr, SPK_ERR_IO, "%s: %s", roi_path, strerror(errno));
  *roi_len = (int64_t)st.st_size;
  if ((int64_t)st.st_size > roi_cap)
    return fail(nullptr, SPK_ERR_CAPACITY, "spk_bin_load: %s has %lld bytes, the buffer %lld", roi_path, (long long)st.st_size,
                (long long)roi_cap);
#ifdef POSIX_FADV_SEQUENTIAL
  posix_fadvise(f.fd, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
  if (st.st_size > 0) {
    if (!roi_buf) return fail(nullptr, SPK_ERR_INVALID, "spk_bin_load: null .roi buffer");
    if (!read_all(f.fd, roi_buf, (int64_t)st.st_size)) return fail(nullptr, SPK_ERR_IO, "%s: short read", roi_path);
  }
  *n_out = n;
  // ---- geometry: "Faulty raw data" / zero-pixel resize, as the reference finds out while decoding
  int64_t bad = -1;
  return spk_rois_validate(width, height, start, n, (int64_t)st.st_size, target_h, target_w, &bad);
} catch (const std::exception& e) {
  return fail(nullptr, SPK_ERR_STATE, "spk_bin_load: %s", e.what());
}

int spk_prob_csv_write(const char* path, const char* header_line, const int32_t* roi_id, const float* probs, int64_t n, int k,
                       int64_t* bytes_written) try {
  if (!path || !header_line || n < 0 || k < 0) return fail(nullptr, SPK_ERR_INVALID, "spk_prob_csv_write: bad arguments");
  const int64_t cap = (int64_t)strlen(header_line) + n * (12 + 8 * (int64_t)k + 1) + 16;
  std::vector<char> text((size_t)cap);
  int64_t len = 0;
  int rc = spk_format_prob_csv(header_line, roi_id, probs, n, k, text.data(), cap, &len);
  if (rc != SPK_OK) return rc;
  Fd f;
  f.fd = open(path, O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0666);
  if (f.fd < 0) return fail(nullptr, SPK_ERR_IO, "%s: %s", path, strerror(errno));
  int64_t put = 0;
  while (put < len) {
    const ssize_t r = write(f.fd, text.data() + put, (size_t)(len - put));
    if (r < 0) {
      if (errno == EINTR) continue;
      return fail(nullptr, SPK_ERR_IO, "%s: %s", path, strerror(errno));
    }
    put += r;
  }
  if (bytes_written) *bytes_written = len;
  return SPK_OK;
} catch (const std::exception& e) {
  return fail(nullptr, SPK_ERR_STATE, "spk_prob_csv_write: %s", e.what());
}

namespace {
// one line of text [b, e) without its line terminator; false at the end of the text
inline bool next_line(const char*& cur, const char* end, const char*& b, const char*& e) {
  if (cur >= end) return false;
  b = cur;
  const char* nl = (const char*)memchr(cur, '\n', (size_t)(end - cur));
  e = nl ? nl : end;
  cur = nl ? nl + 1 : end;
  if (e > b && e[-1] == '\r') --e;
  return true;
}
const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
// [sign] digits [. digits] with at most 15 significant digits: mantissa and power of ten are both exact doubles, so one
// division is the correctly rounded value (what pandas' default parser returns for such text, checked for all 100001
// five-decimal values in tests/test_host_probcsv.py).  Nothing else is accepted.
bool parse_decimal(const char* b, const char* e, double* out) {
  const char* p = b;
  bool neg = false;
  if (p < e && (*p == '-' || *p == '+')) neg = (*p++ == '-');
  uint64_t mant = 0;
  int digits = 0, frac = 0;
  bool dot = false, any = false;
  for (; p < e; ++p) {
    if (*p >= '0' && *p <= '9') {
      any = true;
      if (mant || *p != '0') ++digits;
      if (digits > 15) break;
      mant = mant * 10 + (uint64_t)(*p - '0');
      if (dot) ++frac;
    } else if (*p == '.' && !dot) {
      dot = true;
    } else {
      break;
    }
  }
  if (p == e && any && frac <= 22) {
    const double v = (double)mant / kPow10[frac];
    *out = neg ? -v : v;
    return true;
  }
  return false;  // exponents, inf / nan, longer mantissas: not what the writers produce -- the caller lets pandas decide
}
}  // namespace

int spk_prob_csv_shape(const char* text, int64_t len, int64_t* n_rows, int* n_cols) {
  if (!text || len < 0 || !n_rows || !n_cols) return fail(nullptr, SPK_ERR_INVALID, "spk_prob_csv_shape: bad argument");
  const char *cur = text, *end = text + len, *b, *e;
  if (!next_line(cur, end, b, e) || b == e) return fail(nullptr, SPK_ERR_PARSE, "spk_prob_csv_shape: no header line");
  if (memchr(b, '"', (size_t)(e - b))) return fail(nullptr, SPK_ERR_PARSE, "spk_prob_csv_shape: quoted header");
  int fields = 1;
  for (const char* p = b; p < e; ++p) fields += *p == ',';
  int64_t rows = 0;
  while (next_line(cur, end, b, e))
    if (b != e) ++rows;  // pandas skips blank lines
  *n_rows = rows;
  *n_cols = fields - 1;
  return SPK_OK;
}

int spk_prob_csv_parse(const char* text, int64_t len, int64_t n_rows, int n_cols, int64_t* roi, double* values) {
  if (!text || len < 0 || n_rows < 0 || n_cols < 0 || (n_rows > 0 && (!roi || (n_cols > 0 && !values))))
    return fail(nullptr, SPK_ERR_INVALID, "spk_prob_csv_parse: bad argument");
  const char *cur = text, *end = text + len, *b, *e;
  if (!next_line(cur, end, b, e)) return fail(nullptr, SPK_ERR_PARSE, "spk_prob_csv_parse: no header line");
  int64_t row = 0;
  while (next_line(cur, end, b, e)) {
    if (b == e) continue;
    if (row >= n_rows) return fail(nullptr, SPK_ERR_PARSE, "spk_prob_csv_parse: more rows than announced");
    const char* p = b;
    // the ROI id: plain digits (what both writers produce)
    int64_t id = 0;
    int nd = 0;
    for (; p < e && *p >= '0' && *p <= '9' && nd < 18; ++p, ++nd) id = id * 10 + (*p - '0');
    if (nd == 0 || (p < e && *p != ',') || (p == e && n_cols > 0))
      return fail(nullptr, SPK_ERR_PARSE, "spk_prob_csv_parse: line %lld: bad ROI id", (long long)(row + 2));
    roi[row] = id;
    for (int c = 0; c < n_cols; ++c) {
      if (p >= e || *p != ',') return fail(nullptr, SPK_ERR_PARSE, "spk_prob_csv_parse: line %lld: %d values expected", (long long)(row + 2), n_cols);
      ++p;
      const char* q = (const char*)memchr(p, ',', (size_t)(e - p));
      if (!q) q = e;
      if (!parse_decimal(p, q, &values[row * n_cols + c]))
        return fail(nullptr, SPK_ERR_PARSE, "spk_prob_csv_parse: line %lld, value %d is not a number", (long long)(row + 2), c + 1);
      p = q;
    }
    if (p != e) return fail(nullptr, SPK_ERR_PARSE, "spk_prob_csv_parse: line %lld: more than %d values", (long long)(row + 2), n_cols);
    ++row;
  }
  if (row != n_rows) return fail(nullptr, SPK_ERR_PARSE, "spk_prob_csv_parse: fewer rows than announced");
  return SPK_OK;
}

}  // extern "C"
