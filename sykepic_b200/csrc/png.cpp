// Host-side PNG decode for the image mode of `sykepic prob` (--image-dir / --images, sykepic/compute/probability.py:28-36,
// 165-177): the reference reads every ROI image with cv2.imread (sykepic/train/data.py:217-219), one file per DataLoader item.
// Here a whole sample's files are decoded by a few threads straight into ONE byte stream (the layout of a `.roi` file), which
// then takes the device path of a raw bin.  8-bit gray, gray + alpha, RGB(A) with equal colour channels; non-interlaced.
// Container parsing, CRC check and inflate (zlib) and the five scanline filters; no CUDA here.
#include <zlib.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <thread>
#include <vector>

#include "spk_internal.h"

namespace spk {
namespace {

inline uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }

struct PngHeader {
  uint32_t w = 0, h = 0;
  int depth = 0, ctype = 0, interlace = 0, chans = 0;
};

const unsigned char kSig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};

// IHDR is the first chunk: signature (8) + length (4) + type (4) + 13 bytes of header
bool parse_ihdr(const unsigned char* d, size_t n, PngHeader* hd, std::string* err) {
  if (n < 33 || memcmp(d, kSig, 8) != 0) {
    *err = "not a PNG file";
    return false;
  }
  if (be32(d + 8) != 13 || memcmp(d + 12, "IHDR", 4) != 0) {
    *err = "no IHDR";
    return false;
  }
  hd->w = be32(d + 16);
  hd->h = be32(d + 20);
  hd->depth = d[24];
  hd->ctype = d[25];
  hd->interlace = d[28];
  hd->chans = hd->ctype == 0 ? 1 : hd->ctype == 2 ? 3 : hd->ctype == 4 ? 2 : hd->ctype == 6 ? 4 : 0;
  if (hd->depth != 8 || hd->chans == 0 || hd->interlace != 0) {
    char b[96];
    snprintf(b, sizeof b, "unsupported PNG (depth %d, colour type %d, interlace %d)", hd->depth, hd->ctype, hd->interlace);
    *err = b;
    return false;
  }
  // (an IFCB ROI is at most 1380 x 1034; the cap keeps a forged header from asking for gigabytes)
  if (hd->w == 0 || hd->h == 0 || hd->w > 65535u || hd->h > 65535u || (uint64_t)hd->w * hd->h * (uint64_t)hd->chans > (1ull << 28)) {
    *err = "bad image size";
    return false;
  }
  return true;
}

bool read_file(const char* path, size_t limit, std::vector<unsigned char>* out, std::string* err) {
  FILE* f = fopen(path, "rb");
  if (!f) {
    *err = "cannot open";
    return false;
  }
  out->clear();
  unsigned char buf[1 << 16];
  size_t got;
  while ((got = fread(buf, 1, limit && out->size() + sizeof buf > limit ? limit - out->size() : sizeof buf, f)) > 0) {
    out->insert(out->end(), buf, buf + got);
    if (limit && out->size() >= limit) break;
  }
  fclose(f);
  return true;
}

}  // namespace

// the five scanline filters of an inflated IDAT stream (shared with spk_png_unfilter); returns the row of a bad filter type, -1 if none
int64_t png_unfilter_rows(const uint8_t* raw, int64_t h, int64_t stride, int bpp, uint8_t* out) {
  const uint8_t* prev = nullptr;
  for (int64_t y = 0; y < h; ++y) {
    const uint8_t* line = raw + y * (stride + 1);
    const int f = line[0];
    ++line;
    uint8_t* cur = out + y * stride;
    const int64_t head = stride < bpp ? stride : bpp;  // the first pixel has no left neighbour
    switch (f) {
      case 0:
        memcpy(cur, line, (size_t)stride);
        break;
      case 1:  // Sub
        for (int64_t x = 0; x < head; ++x) cur[x] = line[x];
        for (int64_t x = head; x < stride; ++x) cur[x] = (uint8_t)(line[x] + cur[x - bpp]);
        break;
      case 2:  // Up
        if (prev)
          for (int64_t x = 0; x < stride; ++x) cur[x] = (uint8_t)(line[x] + prev[x]);
        else
          memcpy(cur, line, (size_t)stride);
        break;
      case 3:  // Average
        for (int64_t x = 0; x < head; ++x) cur[x] = (uint8_t)(line[x] + ((prev ? prev[x] : 0) >> 1));
        for (int64_t x = head; x < stride; ++x) cur[x] = (uint8_t)(line[x] + ((cur[x - bpp] + (prev ? prev[x] : 0)) >> 1));
        break;
      case 4:  // Paeth
        for (int64_t x = 0; x < head; ++x) cur[x] = (uint8_t)(line[x] + (prev ? prev[x] : 0));  // a = c = 0: the predictor is b
        for (int64_t x = head; x < stride; ++x) {
          const int a = cur[x - bpp], b = prev ? prev[x] : 0, c = prev ? prev[x - bpp] : 0;
          const int pp = a + b - c;
          const int pa = pp > a ? pp - a : a - pp, pb = pp > b ? pp - b : b - pp, pc = pp > c ? pp - c : c - pp;
          const int pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
          cur[x] = (uint8_t)(line[x] + pred);
        }
        break;
      default:
        return y;
    }
    prev = cur;
  }
  return -1;
}

namespace {

// one file -> gray plane of w * h bytes at `out`; the IHDR must say (w, h)
bool decode_one(const char* path, int32_t want_w, int32_t want_h, uint8_t* out, std::string* err) {
  std::vector<unsigned char> file;
  if (!read_file(path, 0, &file, err)) return false;
  PngHeader hd;
  if (!parse_ihdr(file.data(), file.size(), &hd, err)) return false;
  if ((int64_t)hd.w != want_w || (int64_t)hd.h != want_h) {
    *err = "image size changed since it was probed";
    return false;
  }
  const int64_t stride = (int64_t)hd.w * hd.chans, raw_len = (int64_t)hd.h * (stride + 1);
  std::vector<uint8_t> raw((size_t)raw_len);
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit(&zs) != Z_OK) {
    *err = "inflateInit failed";
    return false;
  }
  zs.next_out = raw.data();
  zs.avail_out = (uInt)raw_len;
  bool ended = false, done = false;
  size_t pos = 8;
  while (pos + 12 <= file.size() && !ended) {
    const uint32_t len = be32(file.data() + pos);
    const unsigned char* typ = file.data() + pos + 4;
    if ((size_t)len > file.size() - pos - 12) {
      *err = "truncated chunk";
      break;
    }
    const unsigned char* body = typ + 4;
    if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), typ, len + 4) != be32(body + len)) {
      *err = "CRC error";  // libpng (cv2.imread) rejects the file too
      break;
    }
    if (memcmp(typ, "IDAT", 4) == 0 && !done) {
      zs.next_in = const_cast<unsigned char*>(body);
      zs.avail_in = len;
      const int rc = inflate(&zs, Z_NO_FLUSH);
      if (rc == Z_STREAM_END) {
        done = true;
      } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
        *err = "corrupt image data (inflate)";
        break;
      } else if (zs.avail_in != 0) {  // output full but the stream goes on
        *err = "more image data than the header announces";
        break;
      }
    } else if (memcmp(typ, "IEND", 4) == 0) {
      ended = true;
    }
    pos += 12 + (size_t)len;
  }
  const int64_t produced = raw_len - (int64_t)zs.avail_out;
  inflateEnd(&zs);
  if (!err->empty()) return false;
  if (produced != raw_len) {
    char b[96];
    snprintf(b, sizeof b, "%lld bytes of image data, expected %lld", (long long)produced, (long long)raw_len);
    *err = b;
    return false;
  }
  if (hd.chans == 1) {
    const int64_t bad = png_unfilter_rows(raw.data(), hd.h, stride, 1, out);
    if (bad >= 0) {
      *err = "bad filter type";
      return false;
    }
    return true;
  }
  std::vector<uint8_t> px((size_t)((int64_t)hd.h * stride));
  if (png_unfilter_rows(raw.data(), hd.h, stride, hd.chans, px.data()) >= 0) {
    *err = "bad filter type";
    return false;
  }
  const int64_t n = (int64_t)hd.w * hd.h;
  const int c = hd.chans;
  for (int64_t i = 0; i < n; ++i) {
    const uint8_t* p = px.data() + i * c;
    if (c >= 3 && (p[0] != p[1] || p[0] != p[2])) {
      *err = "colour PNG; IFCB ROI images are grayscale";
      return false;
    }
    out[i] = p[0];
  }
  return true;
}

template <typename F>
int64_t parallel_first_bad(int64_t n, int threads, std::string* err, F&& job) {
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads > 16) threads = 16;
  if (threads < 1) threads = 1;
  if ((int64_t)threads > n) threads = (int)(n > 0 ? n : 1);
  std::atomic<int64_t> next(0), first_bad(INT64_MAX);
  std::vector<std::string> errs((size_t)threads);
  std::vector<int64_t> bad_idx((size_t)threads, INT64_MAX);
  auto worker = [&](int t) {
    for (;;) {
      const int64_t i = next.fetch_add(1);
      if (i >= n || i > first_bad.load()) return;
      std::string e;
      bool ok = false;
      try {
        ok = job(i, &e);
      } catch (const std::exception& ex) {  // nothing may unwind through a thread or the C ABI
        e = ex.what();
      }
      if (!ok) {
        if (i < bad_idx[(size_t)t]) {
          bad_idx[(size_t)t] = i;
          errs[(size_t)t] = e;
        }
        int64_t cur = first_bad.load();
        while (i < cur && !first_bad.compare_exchange_weak(cur, i)) {
        }
      }
    }
  };
  if (threads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker, t);
    for (auto& th : pool) th.join();
  }
  const int64_t fb = first_bad.load();
  if (fb == INT64_MAX) return -1;
  for (int t = 0; t < threads; ++t)
    if (bad_idx[(size_t)t] == fb) *err = errs[(size_t)t];
  return fb;
}

}  // namespace
}  // namespace spk

using namespace spk;

extern "C" {

int spk_png_unfilter(const uint8_t* raw, int64_t h, int64_t stride, int bpp, uint8_t* out) {
  if (!raw || !out || h < 0 || stride < 0 || bpp < 1 || bpp > 4) return fail(nullptr, SPK_ERR_INVALID, "spk_png_unfilter: bad argument");
  const int64_t bad = png_unfilter_rows(raw, h, stride, bpp, out);
  if (bad >= 0) return fail(nullptr, SPK_ERR_PARSE, "spk_png_unfilter: filter type %d in row %lld", (int)raw[bad * (stride + 1)], (long long)bad);
  return SPK_OK;
}

int spk_png_probe(const char* const* paths, int64_t n, int32_t* width, int32_t* height, int threads, int64_t* first_bad) {
  if (n < 0 || (n > 0 && (!paths || !width || !height))) return fail(nullptr, SPK_ERR_INVALID, "spk_png_probe: bad argument");
  std::string err;
  const int64_t bad = parallel_first_bad(n, threads, &err, [&](int64_t i, std::string* e) {
    std::vector<unsigned char> head;
    PngHeader hd;
    if (!read_file(paths[i], 33, &head, e) || !parse_ihdr(head.data(), head.size(), &hd, e)) return false;
    width[i] = (int32_t)hd.w;
    height[i] = (int32_t)hd.h;
    return true;
  });
  if (first_bad) *first_bad = bad;
  if (bad >= 0) return fail(nullptr, SPK_ERR_PARSE, "%s: %s", paths[bad], err.c_str());
  return SPK_OK;
}

int spk_png_decode_batch(const char* const* paths, int64_t n, const int32_t* width, const int32_t* height, const int64_t* start,
                         uint8_t* out, int64_t out_len, int threads, int64_t* first_bad) {
  if (n < 0 || (n > 0 && (!paths || !width || !height || !start || !out))) return fail(nullptr, SPK_ERR_INVALID, "spk_png_decode_batch: bad argument");
  for (int64_t i = 0; i < n; ++i)
    if (width[i] < 1 || height[i] < 1 || start[i] < 0 || start[i] + (int64_t)width[i] * height[i] > out_len)
      return fail(nullptr, SPK_ERR_CAPACITY, "spk_png_decode_batch: image %lld does not fit the output buffer", (long long)i);
  std::string err;
  const int64_t bad = parallel_first_bad(n, threads, &err, [&](int64_t i, std::string* e) { return decode_one(paths[i], width[i], height[i], out + start[i], e); });
  if (first_bad) *first_bad = bad;
  if (bad >= 0) return fail(nullptr, SPK_ERR_PARSE, "%s: %s", paths[bad], err.c_str());
  return SPK_OK;
}

}  // extern "C"
