// json_io.cpp — native reader / writer for the two JSON schemas on either side of the path (host code).
//
// In : the submission / annotation files the reference reads with json.load — a list of
//      {"image_id": str, "category_id": int, "bbox": [x, y, w, h], "score": float}
//      (detnet/data/coco.py:229-252 writes them; detnet/ensemble.py:79 and
//      tracking/utils.py:65-67 read them, the latter also accepts {"annotations": [...]} and rows
//      without "score").
// Out: what the reference writes with json.dump (default separators ", " / ": ", ensure_ascii):
//      the ensemble rows of detnet/ensemble.py:61-62,159-160 and the tracker rows of
//      tracking/utils.py:52-58 / tracking/track.py:50.  Floats are printed like Python's
//      float.__repr__ (shortest digits that round-trip; fixed notation for 1e-4 <= |x| < 1e16,
//      else d.ddde[+-]XX), so the files are byte-identical to the reference's.
//
// At test scale (15 M detections per file) Python's json + dict loops cost minutes on both sides
// of a pipeline that runs in tens of milliseconds; this module parses into flat arrays directly.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <charconv>
#include <climits>
#include <memory>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "w2t.h"

namespace w2t {
void set_last_error(const char *fmt, ...);
}

struct w2t_json_dets {
  std::vector<int32_t> image_index, category;
  std::vector<double> bbox, score;
  std::vector<uint8_t> has_score;
  std::vector<std::string> image_ids;  // unique, first-appearance order
  std::unordered_map<std::string, int32_t> lookup;
  std::string names;  // image ids joined by '\n' (built on demand)
  int32_t last_index = -1;  // image of the previous row (parser cache)
};

namespace {

// host threads this process may use: the affinity mask (one rank per GPU binds itself to a share of the cores),
// capped by W2T_JSON_THREADS
int usable_cores() {
  int T = (int)std::thread::hardware_concurrency();
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof set, &set) == 0) T = std::max(1, CPU_COUNT(&set));
  if (const char *e = getenv("W2T_JSON_THREADS")) T = atoi(e);
  return std::max(T, 1);
}

// W2T_JSON_TIMING=1: phase times of the host packers / writers on stderr
struct Lap {
  std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
  const bool on = getenv("W2T_JSON_TIMING") != nullptr;
  void operator()(const char *what) {
    const auto now = std::chrono::steady_clock::now();
    if (on) fprintf(stderr, "[w2t json] %s %.3f s\n", what, std::chrono::duration<double>(now - last).count());
    last = now;
  }
};

struct Parser {
  const char *p, *end;
  const char *err = nullptr;

  void ws() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
  }
  bool fail(const char *what) {
    if (!err) err = what;
    return false;
  }
  bool expect(char c) {
    ws();
    if (p < end && *p == c) { p++; return true; }
    return fail("unexpected character");
  }
  static void put_utf8(std::string &out, uint32_t cp) {
    if (cp < 0x80) out += (char)cp;
    else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) {
      out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F));
    } else {
      out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F));
      out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F));
    }
  }
  bool hex4(uint32_t &v) {
    if (end - p < 4) return fail("truncated \\u escape");
    v = 0;
    for (int i = 0; i < 4; i++) {
      const char c = *p++;
      v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0';
      else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
      else return fail("bad \\u escape");
    }
    return true;
  }
  // string body after the opening quote; fast path: no escapes -> view into the buffer
  bool string(std::string_view &view, std::string &scratch) {
    ws();
    if (p >= end || *p != '"') return fail("string expected");
    p++;
    const char *s = p;
    while (p < end && *p != '"' && *p != '\\') p++;
    if (p >= end) return fail("unterminated string");
    if (*p == '"') { view = std::string_view(s, p - s); p++; return true; }
    scratch.assign(s, p - s);
    while (p < end && *p != '"') {
      if (*p != '\\') { scratch += *p++; continue; }
      p++;
      if (p >= end) return fail("unterminated string");
      const char c = *p++;
      switch (c) {
        case '"': scratch += '"'; break;
        case '\\': scratch += '\\'; break;
        case '/': scratch += '/'; break;
        case 'b': scratch += '\b'; break;
        case 'f': scratch += '\f'; break;
        case 'n': scratch += '\n'; break;
        case 'r': scratch += '\r'; break;
        case 't': scratch += '\t'; break;
        case 'u': {
          uint32_t cp;
          if (!hex4(cp)) return false;
          if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
            p += 2;
            uint32_t lo;
            if (!hex4(lo)) return false;
            if (lo >= 0xDC00 && lo < 0xE000) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          }
          put_utf8(scratch, cp);
          break;
        }
        default: return fail("bad escape");
      }
    }
    if (p >= end) return fail("unterminated string");
    p++;
    view = scratch;
    return true;
  }
  bool number(double &v) {
    ws();
    const char *s = p;
    {
      // plain integers of at most 15 digits (pixel boxes, category ids): exact in a double, no from_chars
      const char *q = p;
      const bool neg = q < end && *q == '-';
      if (neg) q++;
      const char *d0 = q;
      uint64_t acc = 0;
      while (q < end && *q >= '0' && *q <= '9' && q - d0 < 15) acc = acc * 10 + (uint64_t)(*q++ - '0');
      if (q > d0 && q < end && (*q == ',' || *q == ']' || *q == '}' || *q == ' ' || *q == '\n') &&
          !(q - d0 > 1 && *d0 == '0')) {       // leading zeros: the general path decides
        v = neg ? -(double)acc : (double)acc;
        if (neg && acc == 0) v = -0.0;
        p = q;
        return true;
      }
    }
    if (p < end && (*p == '-' || *p == '+')) p++;
    while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+')) p++;
    if (s == p) {
      // json.load also accepts NaN / Infinity / -Infinity
      if (end - p >= 3 && !strncmp(p, "NaN", 3)) { p += 3; v = NAN; return true; }
      if (end - p >= 8 && !strncmp(p, "Infinity", 8)) { p += 8; v = INFINITY; return true; }
      return fail("number expected");
    }
    if (p - s == 1 && *s == '-' && end - p >= 8 && !strncmp(p, "Infinity", 8)) { p += 8; v = -INFINITY; return true; }
    const char *b = (*s == '+') ? s + 1 : s;
    const auto r = std::from_chars(b, p, v);  // correctly rounded, like Python's float()
    if (r.ec != std::errc() && r.ec != std::errc::result_out_of_range) return fail("bad number");
    if (r.ptr != p) return fail("bad number");  // "1-2", "1e": json.load raises as well
    return true;
  }
  bool skip_value() {
    ws();
    if (p >= end) return fail("value expected");
    if (*p == '"') { std::string_view v; std::string s; return string(v, s); }
    if (*p == '{') {
      p++;
      ws();
      if (p < end && *p == '}') { p++; return true; }
      for (;;) {
        std::string_view k; std::string s;
        if (!string(k, s) || !expect(':') || !skip_value()) return false;
        ws();
        if (p < end && *p == ',') { p++; continue; }
        return expect('}');
      }
    }
    if (*p == '[') {
      p++;
      ws();
      if (p < end && *p == ']') { p++; return true; }
      for (;;) {
        if (!skip_value()) return false;
        ws();
        if (p < end && *p == ',') { p++; continue; }
        return expect(']');
      }
    }
    if (end - p >= 4 && !strncmp(p, "true", 4)) { p += 4; return true; }
    if (end - p >= 4 && !strncmp(p, "null", 4)) { p += 4; return true; }
    if (end - p >= 5 && !strncmp(p, "false", 5)) { p += 5; return true; }
    double d;
    return number(d);
  }

  bool detection(w2t_json_dets &out) {
    if (!expect('{')) return false;
    bool have_id = false, have_cat = false, have_box = false, have_score = false;
    double cat = 0, box[4] = {0, 0, 0, 0}, score = 1.0;
    std::string id_scratch, key_scratch;
    std::string_view id;
    ws();
    if (p < end && *p == '}') { p++; return fail("detection without image_id"); }
    for (;;) {
      std::string_view key;
      if (!string(key, key_scratch) || !expect(':')) return false;
      if (key == "image_id") {
        if (!string(id, id_scratch)) return false;
        have_id = true;
      } else if (key == "category_id") {
        if (!number(cat)) return false;
        have_cat = true;
      } else if (key == "score") {
        if (!number(score)) return false;
        have_score = true;
      } else if (key == "bbox") {
        if (!expect('[')) return false;
        for (int i = 0; i < 4; i++) {
          if (!number(box[i])) return false;
          if (i < 3 && !expect(',')) return false;
        }
        ws();
        // longer boxes would change `[score] + bbox` in the reference; refuse instead of guessing
        if (!expect(']')) return fail("bbox must have exactly four numbers");
        have_box = true;
      } else if (!skip_value()) {
        return false;
      }
      ws();
      if (p < end && *p == ',') { p++; continue; }
      if (!expect('}')) return false;
      break;
    }
    if (!have_id || !have_cat || !have_box) return fail("detection needs image_id, category_id and bbox");
    int32_t idx;
    if (out.last_index >= 0 && id == std::string_view(out.image_ids[(size_t)out.last_index])) {
      idx = out.last_index;                  // rows of one image are consecutive in detector output
    } else {
      auto it = out.lookup.find(std::string(id));
      if (it == out.lookup.end()) {
        idx = (int32_t)out.image_ids.size();
        out.image_ids.emplace_back(id);
        out.lookup.emplace(out.image_ids.back(), idx);
      } else {
        idx = it->second;
      }
      out.last_index = idx;
    }
    if (!(cat == (double)(int32_t)cat)) return fail("category_id must be an integer");
    out.image_index.push_back(idx);
    out.category.push_back((int32_t)cat);
    out.bbox.insert(out.bbox.end(), box, box + 4);
    out.score.push_back(score);
    out.has_score.push_back(have_score ? 1 : 0);
    return true;
  }

  // The list of detections, split over host threads: candidate split points are places where one record ends
  // and the next begins (`}` `,` `{`); a thread parses from its split point up to exactly the next one.  A split
  // point that fell inside a string shows up as a parse error or a missed rendezvous, and the caller then parses
  // sequentially — so the result never depends on the split.
  bool list_parallel(w2t_json_dets &out, int T) {
    std::vector<const char *> starts(1, p);
    const size_t total = (size_t)(end - p);
    for (int t = 1; t < T; t++) {
      const char *q = p + total * (size_t)t / (size_t)T;
      if (q <= starts.back()) continue;
      const char *hit = nullptr;
      for (; q + 1 < end; q++) {
        if (*q != '}') continue;
        const char *r = q + 1;
        while (r < end && (*r == ' ' || *r == '\n' || *r == '\t' || *r == '\r')) r++;
        if (r >= end || *r != ',') continue;
        r++;
        while (r < end && (*r == ' ' || *r == '\n' || *r == '\t' || *r == '\r')) r++;
        if (r < end && *r == '{') { hit = r; break; }
      }
      if (hit == nullptr) break;
      if (hit > starts.back()) starts.push_back(hit);
    }
    const int n = (int)starts.size();
    if (n < 2) return false;
    std::vector<w2t_json_dets> part((size_t)n);
    std::vector<const char *> stop((size_t)n, nullptr);
    std::vector<char> good((size_t)n, 0);
    auto work = [&](int t) {
      Parser sub{starts[t], end};
      const char *limit = (t + 1 < n) ? starts[t + 1] : nullptr;
      w2t_json_dets &o = part[t];
      const size_t guess = (size_t)((limit ? limit : end) - starts[t]) / 90 + 16;
      o.image_index.reserve(guess); o.category.reserve(guess); o.bbox.reserve(4 * guess);
      o.score.reserve(guess); o.has_score.reserve(guess);
      for (;;) {
        if (!sub.detection(o)) return;
        sub.ws();
        if (sub.p < sub.end && *sub.p == ',') {
          sub.p++;
          sub.ws();
          if (limit != nullptr) {
            if (sub.p == limit) break;       // rendezvous with the next thread's first record
            if (sub.p > limit) return;       // ran past it: the split point was not a record boundary
          }
          continue;
        }
        if (limit != nullptr) return;        // the list ended inside a middle chunk?!
        if (!sub.expect(']')) return;
        break;
      }
      stop[t] = sub.p;
      good[t] = 1;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n; t++) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
    for (int t = 0; t < n; t++)
      if (!good[t]) return false;
    // merge: image ids keep first-appearance order (sequential, ids only); the rows are copied by the threads
    std::vector<std::vector<int32_t>> remap((size_t)n);
    std::vector<size_t> base((size_t)n + 1, out.image_index.size());
    for (int t = 0; t < n; t++) {
      w2t_json_dets &o = part[t];
      remap[t].resize(o.image_ids.size());
      for (size_t i = 0; i < o.image_ids.size(); i++) {
        auto it = out.lookup.find(o.image_ids[i]);
        if (it == out.lookup.end()) {
          const int32_t idx = (int32_t)out.image_ids.size();
          out.image_ids.push_back(std::move(o.image_ids[i]));
          out.lookup.emplace(out.image_ids.back(), idx);
          remap[t][i] = idx;
        } else {
          remap[t][i] = it->second;
        }
      }
      base[t + 1] = base[t] + o.image_index.size();
    }
    const size_t total_rows = base[n];
    out.image_index.resize(total_rows); out.category.resize(total_rows); out.bbox.resize(4 * total_rows);
    out.score.resize(total_rows); out.has_score.resize(total_rows);
    auto copy = [&](int t) {
      const w2t_json_dets &o = part[t];
      const size_t b = base[t], m = o.image_index.size();
      for (size_t i = 0; i < m; i++) out.image_index[b + i] = remap[t][(size_t)o.image_index[i]];
      if (m) {
        memcpy(&out.category[b], o.category.data(), 4 * m);
        memcpy(&out.bbox[4 * b], o.bbox.data(), 32 * m);
        memcpy(&out.score[b], o.score.data(), 8 * m);
        memcpy(&out.has_score[b], o.has_score.data(), m);
      }
    };
    pool.clear();
    for (int t = 1; t < n; t++) pool.emplace_back(copy, t);
    copy(0);
    for (auto &th : pool) th.join();
    out.last_index = -1;
    p = stop[n - 1];
    return true;
  }

  static int list_threads(size_t bytes) {
    int T = std::min(usable_cores(), 32);
    T = std::min<int>(T, (int)(bytes >> 21));  // at least 2 MB of text per thread
    return std::max(T, 1);
  }

  bool list(w2t_json_dets &out) {
    if (!expect('[')) return false;
    ws();
    if (p < end && *p == ']') { p++; return true; }
    {
      const int T = list_threads((size_t)(end - p));
      const char *p0 = p;
      const size_t rows0 = out.image_index.size(), ids0 = out.image_ids.size();
      if (T > 1) {
        if (list_parallel(out, T)) return true;
        // undo anything a failed attempt merged (it merges only after every thread succeeded: nothing), reparse
        p = p0;
        err = nullptr;
        (void)rows0; (void)ids0;
      }
    }
    {
      const size_t guess = out.image_index.size() + (size_t)(end - p) / 90 + 16;
      out.image_index.reserve(guess); out.category.reserve(guess); out.bbox.reserve(4 * guess);
      out.score.reserve(guess); out.has_score.reserve(guess);
    }
    for (;;) {
      if (!detection(out)) return false;
      ws();
      if (p < end && *p == ',') { p++; continue; }
      return expect(']');
    }
  }

  bool document(w2t_json_dets &out) {
    ws();
    if (p < end && *p == '[') return list(out);
    if (p < end && *p == '{') {  // {"annotations": [...], ...}  (tracking/utils.py:66-67)
      p++;
      bool found = false;
      ws();
      if (p < end && *p == '}') return fail("object without \"annotations\"");
      for (;;) {
        std::string_view key; std::string s;
        if (!string(key, s) || !expect(':')) return false;
        if (key == "annotations") {
          if (!list(out)) return false;
          found = true;
        } else if (!skip_value()) {
          return false;
        }
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (!expect('}')) return false;
        break;
      }
      return found ? true : fail("object without \"annotations\"");
    }
    return fail("a JSON list or an object with \"annotations\" expected");
  }
};

// the file mapped read-only: no copy, and the parser threads fault their own parts of it in
struct MappedFile {
  const char *data = nullptr;
  size_t size = 0;
  bool mapped = false;
  std::string fallback;      // pipes and other unmappable inputs
  bool open(const char *path) {
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode)) {
      size = (size_t)st.st_size;
      if (size == 0) { ::close(fd); data = ""; return true; }
      void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m != MAP_FAILED) {
        madvise(m, size, MADV_WILLNEED);
        data = static_cast<const char *>(m);
        mapped = true;
        ::close(fd);
        return true;
      }
    }
    char chunk[1 << 16];
    for (;;) {
      const ssize_t got = ::read(fd, chunk, sizeof chunk);
      if (got < 0) { ::close(fd); return false; }
      if (got == 0) break;
      fallback.append(chunk, (size_t)got);
    }
    ::close(fd);
    data = fallback.data();
    size = fallback.size();
    return true;
  }
  ~MappedFile() {
    if (mapped) munmap(const_cast<char *>(data), size);
  }
};

// Python's float.__repr__
void py_float(std::string &out, double x) {
  if (std::isnan(x)) { out += "NaN"; return; }          // json.dump spellings
  if (std::isinf(x)) { out += x > 0 ? "Infinity" : "-Infinity"; return; }
  if (x == 0) { out += std::signbit(x) ? "-0.0" : "0.0"; return; }
  char sci[40], text[64];
  const auto r = std::to_chars(sci, sci + sizeof sci, x, std::chars_format::scientific);  // shortest round-trip digits
  // sci = [-]d[.ddd]e[+-]XX[X]
  const char *q = sci;
  char *w = text;
  if (*q == '-') *w++ = *q++;
  char digits[24];
  int nd = 0;
  for (; *q != 'e'; q++)
    if (*q != '.') digits[nd++] = *q;
  q++;
  const bool eneg = *q == '-';
  q++;
  int a = 0;
  for (; q < r.ptr; q++) a = a * 10 + (*q - '0');
  const int exp10 = eneg ? -a : a;
  if (exp10 >= -4 && exp10 < 16) {
    if (exp10 < 0) {
      *w++ = '0'; *w++ = '.';
      for (int i = 0; i < -exp10 - 1; i++) *w++ = '0';
      memcpy(w, digits, (size_t)nd); w += nd;
    } else if (nd <= exp10 + 1) {
      memcpy(w, digits, (size_t)nd); w += nd;
      for (int i = 0; i < exp10 + 1 - nd; i++) *w++ = '0';
      *w++ = '.'; *w++ = '0';
    } else {
      memcpy(w, digits, (size_t)exp10 + 1); w += exp10 + 1;
      *w++ = '.';
      memcpy(w, digits + exp10 + 1, (size_t)(nd - exp10 - 1)); w += nd - exp10 - 1;
    }
  } else {
    *w++ = digits[0];
    if (nd > 1) { *w++ = '.'; memcpy(w, digits + 1, (size_t)nd - 1); w += nd - 1; }
    *w++ = 'e';
    *w++ = eneg ? '-' : '+';
    if (a < 10) *w++ = '0';
    if (a >= 100) *w++ = (char)('0' + a / 100);
    if (a >= 10) *w++ = (char)('0' + (a / 10) % 10);
    *w++ = (char)('0' + a % 10);
  }
  out.append(text, (size_t)(w - text));
}

inline void put_int(std::string &out, long long v) {
  char buf[24];
  const auto r = std::to_chars(buf, buf + sizeof buf, v);
  out.append(buf, (size_t)(r.ptr - buf));
}

// json.dumps(str) with ensure_ascii=True
void py_string(std::string &out, const char *s) {
  static const char *hex = "0123456789abcdef";
  out += '"';
  const unsigned char *u = reinterpret_cast<const unsigned char *>(s);
  while (*u) {
    uint32_t cp = *u;
    int extra = 0;
    if (cp >= 0xF0) { cp &= 0x07; extra = 3; }
    else if (cp >= 0xE0) { cp &= 0x0F; extra = 2; }
    else if (cp >= 0xC0) { cp &= 0x1F; extra = 1; }
    u++;
    for (int i = 0; i < extra && *u; i++) cp = (cp << 6) | (*u++ & 0x3F);
    auto u16 = [&](uint32_t v) {
      out += "\\u";
      for (int sh = 12; sh >= 0; sh -= 4) out += hex[(v >> sh) & 0xF];
    };
    if (cp == '"') out += "\\\"";
    else if (cp == '\\') out += "\\\\";
    else if (cp == '\n') out += "\\n";
    else if (cp == '\r') out += "\\r";
    else if (cp == '\t') out += "\\t";
    else if (cp == '\b') out += "\\b";
    else if (cp == '\f') out += "\\f";
    else if (cp < 0x20) u16(cp);
    else if (cp < 0x80) out += (char)cp;
    else if (cp < 0x10000) u16(cp);
    else { cp -= 0x10000; u16(0xD800 + (cp >> 10)); u16(0xDC00 + (cp & 0x3FF)); }
  }
  out += '"';
}

bool write_all(const char *path, const std::string &data, bool append) {
  FILE *f = fopen(path, append ? "ab" : "wb");
  if (!f) return false;
  const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
  return (fclose(f) == 0) && ok;
}

}  // namespace

namespace {
// parse `path` into a fresh handle; on failure nullptr and the message in `error` (the library's last error is
// thread-local, and this also runs on worker threads)
w2t_json_dets *load_file(const char *path, std::string &error) {
  MappedFile file;
  if (!file.open(path)) {
    error = std::string("cannot read ") + path;
    return nullptr;
  }
  auto *h = new w2t_json_dets;
  Parser ps{file.data, file.data + file.size};
  bool ok = ps.document(*h);
  if (ok) {
    ps.ws();
    if (ps.p != ps.end) ok = ps.fail("trailing data after the JSON document");
  }
  if (!ok) {
    error = std::string(ps.err ? ps.err : "parse error") + " at byte " + std::to_string((long long)(ps.p - file.data)) +
            " of " + path;
    delete h;
    return nullptr;
  }
  h->lookup.clear();
  return h;
}
}  // namespace

extern "C" int w2t_json_load(const char *path, w2t_json_dets_t **out) {
  if (!path || !out) { w2t::set_last_error("w2t_json_load: bad argument"); return W2T_ERR_ARG; }
  std::string error;
  *out = load_file(path, error);
  if (!*out) w2t::set_last_error("w2t_json_load: %s", error.c_str());
  return *out ? W2T_OK : W2T_ERR_ARG;
}

extern "C" int64_t w2t_json_count(const w2t_json_dets_t *h) { return h ? (int64_t)h->image_index.size() : 0; }
extern "C" int64_t w2t_json_n_images(const w2t_json_dets_t *h) { return h ? (int64_t)h->image_ids.size() : 0; }

extern "C" int w2t_json_copy(const w2t_json_dets_t *h, int32_t *image_index, int32_t *category, double *bbox,
                             double *score, uint8_t *has_score) {
  if (!h) { w2t::set_last_error("w2t_json_copy: null handle"); return W2T_ERR_ARG; }
  const size_t n = h->image_index.size();
  if (image_index) memcpy(image_index, h->image_index.data(), 4 * n);
  if (category) memcpy(category, h->category.data(), 4 * n);
  if (bbox) memcpy(bbox, h->bbox.data(), 32 * n);
  if (score) memcpy(score, h->score.data(), 8 * n);
  if (has_score) memcpy(has_score, h->has_score.data(), n);
  return W2T_OK;
}

extern "C" const char *w2t_json_image_ids(w2t_json_dets_t *h, int64_t *bytes) {
  if (!h) return nullptr;
  if (h->names.empty() && !h->image_ids.empty()) {
    size_t total = 0;
    for (const auto &s : h->image_ids) total += s.size() + 1;
    h->names.reserve(total);
    for (const auto &s : h->image_ids) { h->names += s; h->names += '\n'; }
  }
  if (bytes) *bytes = (int64_t)h->names.size();
  return h->names.data();
}

extern "C" void w2t_json_free(w2t_json_dets_t *h) { delete h; }

// ---- grouping of parsed files for the ensemble (and, fused, for the tracker behind it) -------------------------

struct w2t_json_groups {
  std::vector<std::string> image_ids;      // images that keep at least one row, sorted (ensemble.py:91-95)
  std::string names;                       // the same, '\n'-joined
  std::vector<int32_t> category_ids;       // every category id of the files, ascending
  std::vector<int32_t> image_order;        // [n_img] position k of the layout -> index into image_ids
  std::vector<int32_t> stream_img_offsets; // [S+1] (stream layout only)
  std::vector<int64_t> frame_ids;          // [n_img] in layout order (stream layout only)
  std::vector<int32_t> group_offsets;      // [G+1]
  std::vector<int32_t> sub_counts;         // [G, n_files]
  // rows stay in the parsed files until w2t_json_groups_copy gathers them into the caller's buffer:
  std::vector<std::unique_ptr<w2t_json_dets>> files;
  std::vector<std::vector<double>> wscore; // [file][row] score * weight
  std::unique_ptr<int32_t[]> src;          // [N] row of its file (the file follows from sub_counts)
  std::unique_ptr<uint64_t[]> packed;      // [N] W2T_BOX_LTWH_P64 rows when every row fits them exactly
  int64_t n_rows = 0;
  bool packable = true;
  int32_t n_cat = 0, n_files = 0, max_group = 0;
};

namespace {

// `int(text)` for the plain spellings ([+-]digits); anything else (underscores, spaces, non-ASCII digits) is refused
// and the caller falls back to the Python packer
bool plain_int(std::string_view t, int64_t &v) {
  size_t i = 0;
  bool neg = false;
  if (i < t.size() && (t[i] == '+' || t[i] == '-')) neg = t[i++] == '-';
  if (i >= t.size() || t.size() - i > 18) return false;
  int64_t acc = 0;
  for (; i < t.size(); i++) {
    if (t[i] < '0' || t[i] > '9') return false;
    acc = acc * 10 + (t[i] - '0');
  }
  v = neg ? -acc : acc;
  return true;
}

bool pack_row(const double *r, uint64_t &out) {
  const double k = std::nearbyint(r[0] * 1e5);
  if (!(k >= 0 && k < 131072.0) || k / 1e5 != r[0]) return false;
  for (int j = 1; j < 5; j++)
    if (r[j] != std::nearbyint(r[j])) return false;
  if (!(r[1] >= -3072 && r[1] <= 5119 && r[2] >= -1536 && r[2] <= 2559 && r[3] >= 0 && r[3] <= 2047 && r[4] >= 0 &&
        r[4] <= 2047))
    return false;
  out = (uint64_t)k | ((uint64_t)(r[1] + 3072) << 17) | ((uint64_t)(r[2] + 1536) << 30) | ((uint64_t)r[3] << 42) |
        ((uint64_t)r[4] << 53);
  return true;
}

}  // namespace

extern "C" int w2t_json_group_files(const char *const *paths, int32_t n_files, const double *weights, double min_score,
                                    int32_t layout, int32_t n_classes, w2t_json_groups_t **out) {
  if (!paths || n_files < 1 || !out || (layout != W2T_LAYOUT_ENSEMBLE && layout != W2T_LAYOUT_STREAMS) ||
      (layout == W2T_LAYOUT_STREAMS && n_classes < 1)) {
    w2t::set_last_error("w2t_json_group_files: bad argument");
    return W2T_ERR_ARG;
  }
  *out = nullptr;
  const int K = n_files;
  Lap lap;
  // 1. parse, one thread per file (each splits its file further)
  auto g = std::make_unique<w2t_json_groups>();
  std::vector<std::unique_ptr<w2t_json_dets>> &files = g->files;
  files.resize((size_t)K);
  std::vector<std::string> errors((size_t)K);
  {
    std::vector<std::thread> pool;
    auto job = [&](int f) { files[f].reset(load_file(paths[f], errors[f])); };
    for (int f = 1; f < K; f++) pool.emplace_back(job, f);
    job(0);
    for (auto &th : pool) th.join();
  }
  for (int f = 0; f < K; f++)
    if (!files[f]) {
      w2t::set_last_error("w2t_json_load: %s", errors[f].c_str());
      return W2T_ERR_ARG;
    }
  for (int f = 0; f < K; f++)
    if (std::find(files[f]->has_score.begin(), files[f]->has_score.end(), (uint8_t)0) != files[f]->has_score.end()) {
      w2t::set_last_error("w2t_json_group_files: a row of %s has no \"score\"", paths[f]);
      return W2T_ERR_UNSUPPORTED;  // the general path raises the reference's KeyError
    }
  g->n_files = K;
  lap("parse");
  // 2. filters of convert_submission (ensemble.py:37-42): width and height > 0, score * weight >= min_score
  std::vector<std::vector<double>> &wscore = g->wscore;
  wscore.resize((size_t)K);
  std::vector<std::vector<uint8_t>> keep((size_t)K), used((size_t)K);
  std::vector<std::vector<int32_t>> cats((size_t)K);
  {
    auto job = [&](int f) {
      const w2t_json_dets &d = *files[f];
      const size_t n = d.score.size();
      const double w = weights ? weights[f] : 1.0;
      wscore[f].resize(n); keep[f].resize(n); used[f].assign(d.image_ids.size(), 0);
      int32_t last = INT32_MIN;
      for (size_t i = 0; i < n; i++) {
        const double s = d.score[i] * w;
        wscore[f][i] = s;
        const bool k = d.bbox[4 * i + 2] > 0 && d.bbox[4 * i + 3] > 0 && s >= min_score;
        keep[f][i] = k;
        if (k) used[f][(size_t)d.image_index[i]] = 1;
        if (d.category[i] != last) {
          last = d.category[i];
          if (std::find(cats[f].begin(), cats[f].end(), last) == cats[f].end()) cats[f].push_back(last);
        }
      }
    };
    std::vector<std::thread> pool;
    for (int f = 1; f < K; f++) pool.emplace_back(job, f);
    job(0);
    for (auto &th : pool) th.join();
  }
  lap("filter");
  for (int f = 0; f < K; f++) g->category_ids.insert(g->category_ids.end(), cats[f].begin(), cats[f].end());
  std::sort(g->category_ids.begin(), g->category_ids.end());
  g->category_ids.erase(std::unique(g->category_ids.begin(), g->category_ids.end()), g->category_ids.end());
  // 3. images with a surviving row, sorted like Python sorts str (code points = UTF-8 bytes, unsigned)
  {
    std::vector<std::string_view> ids;
    for (int f = 0; f < K; f++)
      for (size_t i = 0; i < used[f].size(); i++)
        if (used[f][i]) ids.emplace_back(files[f]->image_ids[i]);
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    g->image_ids.assign(ids.begin(), ids.end());
  }
  const int64_t n_img = (int64_t)g->image_ids.size();
  // 4. layout: position of every image, and the column of every category
  std::vector<int32_t> pos((size_t)n_img);          // sorted index -> position in the layout
  g->image_order.resize((size_t)n_img);
  int32_t ncat;
  std::vector<int32_t> cat_col;                      // category id - cat_lo -> column, -1 = none
  int32_t cat_lo = 0;
  if (layout == W2T_LAYOUT_ENSEMBLE) {
    for (int64_t i = 0; i < n_img; i++) pos[i] = g->image_order[i] = (int32_t)i;
    ncat = std::max<int32_t>((int32_t)g->category_ids.size(), 1);
    if (!g->category_ids.empty()) {
      cat_lo = g->category_ids.front();
      const int64_t span = (int64_t)g->category_ids.back() - cat_lo + 1;
      if (span > (1 << 24)) { w2t::set_last_error("w2t_json_group_files: category ids too sparse"); return W2T_ERR_UNSUPPORTED; }
      cat_col.assign((size_t)span, -1);
      for (size_t c = 0; c < g->category_ids.size(); c++) cat_col[(size_t)(g->category_ids[c] - cat_lo)] = (int32_t)c;
    }
  } else {
    // the tracker's streams (tracking/utils.py:63-96 on a file that lists images in sorted order): segments, and
    // cameras inside a segment, in first-appearance order; the frames of a stream in numeric order
    ncat = n_classes;
    if (!g->category_ids.empty() && (g->category_ids.front() < 1 || g->category_ids.back() > n_classes)) {
      w2t::set_last_error("w2t_json_group_files: category id outside 1..%d", n_classes);
      return W2T_ERR_UNSUPPORTED;      // the general path raises the reference's IndexError
    }
    cat_lo = 1;
    cat_col.resize((size_t)n_classes);
    for (int c = 0; c < n_classes; c++) cat_col[c] = c;
    std::unordered_map<std::string_view, int32_t> seg_index;
    std::unordered_map<std::string, int32_t> pair_index;
    std::vector<int32_t> pair_seg, pair_of((size_t)n_img);
    std::vector<int64_t> frame((size_t)n_img);
    std::string key;
    for (int64_t i = 0; i < n_img; i++) {
      const std::string_view id = g->image_ids[i];
      const size_t a = id.find('/');
      const size_t b = a == std::string_view::npos ? a : id.find('/', a + 1);
      if (b == std::string_view::npos || id.find('/', b + 1) != std::string_view::npos ||
          !plain_int(id.substr(a + 1, b - a - 1), frame[i])) {
        w2t::set_last_error("w2t_json_group_files: image id %s is not segment/frame/camera", g->image_ids[i].c_str());
        return W2T_ERR_UNSUPPORTED;
      }
      const int32_t sidx = seg_index.emplace(id.substr(0, a), (int32_t)seg_index.size()).first->second;
      key.assign(id.substr(0, a));
      key += '/';
      key.append(id.substr(b + 1));
      auto it = pair_index.find(key);
      if (it == pair_index.end()) {
        it = pair_index.emplace(key, (int32_t)pair_index.size()).first;
        pair_seg.push_back(sidx);
      }
      pair_of[i] = it->second;
    }
    const int32_t n_pair = (int32_t)pair_seg.size();
    std::vector<int32_t> pair_order((size_t)n_pair), pair_rank((size_t)n_pair);
    for (int32_t q = 0; q < n_pair; q++) pair_order[q] = q;
    std::stable_sort(pair_order.begin(), pair_order.end(), [&](int32_t x, int32_t y) { return pair_seg[x] < pair_seg[y]; });
    for (int32_t r = 0; r < n_pair; r++) pair_rank[pair_order[r]] = r;
    for (int64_t i = 0; i < n_img; i++) g->image_order[i] = (int32_t)i;
    std::stable_sort(g->image_order.begin(), g->image_order.end(), [&](int32_t x, int32_t y) {
      const int32_t rx = pair_rank[pair_of[x]], ry = pair_rank[pair_of[y]];
      return rx != ry ? rx < ry : frame[x] < frame[y];
    });
    g->stream_img_offsets.assign((size_t)n_pair + 1, 0);
    g->frame_ids.resize((size_t)n_img);
    for (int64_t k = 0; k < n_img; k++) {
      const int32_t i = g->image_order[k];
      pos[i] = (int32_t)k;
      g->frame_ids[k] = frame[i];
      g->stream_img_offsets[(size_t)pair_rank[pair_of[i]] + 1]++;
    }
    for (int32_t r = 0; r < n_pair; r++) g->stream_img_offsets[r + 1] += g->stream_img_offsets[r];
  }
  lap("layout");
  g->n_cat = ncat;
  const int64_t G = n_img * ncat;
  if (G >= INT32_MAX) { w2t::set_last_error("w2t_json_group_files: too many groups"); return W2T_ERR_UNSUPPORTED; }
  // 5. group key of every kept row; rows per (group, file)
  std::vector<std::vector<int32_t>> key((size_t)K);
  g->sub_counts.assign((size_t)(G * K), 0);
  {
    std::unordered_map<std::string_view, int32_t> sorted_index;
    sorted_index.reserve((size_t)n_img * 2);
    for (int64_t i = 0; i < n_img; i++) sorted_index.emplace(g->image_ids[i], (int32_t)i);
    std::vector<std::vector<int32_t>> local_pos((size_t)K);
    for (int f = 0; f < K; f++) {
      local_pos[f].assign(files[f]->image_ids.size(), -1);
      for (size_t i = 0; i < used[f].size(); i++)
        if (used[f][i]) local_pos[f][i] = pos[(size_t)sorted_index.find(files[f]->image_ids[i])->second];
    }
    auto job = [&](int f) {
      const w2t_json_dets &d = *files[f];
      const size_t n = d.score.size();
      key[f].resize(n);
      for (size_t i = 0; i < n; i++) {
        if (!keep[f][i]) { key[f][i] = -1; continue; }
        const int32_t gk = local_pos[f][(size_t)d.image_index[i]] * ncat + cat_col[(size_t)(d.category[i] - cat_lo)];
        key[f][i] = gk;
        g->sub_counts[(size_t)gk * K + f]++;
      }
    };
    std::vector<std::thread> pool;
    for (int f = 1; f < K; f++) pool.emplace_back(job, f);
    job(0);
    for (auto &th : pool) th.join();
  }
  lap("keys");
  // 6. offsets: groups in order, inside a group the files in order, inside a file the JSON order (tta.py:9-12)
  g->group_offsets.assign((size_t)G + 1, 0);
  std::vector<int64_t> start((size_t)(G * K));
  int64_t total = 0;
  for (int64_t q = 0; q < G; q++) {
    for (int f = 0; f < K; f++) {
      start[(size_t)(q * K + f)] = total;
      total += g->sub_counts[(size_t)(q * K + f)];
    }
    if (total >= INT32_MAX) { w2t::set_last_error("w2t_json_group_files: more than 2^31 rows"); return W2T_ERR_UNSUPPORTED; }
    g->group_offsets[(size_t)q + 1] = (int32_t)total;
    g->max_group = std::max<int32_t>(g->max_group, g->group_offsets[q + 1] - g->group_offsets[q]);
  }
  g->n_rows = total;
  g->src.reset(new int32_t[(size_t)total + 1]);
  g->packed.reset(new uint64_t[(size_t)total + 1]);
  lap("offsets+alloc");
  // 7. scatter, one thread per file (their destinations are disjoint)
  {
    std::vector<char> fits((size_t)K, 1);
    auto job = [&](int f) {
      const w2t_json_dets &d = *files[f];
      const size_t n = d.score.size();
      bool ok = true;
      for (size_t i = 0; i < n; i++) {
        const int32_t gk = key[f][i];
        if (gk < 0) continue;
        const int64_t at = start[(size_t)gk * K + f]++;
        g->src[(size_t)at] = (int32_t)i;
        const double r[5] = {wscore[f][i], d.bbox[4 * i], d.bbox[4 * i + 1], d.bbox[4 * i + 2], d.bbox[4 * i + 3]};
        if (ok) ok = pack_row(r, g->packed[(size_t)at]);
      }
      fits[f] = ok;
    };
    std::vector<std::thread> pool;
    for (int f = 1; f < K; f++) pool.emplace_back(job, f);
    job(0);
    for (auto &th : pool) th.join();
    for (int f = 0; f < K; f++) g->packable = g->packable && fits[f];
  }
  lap("scatter");
  for (const auto &id : g->image_ids) { g->names += id; g->names += '\n'; }
  *out = g.release();
  return W2T_OK;
}

extern "C" int w2t_json_groups_info(const w2t_json_groups_t *g, int64_t info[8]) {
  if (!g || !info) { w2t::set_last_error("w2t_json_groups_info: null argument"); return W2T_ERR_ARG; }
  info[0] = (int64_t)g->image_ids.size();
  info[1] = g->n_cat;
  info[2] = g->n_rows;
  info[3] = g->max_group;
  info[4] = g->stream_img_offsets.empty() ? 0 : (int64_t)g->stream_img_offsets.size() - 1;
  info[5] = g->packable ? 1 : 0;
  info[6] = (int64_t)g->category_ids.size();
  info[7] = g->n_files;
  return W2T_OK;
}

extern "C" int w2t_json_groups_copy(const w2t_json_groups_t *g, int32_t *category_ids, int32_t *image_order,
                                    int32_t *stream_img_offsets, int64_t *frame_ids, int32_t *group_offsets,
                                    int32_t *sub_counts, double *rows, uint64_t *packed) {
  if (!g) { w2t::set_last_error("w2t_json_groups_copy: null handle"); return W2T_ERR_ARG; }
  auto put = [](auto *dst, const auto &v) {
    if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(v[0]));
  };
  put(category_ids, g->category_ids);
  put(image_order, g->image_order);
  put(stream_img_offsets, g->stream_img_offsets);
  put(frame_ids, g->frame_ids);
  put(group_offsets, g->group_offsets);
  put(sub_counts, g->sub_counts);
  if (packed && g->packable && g->n_rows) memcpy(packed, g->packed.get(), (size_t)g->n_rows * 8);
  if (rows && g->n_rows) {
    // gather from the parsed files, group ranges dealt to a few threads
    const int K = g->n_files;
    const int64_t G = (int64_t)g->group_offsets.size() - 1;
    int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int>(usable_cores(), 16),
                                                         g->n_rows / 100000));
    auto job = [&](int t) {
      const int64_t q0 = G * t / T, q1 = G * (t + 1) / T;
      int64_t at = g->group_offsets[(size_t)q0];
      for (int64_t q = q0; q < q1; q++)
        for (int f = 0; f < K; f++) {
          const w2t_json_dets &d = *g->files[f];
          const double *ws = g->wscore[f].data();
          for (int32_t c = g->sub_counts[(size_t)(q * K + f)]; c > 0; c--, at++) {
            const size_t i = (size_t)g->src[(size_t)at];
            double *r = rows + 5 * at;
            r[0] = ws[i];
            r[1] = d.bbox[4 * i]; r[2] = d.bbox[4 * i + 1]; r[3] = d.bbox[4 * i + 2]; r[4] = d.bbox[4 * i + 3];
          }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < T; t++) pool.emplace_back(job, t);
    job(0);
    for (auto &th : pool) th.join();
  }
  return W2T_OK;
}

extern "C" const char *w2t_json_groups_image_ids(const w2t_json_groups_t *g, int64_t *bytes) {
  if (!g) return nullptr;
  if (bytes) *bytes = (int64_t)g->names.size();
  return g->names.data();
}

extern "C" void w2t_json_groups_free(w2t_json_groups_t *g) { delete g; }

// ---- the tracker's input from one file: read_data_file + the packing of its result (tracking/utils.py:63-96) ----

struct w2t_json_tracks {
  std::string stream_names;                 // "segment\tcamera\n" per stream
  std::vector<int32_t> stream_img_offsets;  // [S+1]
  std::vector<int64_t> frame_ids;           // [n_img] in layout order
  std::vector<int32_t> det_start, det_count;  // [n_img * n_classes]
  std::vector<float> det_box;               // [N,4] x1, y1, x2, y2 rounded to float32 (tracker_sort.py:45)
  std::vector<int32_t> class_rank;          // [S * n_classes] position of the category in the stream's tracker dict
};

extern "C" int w2t_json_pack_tracks(const char *path, const double *score_thr, int32_t n_thr, int32_t n_classes,
                                    const char *segment_id, int32_t block_rank, int32_t block_world,
                                    w2t_json_tracks_t **out) {
  if (!path || !score_thr || n_thr < 0 || n_classes < 1 || n_classes > W2T_MAX_CLASSES || !out || block_world < 0 ||
      (block_world > 1 && (block_rank < 0 || block_rank >= block_world))) {
    w2t::set_last_error("w2t_json_pack_tracks: bad argument");
    return W2T_ERR_ARG;
  }
  *out = nullptr;
  Lap lap;
  std::string error;
  std::unique_ptr<w2t_json_dets> file(load_file(path, error));
  if (!file) {
    w2t::set_last_error("w2t_json_load: %s", error.c_str());
    return W2T_ERR_ARG;
  }
  lap("parse");
  const w2t_json_dets &d = *file;
  const int NC = n_classes;
  const int n_ids = (int)d.image_ids.size();
  const size_t n = d.score.size();
  // image id -> (segment, frame, camera); segments and cameras numbered in first-appearance order
  std::unordered_map<std::string_view, int32_t> seg_index, cam_index;
  std::vector<std::string_view> segments, cameras;
  std::vector<int32_t> seg_of((size_t)n_ids), cam_of((size_t)n_ids);
  std::vector<int64_t> frame_of((size_t)n_ids);
  for (int i = 0; i < n_ids; i++) {
    const std::string_view id = d.image_ids[(size_t)i];
    const size_t a = id.find('/');
    const size_t b = a == std::string_view::npos ? a : id.find('/', a + 1);
    if (b == std::string_view::npos || id.find('/', b + 1) != std::string_view::npos ||
        !plain_int(id.substr(a + 1, b - a - 1), frame_of[(size_t)i])) {
      w2t::set_last_error("w2t_json_pack_tracks: image id %s is not segment/frame/camera", d.image_ids[(size_t)i].c_str());
      return W2T_ERR_UNSUPPORTED;  // the general path raises the reference's ValueError (or parses int('1_0'))
    }
    const std::string_view seg = id.substr(0, a), cam = id.substr(b + 1);
    auto si = seg_index.emplace(seg, (int32_t)segments.size());
    if (si.second) segments.push_back(seg);
    auto ci = cam_index.emplace(cam, (int32_t)cameras.size());
    if (ci.second) cameras.push_back(cam);
    seg_of[(size_t)i] = si.first->second;
    cam_of[(size_t)i] = ci.first->second;
  }
  const int n_cam = std::max<int>((int)cameras.size(), 1);
  // streams = (segment, camera) pairs in first-appearance order: segment first appearance, then the pair's own
  // (image ids are interned in first-appearance order, so the smallest image index of a pair is its first image)
  std::vector<int32_t> pair_first((size_t)segments.size() * n_cam, -1), seg_first(segments.size(), INT32_MAX);
  for (int i = 0; i < n_ids; i++) {
    int32_t &pf = pair_first[(size_t)seg_of[(size_t)i] * n_cam + cam_of[(size_t)i]];
    if (pf < 0) pf = i;
    seg_first[(size_t)seg_of[(size_t)i]] = std::min(seg_first[(size_t)seg_of[(size_t)i]], (int32_t)i);
  }
  std::vector<int32_t> stream_pairs;
  for (size_t p = 0; p < pair_first.size(); p++)
    if (pair_first[p] >= 0) stream_pairs.push_back((int32_t)p);
  std::sort(stream_pairs.begin(), stream_pairs.end(), [&](int32_t x, int32_t y) {
    const int32_t sx = seg_first[(size_t)(x / n_cam)], sy = seg_first[(size_t)(y / n_cam)];
    return sx != sy ? sx < sy : pair_first[(size_t)x] < pair_first[(size_t)y];
  });
  if (segment_id != nullptr) {  // track.py --segment-id
    std::vector<int32_t> kept;
    for (int32_t p : stream_pairs)
      if (segments[(size_t)(p / n_cam)] == std::string_view(segment_id)) kept.push_back(p);
    stream_pairs.swap(kept);
  }
  if (block_world > 1) {  // this rank's contiguous block of the remaining segments (sharding.block)
    std::vector<int32_t> seg_order;
    for (int32_t p : stream_pairs)
      if (std::find(seg_order.begin(), seg_order.end(), p / n_cam) == seg_order.end()) seg_order.push_back(p / n_cam);
    const int q = (int)seg_order.size() / block_world, rem = (int)seg_order.size() % block_world;
    const int lo = block_rank * q + std::min(block_rank, rem), cnt = q + (block_rank < rem ? 1 : 0);
    std::vector<char> mine(segments.size(), 0);
    for (int k = lo; k < lo + cnt; k++) mine[(size_t)seg_order[(size_t)k]] = 1;
    std::vector<int32_t> kept;
    for (int32_t p : stream_pairs)
      if (mine[(size_t)(p / n_cam)]) kept.push_back(p);
    stream_pairs.swap(kept);
  }
  const int S = (int)stream_pairs.size();
  std::vector<int32_t> stream_of_pair(pair_first.size(), -1);
  for (int sidx = 0; sidx < S; sidx++) stream_of_pair[(size_t)stream_pairs[(size_t)sidx]] = sidx;
  auto t = std::make_unique<w2t_json_tracks>();
  for (int sidx = 0; sidx < S; sidx++) {
    const int32_t p = stream_pairs[(size_t)sidx];
    t->stream_names.append(segments[(size_t)(p / n_cam)]);
    t->stream_names += '\t';
    t->stream_names.append(cameras[(size_t)(p % n_cam)]);
    t->stream_names += '\n';
  }
  // images of each stream in frame order (ties: first appearance), every image of the file whether or not a row
  // of it survives the filters
  std::vector<int32_t> img_stream((size_t)n_ids), img_order;
  for (int i = 0; i < n_ids; i++) {
    img_stream[(size_t)i] = stream_of_pair[(size_t)seg_of[(size_t)i] * n_cam + cam_of[(size_t)i]];
    if (img_stream[(size_t)i] >= 0) img_order.push_back(i);
  }
  std::stable_sort(img_order.begin(), img_order.end(), [&](int32_t x, int32_t y) {
    return img_stream[(size_t)x] != img_stream[(size_t)y] ? img_stream[(size_t)x] < img_stream[(size_t)y]
                                                          : frame_of[(size_t)x] < frame_of[(size_t)y];
  });
  const int64_t n_img = (int64_t)img_order.size();
  std::vector<int32_t> new_img((size_t)n_ids, -1);
  t->frame_ids.resize((size_t)n_img);
  t->stream_img_offsets.assign((size_t)S + 1, 0);
  for (int64_t k = 0; k < n_img; k++) {
    const int32_t i = img_order[(size_t)k];
    new_img[(size_t)i] = (int32_t)k;
    t->frame_ids[(size_t)k] = frame_of[(size_t)i];
    t->stream_img_offsets[(size_t)img_stream[(size_t)i] + 1]++;
  }
  for (int sidx = 0; sidx < S; sidx++) t->stream_img_offsets[(size_t)sidx + 1] += t->stream_img_offsets[(size_t)sidx];
  if (n_img * NC >= INT32_MAX) { w2t::set_last_error("w2t_json_pack_tracks: too many groups"); return W2T_ERR_UNSUPPORTED; }
  lap("layout");
  // filters of read_data_file in its order (utils.py:79-87): box validity, then the category's threshold
  const int64_t G = n_img * NC;
  const int cat_hi = std::min(NC, (int)n_thr);
  t->det_count.assign((size_t)G, 0);
  std::vector<int32_t> key(n);
  for (size_t i = 0; i < n; i++) {
    key[i] = -1;
    const int32_t img = new_img[(size_t)d.image_index[i]];
    if (img < 0) continue;
    if (d.bbox[4 * i + 2] < 1 || d.bbox[4 * i + 3] < 1) continue;
    const int32_t c = d.category[i];
    if (c < 1 || c > cat_hi) {
      w2t::set_last_error("w2t_json_pack_tracks: category_id %d with %d thresholds", (int)c, cat_hi);
      return W2T_ERR_UNSUPPORTED;  // the general path raises the reference's IndexError
    }
    if (d.score[i] < score_thr[c - 1]) continue;
    key[i] = img * NC + (c - 1);
    t->det_count[(size_t)key[i]]++;
  }
  t->det_start.resize((size_t)G);
  int64_t total = 0;
  for (int64_t q = 0; q < G; q++) {
    t->det_start[(size_t)q] = (int32_t)total;
    total += t->det_count[(size_t)q];
    if (total >= INT32_MAX) { w2t::set_last_error("w2t_json_pack_tracks: more than 2^31 rows"); return W2T_ERR_UNSUPPORTED; }
  }
  // scatter in file order inside a group; category rank = order of the first surviving row in (image, file) order
  t->det_box.resize((size_t)total * 4);
  std::vector<int32_t> cursor(t->det_start);
  std::vector<int64_t> pos_first((size_t)S * NC, INT64_MAX);
  for (size_t i = 0; i < n; i++) {
    const int32_t k = key[i];
    if (k < 0) continue;
    float *b = &t->det_box[(size_t)cursor[(size_t)k]++ * 4];
    const double x = d.bbox[4 * i], y = d.bbox[4 * i + 1];
    b[0] = (float)x; b[1] = (float)y; b[2] = (float)(x + d.bbox[4 * i + 2]); b[3] = (float)(y + d.bbox[4 * i + 3]);
    const int32_t img = k / NC;
    const int32_t sidx = img_stream[(size_t)img_order[(size_t)img]];
    int64_t &pf = pos_first[(size_t)sidx * NC + (k % NC)];
    pf = std::min(pf, (int64_t)img * ((int64_t)n + 1) + (int64_t)i);
  }
  t->class_rank.resize((size_t)S * NC);
  for (int sidx = 0; sidx < S; sidx++) {
    int order[W2T_MAX_CLASSES];
    for (int c = 0; c < NC; c++) order[c] = c;
    std::stable_sort(order, order + NC, [&](int x, int y) { return pos_first[(size_t)sidx * NC + x] < pos_first[(size_t)sidx * NC + y]; });
    for (int r = 0; r < NC; r++) t->class_rank[(size_t)sidx * NC + order[r]] = r;
  }
  lap("filter + scatter");
  *out = t.release();
  return W2T_OK;
}

extern "C" int w2t_json_tracks_info(const w2t_json_tracks_t *t, int64_t info[4]) {
  if (!t || !info) { w2t::set_last_error("w2t_json_tracks_info: null argument"); return W2T_ERR_ARG; }
  info[0] = (int64_t)t->stream_img_offsets.size() - 1;
  info[1] = (int64_t)t->frame_ids.size();
  info[2] = (int64_t)t->det_box.size() / 4;
  info[3] = (int64_t)t->stream_names.size();
  return W2T_OK;
}

extern "C" int w2t_json_tracks_copy(const w2t_json_tracks_t *t, int32_t *stream_img_offsets, int64_t *frame_ids,
                                    int32_t *det_start, int32_t *det_count, float *det_box, int32_t *class_rank,
                                    char *stream_names) {
  if (!t) { w2t::set_last_error("w2t_json_tracks_copy: null handle"); return W2T_ERR_ARG; }
  auto put = [](auto *dst, const auto &v) {
    if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(v[0]));
  };
  put(stream_img_offsets, t->stream_img_offsets);
  put(frame_ids, t->frame_ids);
  put(det_start, t->det_start);
  put(det_count, t->det_count);
  put(det_box, t->det_box);
  put(class_rank, t->class_rank);
  if (stream_names && !t->stream_names.empty()) memcpy(stream_names, t->stream_names.data(), t->stream_names.size());
  return W2T_OK;
}

extern "C" void w2t_json_tracks_free(w2t_json_tracks_t *t) { delete t; }

namespace {
// rows [0, n) formatted by `row(out, i)` on a few host threads (each its own range and buffer), written in order
template <class ROW>
bool write_rows_parallel(const char *path, int64_t n, ROW row) {
  const int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(usable_cores(), 32), n / 20000));
  std::vector<std::string> buf((size_t)T);
  Lap lap;
  auto work = [&](int t) {
    const int64_t i0 = n * t / T, i1 = n * (t + 1) / T;
    std::string &out = buf[t];
    out.reserve((size_t)(i1 - i0) * 256 + 16);
    if (t == 0) out += '[';
    for (int64_t i = i0; i < i1; i++) {
      if (i) out += ", ";
      row(out, i);
    }
    if (t == T - 1) out += ']';
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < T; t++) pool.emplace_back(work, t);
  work(0);
  for (auto &th : pool) th.join();
  lap("format");
  // every thread writes its own byte range of the file
  const int fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0666);
  if (fd < 0) return false;
  std::vector<size_t> at((size_t)T + 1, 0);
  for (int t = 0; t < T; t++) at[t + 1] = at[t] + buf[t].size();
  std::vector<char> good((size_t)T, 1);
  auto put = [&](int t) {
    const char *src = buf[t].data();
    size_t left = buf[t].size(), off = at[t];
    while (left) {
      const ssize_t w = ::pwrite(fd, src, left, (off_t)off);
      if (w <= 0) { good[t] = 0; return; }
      src += w; off += (size_t)w; left -= (size_t)w;
    }
  };
  pool.clear();
  for (int t = 1; t < T; t++) pool.emplace_back(put, t);
  put(0);
  for (auto &th : pool) th.join();
  bool ok = ::close(fd) == 0;
  lap("write");
  for (int t = 0; t < T; t++) ok = ok && good[t];
  return ok;
}

// the quoted image id of row i; rows of one image are consecutive, so the previous one is usually reused
std::atomic<uint64_t> g_write_call{0};
inline void put_image_id(std::string &out, uint64_t call, const char *const *image_ids, int32_t img) {
  thread_local uint64_t cached_call = 0;
  thread_local int32_t cached_img = -1;
  thread_local std::string cached;
  if (cached_call != call || cached_img != img) {
    cached.clear();
    py_string(cached, image_ids[img]);
    cached_call = call;
    cached_img = img;
  }
  out += cached;
}
}  // namespace

extern "C" int w2t_json_write_tracks(const char *path, int64_t n, const char *const *image_ids, const int32_t *image,
                                     const double *bbox, const double *score, const int32_t *category,
                                     const int64_t *object_id) {
  if (!path || n < 0 || (n > 0 && (!image_ids || !image || !bbox || !score || !category || !object_id))) {
    w2t::set_last_error("w2t_json_write_tracks: bad argument");
    return W2T_ERR_ARG;
  }
  if (n >= 40000) {
    const uint64_t call = ++g_write_call;
    const bool ok = write_rows_parallel(path, n, [&](std::string &out, int64_t i) {
      out += "{\"image_id\": ";
      put_image_id(out, call, image_ids, image[i]);
      out += ", \"bbox\": [";
      for (int k = 0; k < 4; k++) {
        if (k) out += ", ";
        py_float(out, bbox[4 * i + k]);
      }
      out += "], \"score\": ";
      py_float(out, score[i]);
      out += ", \"category_id\": ";
      put_int(out, category[i]);
      out += ", \"object_id\": \"";
      put_int(out, (long long)object_id[i]);
      out += "\"}";
    });
    if (!ok) { w2t::set_last_error("w2t_json_write_tracks: cannot write %s", path); return W2T_ERR_ARG; }
    return W2T_OK;
  }
  std::string out;
  out.reserve(1 << 22);
  out += '[';
  bool first_chunk = true;
  for (int64_t i = 0; i < n; i++) {
    if (i) out += ", ";
    out += "{\"image_id\": ";
    py_string(out, image_ids[image[i]]);
    out += ", \"bbox\": [";
    for (int k = 0; k < 4; k++) {
      if (k) out += ", ";
      py_float(out, bbox[4 * i + k]);
    }
    out += "], \"score\": ";
    py_float(out, score[i]);
    out += ", \"category_id\": ";
    put_int(out, category[i]);
    out += ", \"object_id\": \"";
    put_int(out, (long long)object_id[i]);
    out += "\"}";
    if (out.size() > (1u << 22) - 512) {
      if (!write_all(path, out, !first_chunk)) { w2t::set_last_error("w2t_json_write_tracks: cannot write %s", path); return W2T_ERR_ARG; }
      first_chunk = false;
      out.clear();
    }
  }
  out += ']';
  if (!write_all(path, out, !first_chunk)) { w2t::set_last_error("w2t_json_write_tracks: cannot write %s", path); return W2T_ERR_ARG; }
  return W2T_OK;
}

extern "C" int w2t_json_write_detections(const char *path, int64_t n, const char *const *image_ids,
                                         const int32_t *image, const int32_t *category, const int32_t *bbox,
                                         const double *score) {
  if (!path || n < 0 || (n > 0 && (!image_ids || !image || !category || !bbox || !score))) {
    w2t::set_last_error("w2t_json_write_detections: bad argument");
    return W2T_ERR_ARG;
  }
  if (n >= 40000) {
    const uint64_t call = ++g_write_call;
    const bool ok = write_rows_parallel(path, n, [&](std::string &out, int64_t i) {
      out += "{\"image_id\": ";
      put_image_id(out, call, image_ids, image[i]);
      out += ", \"category_id\": ";
      put_int(out, category[i]);
      out += ", \"bbox\": [";
      for (int k = 0; k < 4; k++) {
        if (k) out += ", ";
        put_int(out, bbox[4 * i + k]);
      }
      out += "], \"score\": ";
      py_float(out, score[i]);
      out += '}';
    });
    if (!ok) { w2t::set_last_error("w2t_json_write_detections: cannot write %s", path); return W2T_ERR_ARG; }
    return W2T_OK;
  }
  std::string out;
  out.reserve(1 << 22);
  out += '[';
  bool first_chunk = true;
  for (int64_t i = 0; i < n; i++) {
    if (i) out += ", ";
    out += "{\"image_id\": ";
    py_string(out, image_ids[image[i]]);
    out += ", \"category_id\": ";
    put_int(out, category[i]);
    out += ", \"bbox\": [";
    for (int k = 0; k < 4; k++) {
      if (k) out += ", ";
      put_int(out, bbox[4 * i + k]);
    }
    out += "], \"score\": ";
    py_float(out, score[i]);
    out += '}';
    if (out.size() > (1u << 22) - 512) {
      if (!write_all(path, out, !first_chunk)) { w2t::set_last_error("w2t_json_write_detections: cannot write %s", path); return W2T_ERR_ARG; }
      first_chunk = false;
      out.clear();
    }
  }
  out += ']';
  if (!write_all(path, out, !first_chunk)) { w2t::set_last_error("w2t_json_write_detections: cannot write %s", path); return W2T_ERR_ARG; }
  return W2T_OK;
}
