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
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

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
};

namespace {

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
    auto it = out.lookup.find(std::string(id));
    if (it == out.lookup.end()) {
      idx = (int32_t)out.image_ids.size();
      out.image_ids.emplace_back(id);
      out.lookup.emplace(out.image_ids.back(), idx);
    } else {
      idx = it->second;
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
    // merge: image ids keep first-appearance order
    for (int t = 0; t < n; t++) {
      w2t_json_dets &o = part[t];
      std::vector<int32_t> remap(o.image_ids.size());
      for (size_t i = 0; i < o.image_ids.size(); i++) {
        auto it = out.lookup.find(o.image_ids[i]);
        if (it == out.lookup.end()) {
          const int32_t idx = (int32_t)out.image_ids.size();
          out.image_ids.push_back(o.image_ids[i]);
          out.lookup.emplace(out.image_ids.back(), idx);
          remap[i] = idx;
        } else {
          remap[i] = it->second;
        }
      }
      const size_t base = out.image_index.size();
      out.image_index.resize(base + o.image_index.size());
      for (size_t i = 0; i < o.image_index.size(); i++) out.image_index[base + i] = remap[o.image_index[i]];
      out.category.insert(out.category.end(), o.category.begin(), o.category.end());
      out.bbox.insert(out.bbox.end(), o.bbox.begin(), o.bbox.end());
      out.score.insert(out.score.end(), o.score.begin(), o.score.end());
      out.has_score.insert(out.has_score.end(), o.has_score.begin(), o.has_score.end());
    }
    p = stop[n - 1];
    return true;
  }

  static int list_threads(size_t bytes) {
    int T = (int)std::thread::hardware_concurrency();
    if (const char *e = getenv("W2T_JSON_THREADS")) T = atoi(e);
    T = std::min(T, 32);
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

bool read_file(const char *path, std::string &buf) {
  FILE *f = fopen(path, "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize(n > 0 ? (size_t)n : 0);
  const size_t got = n > 0 ? fread(&buf[0], 1, (size_t)n, f) : 0;
  fclose(f);
  return got == buf.size();
}

// Python's float.__repr__
void py_float(std::string &out, double x) {
  if (std::isnan(x)) { out += "NaN"; return; }          // json.dump spellings
  if (std::isinf(x)) { out += x > 0 ? "Infinity" : "-Infinity"; return; }
  if (x == 0) { out += std::signbit(x) ? "-0.0" : "0.0"; return; }
  char buf[48];
  const auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);  // shortest round-trip digits
  const std::string_view s(buf, r.ptr - buf);
  const size_t epos = s.find('e');
  std::string_view mant = s.substr(0, epos);
  const int exp10 = atoi(std::string(s.substr(epos + 1)).c_str());
  bool neg = false;
  if (mant[0] == '-') { neg = true; mant.remove_prefix(1); }
  std::string digits;
  for (char c : mant)
    if (c != '.') digits += c;
  if (neg) out += '-';
  if (exp10 >= -4 && exp10 < 16) {
    if (exp10 < 0) {
      out += "0.";
      out.append((size_t)(-exp10 - 1), '0');
      out += digits;
    } else if ((int)digits.size() <= exp10 + 1) {
      out += digits;
      out.append((size_t)(exp10 + 1 - (int)digits.size()), '0');
      out += ".0";
    } else {
      out.append(digits, 0, (size_t)exp10 + 1);
      out += '.';
      out.append(digits, (size_t)exp10 + 1, std::string::npos);
    }
  } else {
    out += digits[0];
    if (digits.size() > 1) { out += '.'; out.append(digits, 1, std::string::npos); }
    out += 'e';
    out += exp10 < 0 ? '-' : '+';
    const int a = exp10 < 0 ? -exp10 : exp10;
    if (a < 10) out += '0';
    out += std::to_string(a);
  }
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

extern "C" int w2t_json_load(const char *path, w2t_json_dets_t **out) {
  if (!path || !out) { w2t::set_last_error("w2t_json_load: bad argument"); return W2T_ERR_ARG; }
  *out = nullptr;
  std::string buf;
  if (!read_file(path, buf)) {
    w2t::set_last_error("w2t_json_load: cannot read %s", path);
    return W2T_ERR_ARG;
  }
  auto *h = new w2t_json_dets;
  const size_t guess = buf.size() / 90 + 16;
  h->image_index.reserve(guess); h->category.reserve(guess); h->bbox.reserve(4 * guess);
  h->score.reserve(guess); h->has_score.reserve(guess);
  Parser ps{buf.data(), buf.data() + buf.size()};
  bool ok = ps.document(*h);
  if (ok) {
    ps.ws();
    if (ps.p != ps.end) ok = ps.fail("trailing data after the JSON document");
  }
  if (!ok) {
    w2t::set_last_error("w2t_json_load: %s at byte %lld of %s", ps.err ? ps.err : "parse error",
                        (long long)(ps.p - buf.data()), path);
    delete h;
    return W2T_ERR_ARG;
  }
  h->lookup.clear();
  *out = h;
  return W2T_OK;
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

namespace {
// rows [0, n) formatted by `row(out, i)` on a few host threads (each its own range and buffer), written in order
template <class ROW>
bool write_rows_parallel(const char *path, int64_t n, ROW row) {
  int T = (int)std::thread::hardware_concurrency();
  if (const char *e = getenv("W2T_JSON_THREADS")) T = atoi(e);
  T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(T, 32), n / 20000));
  std::vector<std::string> buf((size_t)T);
  auto work = [&](int t) {
    const int64_t i0 = n * t / T, i1 = n * (t + 1) / T;
    std::string &out = buf[t];
    out.reserve((size_t)(i1 - i0) * 160 + 16);
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
  FILE *f = fopen(path, "wb");
  if (!f) return false;
  bool ok = true;
  for (int t = 0; t < T && ok; t++) ok = fwrite(buf[t].data(), 1, buf[t].size(), f) == buf[t].size();
  return (fclose(f) == 0) && ok;
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
    const bool ok = write_rows_parallel(path, n, [&](std::string &out, int64_t i) {
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
      out += std::to_string(category[i]);
      out += ", \"object_id\": \"";
      out += std::to_string((long long)object_id[i]);
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
    out += std::to_string(category[i]);
    out += ", \"object_id\": \"";
    out += std::to_string((long long)object_id[i]);
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
    const bool ok = write_rows_parallel(path, n, [&](std::string &out, int64_t i) {
      out += "{\"image_id\": ";
      py_string(out, image_ids[image[i]]);
      out += ", \"category_id\": ";
      out += std::to_string(category[i]);
      out += ", \"bbox\": [";
      for (int k = 0; k < 4; k++) {
        if (k) out += ", ";
        out += std::to_string(bbox[4 * i + k]);
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
    out += std::to_string(category[i]);
    out += ", \"bbox\": [";
    for (int k = 0; k < 4; k++) {
      if (k) out += ", ";
      out += std::to_string(bbox[4 * i + k]);
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
