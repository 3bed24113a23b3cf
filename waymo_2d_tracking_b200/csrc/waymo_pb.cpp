// waymo_pb.cpp — Waymo Open Dataset `Submission` / `Objects` protobuf writer (host code).
//
// Replaces create_pd_object / create_pd_objects / create_pb_submission of coco_to_waymo.py:16-82: the
// consumer of the path's output JSON (README.md:46,55 of the reference).  The reference builds the messages
// with the generated classes of the `waymo_open_dataset` package, which is neither under
// /root/reference nor installable here; the wire format is written by hand from the package's
// published .proto files (waymo-open-dataset v1.2, May 2020):
//
//   label.proto       Label      { Box box = 1; Type type = 3; string id = 4; }
//                     Label.Box  { double center_x = 1, center_y = 2, center_z = 3, width = 4,
//                                  length = 5, height = 6, heading = 7; }
//   metrics.proto     Object     { Label object = 1; float score = 2; string context_name = 4;
//                                  int64 frame_timestamp_micros = 5; CameraName.Name camera_name = 6; }
//                     Objects    { repeated Object objects = 1; }
//   submission.proto  Submission { Task task = 1; string account_name = 2, unique_method_name = 3;
//                                  repeated string authors = 4; string affiliation = 5, description = 6,
//                                  method_link = 7; SensorType sensor_type = 8;
//                                  int32 number_past_frames_exclude_current = 9,
//                                  number_future_frames_exclude_current = 10;
//                                  Objects inference_results = 11; }
//   dataset.proto     CameraName.Name { UNKNOWN = 0, FRONT = 1, FRONT_LEFT = 2, FRONT_RIGHT = 3,
//                                       SIDE_LEFT = 4, SIDE_RIGHT = 5 }
//
// PARITY UNPINNED at the schema: the field numbers above are restated from the published protos and
// cannot be checked against the package in this environment.  Given the schema the bytes are those
// of the protobuf runtime (proto2 presence: every field the reference assigns is written, default or
// not; fields in field-number order) — tests/test_waymo_pb.py builds the same schema dynamically and
// compares byte for byte.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "w2t.h"

namespace w2t {
void set_last_error(const char *fmt, ...);
}

namespace {

inline void put_varint(std::string &o, uint64_t v) {
  while (v >= 0x80) {
    o += (char)((v & 0x7f) | 0x80);
    v >>= 7;
  }
  o += (char)v;
}
inline int varint_size(uint64_t v) {
  int n = 1;
  while (v >= 0x80) { v >>= 7; n++; }
  return n;
}
inline void put_double(std::string &o, int field, double v) {
  o += (char)((field << 3) | 1);
  char b[8];
  std::memcpy(b, &v, 8);
  o.append(b, 8);
}
inline void put_string(std::string &o, int field, const char *s, size_t n) {
  put_varint(o, (uint64_t)((field << 3) | 2));
  put_varint(o, n);
  o.append(s, n);
}

struct ImageKey {
  std::string context;
  int64_t timestamp;
  int camera;
};

// 'segment/timestamp/camera' (coco_to_waymo.py:58); camera by name as CameraName.Name.Value() does
bool parse_image_id(const char *id, ImageKey &k) {
  const char *a = std::strchr(id, '/');
  if (!a) return false;
  const char *b = std::strchr(a + 1, '/');
  if (!b || std::strchr(b + 1, '/')) return false;
  k.context.assign(id, a - id);
  if (b == a + 1) return false;
  char *end = nullptr;
  k.timestamp = std::strtoll(a + 1, &end, 10);
  if (end != b) return false;
  static const char *names[] = {"UNKNOWN", "FRONT", "FRONT_LEFT", "FRONT_RIGHT", "SIDE_LEFT", "SIDE_RIGHT"};
  k.camera = -1;
  for (int i = 0; i < 6; i++)
    if (std::strcmp(b + 1, names[i]) == 0) k.camera = i;
  return k.camera >= 0;
}

// one metrics.Object, appended to `o`
void put_object(std::string &o, const ImageKey &k, const double *bbox, double score, int category, const char *oid,
                size_t oid_len, std::string &label) {
  label.clear();
  // Label.Box (36 bytes): center_x, center_y, width = bbox[3], length = bbox[2] (coco_to_waymo.py:35-38)
  label += (char)0x0a;
  label += (char)36;
  put_double(label, 1, bbox[0] + bbox[2] * 0.5);
  put_double(label, 2, bbox[1] + bbox[3] * 0.5);
  put_double(label, 4, bbox[3]);
  put_double(label, 5, bbox[2]);
  label += (char)0x18;  // type
  put_varint(label, (uint64_t)(int64_t)category);
  if (oid) put_string(label, 4, oid, oid_len);
  const float sc = (float)score;
  const size_t body = 1 + varint_size(label.size()) + label.size() + 5 + 1 + varint_size(k.context.size()) + k.context.size() +
                      1 + varint_size((uint64_t)k.timestamp) + 1 + varint_size((uint64_t)k.camera);
  o += (char)0x0a;  // Objects.objects
  put_varint(o, body);
  o += (char)0x0a;  // Object.object
  put_varint(o, label.size());
  o += label;
  o += (char)0x15;  // Object.score (float)
  char b[4];
  std::memcpy(b, &sc, 4);
  o.append(b, 4);
  put_string(o, 4, k.context.data(), k.context.size());
  o += (char)0x28;
  put_varint(o, (uint64_t)k.timestamp);
  o += (char)0x30;
  put_varint(o, (uint64_t)k.camera);
}

}  // namespace

extern "C" int w2t_pb_write_submission(const char *path, int32_t objects_only, int32_t task, const char *account_name,
                                       const char *unique_method_name, const char *const *authors, int32_t n_authors,
                                       const char *affiliation, const char *description, const char *method_link,
                                       int32_t sensor_type, int64_t n, const char *const *image_ids, int64_t n_images,
                                       const int32_t *image, const double *bbox, const double *score,
                                       const int32_t *category, const int64_t *object_id) {
  if (!path || n < 0 || n_images < 0 || (n > 0 && (!image_ids || !image || !bbox || !score || !category))) {
    w2t::set_last_error("w2t_pb_write_submission: bad argument");
    return W2T_ERR_ARG;
  }
  std::vector<ImageKey> keys((size_t)n_images);
  for (int64_t i = 0; i < n_images; i++)
    if (!parse_image_id(image_ids[i], keys[(size_t)i])) {
      w2t::set_last_error("w2t_pb_write_submission: image id '%s' is not segment/timestamp/camera with a Waymo camera name",
                          image_ids[i]);
      return W2T_ERR_ARG;
    }
  std::string objects, label;
  objects.reserve((size_t)n * 96 + 16);
  char num[24];
  for (int64_t i = 0; i < n; i++) {
    if (image[i] < 0 || image[i] >= n_images || category[i] == 0) {
      w2t::set_last_error("w2t_pb_write_submission: row %lld has a bad image index or category 0 (TYPE_UNKNOWN)", (long long)i);
      return W2T_ERR_ARG;
    }
    size_t len = 0;
    if (object_id) len = (size_t)std::snprintf(num, sizeof num, "%lld", (long long)object_id[i]);
    put_object(objects, keys[(size_t)image[i]], bbox + 4 * i, score[i], category[i], object_id ? num : nullptr, len, label);
  }
  std::string head;
  if (!objects_only) {
    auto str = [&](int field, const char *s) {
      if (s) put_string(head, field, s, std::strlen(s));
    };
    head += (char)0x08;
    put_varint(head, (uint64_t)(int64_t)task);
    str(2, account_name);
    str(3, unique_method_name);
    for (int32_t a = 0; a < n_authors; a++) str(4, authors[a]);
    str(5, affiliation);
    str(6, description);
    str(7, method_link);
    head += (char)0x40;
    put_varint(head, (uint64_t)(int64_t)sensor_type);
    head += (char)0x48;  // number_past_frames_exclude_current = 0
    head += (char)0x00;
    head += (char)0x50;  // number_future_frames_exclude_current = 0
    head += (char)0x00;
    head += (char)0x5a;  // inference_results
    put_varint(head, objects.size());
  }
  FILE *f = std::fopen(path, "wb");
  if (!f) {
    w2t::set_last_error("w2t_pb_write_submission: cannot open %s", path);
    return W2T_ERR_ARG;
  }
  const bool ok = std::fwrite(head.data(), 1, head.size(), f) == head.size() &&
                  std::fwrite(objects.data(), 1, objects.size(), f) == objects.size();
  if (std::fclose(f) != 0 || !ok) {
    w2t::set_last_error("w2t_pb_write_submission: cannot write %s", path);
    return W2T_ERR_ARG;
  }
  return W2T_OK;
}
