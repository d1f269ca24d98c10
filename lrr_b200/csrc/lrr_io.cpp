// lrr_io.cpp — on-disk formats either side of the hot path (include/lrr_io.h).  Host only, no CUDA.
// VTK XML PolyData feature sets, MetaImage label volumes, the driver's results text and ITK transform text files.
// Reference call sites: cdcl/io/vtkPolyDataToFeaturesICPFilter.cxx:150-235 (arrays and direction assembly),
// features/vtkITKPointSetToPolyDataFilter.cxx:146-245 (writer side), rrl/itkLocationRegistration.cxx:750-841 (readers),
// :1071-1150 (cdcl <-> ITK transform convention), :1874-2030 (WriteResults).
#include "../../include/lrr_io.h"
#include <zlib.h>
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

struct io_error : std::runtime_error { using std::runtime_error::runtime_error; };

int fail(char* err, int errlen, const std::string& msg) {
  if (err && errlen > 0) { std::snprintf(err, (size_t)errlen, "%s", msg.c_str()); }
  return 1;
}

std::string slurp(const char* path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw io_error(std::string("cannot open ") + path);
  std::ostringstream s; s << f.rdbuf();
  return s.str();
}

// ------------------------------------------------------------------------------------------------ minimal XML element scanner
struct Tag {
  std::string name;
  std::map<std::string, std::string> attr;
  bool closing = false, self_closing = false;
  size_t begin = 0, end = 0;   // [begin, end) of the tag text including <>
  std::string get(const std::string& k, const std::string& d = "") const { auto it = attr.find(k); return it == attr.end() ? d : it->second; }
  bool has(const std::string& k) const { return attr.count(k) != 0; }
};

// next tag at or after pos (comments, processing instructions and <!...> skipped); false at end of text
bool next_tag(const std::string& s, size_t pos, Tag& t) {
  for (;;) {
    size_t lt = s.find('<', pos);
    if (lt == std::string::npos) return false;
    if (s.compare(lt, 4, "<!--") == 0) { size_t e = s.find("-->", lt + 4); if (e == std::string::npos) return false; pos = e + 3; continue; }
    if (lt + 1 < s.size() && (s[lt + 1] == '?' || s[lt + 1] == '!')) { size_t e = s.find('>', lt); if (e == std::string::npos) return false; pos = e + 1; continue; }
    size_t i = lt + 1;
    t = Tag(); t.begin = lt;
    if (i < s.size() && s[i] == '/') { t.closing = true; ++i; }
    size_t n0 = i;
    while (i < s.size() && !std::isspace((unsigned char)s[i]) && s[i] != '>' && s[i] != '/') ++i;
    t.name = s.substr(n0, i - n0);
    for (;;) {
      while (i < s.size() && std::isspace((unsigned char)s[i])) ++i;
      if (i >= s.size()) throw io_error("unterminated tag <" + t.name);
      if (s[i] == '>') { ++i; break; }
      if (s[i] == '/') { t.self_closing = true; ++i; continue; }
      size_t k0 = i;
      while (i < s.size() && s[i] != '=' && !std::isspace((unsigned char)s[i]) && s[i] != '>') ++i;
      std::string key = s.substr(k0, i - k0);
      while (i < s.size() && std::isspace((unsigned char)s[i])) ++i;
      if (i >= s.size() || s[i] != '=') throw io_error("attribute without value in <" + t.name);
      ++i;
      while (i < s.size() && std::isspace((unsigned char)s[i])) ++i;
      if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) throw io_error("unquoted attribute in <" + t.name);
      char q = s[i++]; size_t v0 = i;
      while (i < s.size() && s[i] != q) ++i;
      if (i >= s.size()) throw io_error("unterminated attribute in <" + t.name);
      t.attr[key] = s.substr(v0, i - v0); ++i;
    }
    t.end = i;
    return true;
  }
}

// ------------------------------------------------------------------------------------------------ base64 (concatenated padded segments)
int b64val(unsigned char c) {
  if (c >= 'A' && c <= 'Z') return c - 'A';
  if (c >= 'a' && c <= 'z') return c - 'a' + 26;
  if (c >= '0' && c <= '9') return c - '0' + 52;
  if (c == '+') return 62;
  if (c == '/') return 63;
  return -1;
}
// VTK encodes the length header and the data as separate, individually padded base64 streams (or, for some writers, as one).  Decoding
// four characters at a time and honouring '=' inside the stream yields the concatenated bytes either way.
struct B64Reader {
  const std::string& s; size_t pos, end; std::vector<unsigned char> pending;
  B64Reader(const std::string& str, size_t p, size_t e) : s(str), pos(p), end(e) {}
  bool group() {
    unsigned char c[4]; int n = 0;
    while (n < 4 && pos < end) { unsigned char ch = (unsigned char)s[pos]; if (b64val(ch) >= 0 || ch == '=') c[n++] = ch; else if (!std::isspace(ch)) return false; ++pos; }
    if (n == 0) return false;
    if (n < 4) throw io_error("truncated base64 data");
    int v0 = b64val(c[0]), v1 = b64val(c[1]);
    if (v0 < 0 || v1 < 0) throw io_error("bad base64 data");
    pending.push_back((unsigned char)((v0 << 2) | (v1 >> 4)));
    if (c[2] != '=') { int v2 = b64val(c[2]); pending.push_back((unsigned char)(((v1 & 15) << 4) | (v2 >> 2)));
      if (c[3] != '=') { int v3 = b64val(c[3]); pending.push_back((unsigned char)(((v2 & 3) << 6) | v3)); } }
    return true;
  }
  void read(unsigned char* out, size_t n) {
    size_t got = 0;
    while (got < n) {
      if (pending.empty() && !group()) throw io_error("base64 data ends early");
      size_t k = std::min(n - got, pending.size());
      std::memcpy(out + got, pending.data(), k);
      pending.erase(pending.begin(), pending.begin() + (long)k);
      got += k;
    }
  }
  // a separately padded header ends on a group boundary: drop nothing; data that follow in the same stream stay in `pending`
};
struct RawReader {
  const std::string& s; size_t pos, end;
  void read(unsigned char* out, size_t n) { if (pos + n > end) throw io_error("appended data ends early"); std::memcpy(out, s.data() + pos, n); pos += n; }
};

std::string b64encode(const unsigned char* p, size_t n) {
  static const char* T = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  std::string o; o.reserve((n + 2) / 3 * 4);
  size_t i = 0;
  for (; i + 2 < n; i += 3) { unsigned v = (p[i] << 16) | (p[i + 1] << 8) | p[i + 2]; o += T[v >> 18]; o += T[(v >> 12) & 63]; o += T[(v >> 6) & 63]; o += T[v & 63]; }
  if (i + 1 == n) { unsigned v = p[i] << 16; o += T[v >> 18]; o += T[(v >> 12) & 63]; o += "=="; }
  else if (i + 2 == n) { unsigned v = (p[i] << 16) | (p[i + 1] << 8); o += T[v >> 18]; o += T[(v >> 12) & 63]; o += T[(v >> 6) & 63]; o += '='; }
  return o;
}

// ------------------------------------------------------------------------------------------------ VTK data arrays
struct VtkType { const char* name; int size; char kind; };   // kind: f float, i signed, u unsigned
const VtkType kTypes[] = {{"Float32", 4, 'f'}, {"Float64", 8, 'f'}, {"Int8", 1, 'i'}, {"UInt8", 1, 'u'}, {"Int16", 2, 'i'}, {"UInt16", 2, 'u'},
                          {"Int32", 4, 'i'}, {"UInt32", 4, 'u'}, {"Int64", 8, 'i'}, {"UInt64", 8, 'u'}};
const VtkType& vtk_type(const std::string& n) { for (auto& t : kTypes) if (n == t.name) return t; throw io_error("unsupported DataArray type " + n); }

void swap_bytes(unsigned char* p, size_t count, int size) { if (size > 1) for (size_t i = 0; i < count; ++i) std::reverse(p + i * size, p + (i + 1) * size); }

double elem_to_double(const unsigned char* p, const VtkType& t) {
  switch (t.kind) {
    case 'f': if (t.size == 4) { float v; std::memcpy(&v, p, 4); return v; } else { double v; std::memcpy(&v, p, 8); return v; }
    case 'i': switch (t.size) { case 1: { int8_t v; std::memcpy(&v, p, 1); return v; } case 2: { int16_t v; std::memcpy(&v, p, 2); return v; }
                                case 4: { int32_t v; std::memcpy(&v, p, 4); return v; } default: { int64_t v; std::memcpy(&v, p, 8); return (double)v; } }
    default:  switch (t.size) { case 1: { uint8_t v; std::memcpy(&v, p, 1); return v; } case 2: { uint16_t v; std::memcpy(&v, p, 2); return v; }
                                case 4: { uint32_t v; std::memcpy(&v, p, 4); return v; } default: { uint64_t v; std::memcpy(&v, p, 8); return (double)v; } }
  }
}

struct FileCtx { bool big_endian = false; int header_size = 4; bool zlib = false; size_t appended_begin = 0, appended_end = 0; bool appended_b64 = false; };

template <class R> uint64_t read_header_word(R& r, const FileCtx& fc) {
  unsigned char b[8]; r.read(b, (size_t)fc.header_size);
  if (fc.big_endian) std::reverse(b, b + fc.header_size);
  if (fc.header_size == 4) { uint32_t v; std::memcpy(&v, b, 4); return v; }
  uint64_t v; std::memcpy(&v, b, 8); return v;
}

// one binary block (length header + payload, or the zlib block table + compressed blocks) -> raw bytes
template <class R> std::vector<unsigned char> read_binary_block(R& r, const FileCtx& fc) {
  std::vector<unsigned char> out;
  if (!fc.zlib) {
    uint64_t nbytes = read_header_word(r, fc);
    out.resize(nbytes);
    if (nbytes) r.read(out.data(), nbytes);
    return out;
  }
  uint64_t nblocks = read_header_word(r, fc), bsize = read_header_word(r, fc), last = read_header_word(r, fc);
  std::vector<uint64_t> csz(nblocks);
  for (auto& c : csz) c = read_header_word(r, fc);
  if (nblocks == 0) return out;
  uint64_t total = (nblocks - 1) * bsize + (last ? last : bsize);
  out.resize(total);
  uint64_t off = 0;
  std::vector<unsigned char> comp;
  for (uint64_t b = 0; b < nblocks; ++b) {
    comp.resize(csz[b]); if (csz[b]) r.read(comp.data(), csz[b]);
    uint64_t want = (b + 1 == nblocks && last) ? last : bsize;
    uLongf dl = (uLongf)want;
    if (off + want > total || uncompress(out.data() + off, &dl, comp.data(), (uLong)csz[b]) != Z_OK || dl != want) throw io_error("zlib block does not inflate to its stated size");
    off += want;
  }
  return out;
}

// values of one <DataArray> as doubles (tuple-major, as stored)
std::vector<double> read_data_array(const std::string& s, const Tag& open, const FileCtx& fc, size_t* after) {
  const VtkType& ty = vtk_type(open.get("type"));
  std::string fmt = open.get("format", "ascii");
  size_t content_begin = open.end, content_end = open.end;
  if (!open.self_closing) {
    size_t e = s.find("</DataArray", open.end);
    if (e == std::string::npos) throw io_error("unterminated <DataArray>");
    content_end = e; *after = s.find('>', e) + 1;
  } else *after = open.end;
  std::vector<double> v;
  if (fmt == "ascii") {
    const char* p = s.data() + content_begin; const char* e = s.data() + content_end;
    while (p < e) {
      while (p < e && std::isspace((unsigned char)*p)) ++p;
      if (p >= e) break;
      char* q = nullptr; double d = std::strtod(p, &q);
      if (q == p) throw io_error("bad number in ascii DataArray " + open.get("Name"));
      if (ty.kind == 'f' && ty.size == 4) d = (double)(float)d;   // the array holds floats: the text is parsed into its element type
      v.push_back(d); p = q;
    }
    return v;
  }
  std::vector<unsigned char> raw;
  if (fmt == "binary") { B64Reader r(s, content_begin, content_end); raw = read_binary_block(r, fc); }
  else if (fmt == "appended") {
    if (!fc.appended_end) throw io_error("appended DataArray without <AppendedData>");
    size_t off = (size_t)std::strtoull(open.get("offset", "0").c_str(), nullptr, 10);
    if (fc.appended_begin + off > fc.appended_end) throw io_error("appended offset beyond the data");
    if (fc.appended_b64) { B64Reader r(s, fc.appended_begin + off, fc.appended_end); raw = read_binary_block(r, fc); }
    else { RawReader r{s, fc.appended_begin + off, fc.appended_end}; raw = read_binary_block(r, fc); }
  } else throw io_error("unknown DataArray format " + fmt);
  if (raw.size() % (size_t)ty.size) throw io_error("DataArray byte count is not a multiple of its element size");
  size_t cnt = raw.size() / (size_t)ty.size;
  if (fc.big_endian) swap_bytes(raw.data(), cnt, ty.size);
  v.resize(cnt);
  for (size_t i = 0; i < cnt; ++i) v[i] = elem_to_double(raw.data() + i * ty.size, ty);
  return v;
}

}  // namespace

struct lrr_feature_file {
  int64_t n = 0;
  std::vector<double> xyz, P6, C6, strength, dirs;
  std::vector<uint8_t> shape;
  bool has_cov = false;
};

namespace {

struct Piece { int64_t n = 0; std::vector<double> points; std::map<std::string, std::pair<int, std::vector<double>>> arrays; };

void parse_vtp(const std::string& s, std::vector<Piece>& pieces) {
  FileCtx fc;
  Tag t; size_t pos = 0;
  bool seen_root = false;
  // pass 1: root attributes and the appended section (binary: must not be scanned for tags)
  size_t app = s.find("<AppendedData");
  size_t scan_end = s.size();
  if (app != std::string::npos) {
    Tag a; if (!next_tag(s, app, a)) throw io_error("bad <AppendedData>");
    size_t us = s.find('_', a.end);
    if (us == std::string::npos) throw io_error("<AppendedData> without '_' marker");
    fc.appended_begin = us + 1;
    size_t close = s.rfind("</AppendedData");
    fc.appended_end = close == std::string::npos ? s.size() : close;
    std::string enc = a.get("encoding", "base64");
    fc.appended_b64 = (enc == "base64");
    if (!fc.appended_b64 && enc != "raw") throw io_error("unknown AppendedData encoding " + enc);
    scan_end = app;
  }
  std::string head = s.substr(0, scan_end);
  Piece* cur = nullptr; std::string section;
  while (next_tag(head, pos, t)) {
    pos = t.end;
    if (t.name == "VTKFile" && !t.closing) {
      seen_root = true;
      if (t.get("type") != "PolyData") throw io_error("VTKFile type is '" + t.get("type") + "', expected PolyData");
      fc.big_endian = t.get("byte_order", "LittleEndian") == "BigEndian";
      std::string ht = t.get("header_type", "UInt32");
      if (ht == "UInt32") fc.header_size = 4; else if (ht == "UInt64") fc.header_size = 8; else throw io_error("unknown header_type " + ht);
      std::string comp = t.get("compressor", "");
      if (comp == "vtkZLibDataCompressor") fc.zlib = true; else if (!comp.empty()) throw io_error("unsupported compressor " + comp + " (zlib only)");
    } else if (t.name == "Piece") {
      if (t.closing) { cur = nullptr; continue; }
      pieces.emplace_back(); cur = &pieces.back();
      cur->n = std::strtoll(t.get("NumberOfPoints", "0").c_str(), nullptr, 10);
      if (t.self_closing) cur = nullptr;
    } else if (t.name == "PointData" || t.name == "Points" || t.name == "CellData" || t.name == "Verts" || t.name == "Lines" || t.name == "Strips" || t.name == "Polys") {
      section = (t.closing || t.self_closing) ? "" : t.name;
    } else if (t.name == "DataArray" && !t.closing) {
      if (!cur || (section != "PointData" && section != "Points")) {   // connectivity etc.: skip the content
        if (!t.self_closing) { size_t e = head.find("</DataArray", t.end); if (e == std::string::npos) throw io_error("unterminated <DataArray>"); pos = head.find('>', e) + 1; }
        continue;
      }
      size_t after = t.end;
      std::vector<double> v = read_data_array(s, t, fc, &after);
      pos = after;
      int nc = std::atoi(t.get("NumberOfComponents", "1").c_str()); if (nc < 1) nc = 1;
      if ((int64_t)v.size() != cur->n * nc) throw io_error("DataArray " + t.get("Name") + ": " + std::to_string(v.size()) + " values for " + std::to_string(cur->n) + " points x " + std::to_string(nc));
      if (section == "Points") { if (nc != 3) throw io_error("Points must have 3 components"); cur->points.swap(v); }
      else cur->arrays[t.get("Name")] = {nc, std::move(v)};
    }
  }
  if (!seen_root) throw io_error("not a VTK XML file (no <VTKFile>)");
}

void cross(const double* a, const double* b, double* c) { c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0]; }

}  // namespace

extern "C" {

int lrr_io_read_features_vtp(const char* path, lrr_feature_file** out, char* err, int errlen) {
  if (!path || !out) return fail(err, errlen, "null argument");
  try {
    std::string s = slurp(path);
    std::vector<Piece> pieces; parse_vtp(s, pieces);
    int64_t n = 0; for (auto& p : pieces) n += p.n;
    auto* f = new lrr_feature_file; f->n = n;
    f->xyz.assign(3 * n, 0); f->P6.assign(6 * n, 0); f->strength.assign(n, 0); f->shape.assign(n, 0); f->dirs.assign(9 * n, 0);
    bool all_cov = !pieces.empty();
    for (auto& p : pieces) if (p.n && !p.arrays.count("covariances")) all_cov = false;
    if (all_cov) { f->C6.assign(6 * n, 0); f->has_cov = true; }
    static const int sym[6] = {0, 1, 2, 4, 5, 8};   // xx xy xz yy yz zz of a row-major 3x3 tensor
    int64_t base = 0;
    for (auto& p : pieces) {
      if (!p.n) continue;
      if ((int64_t)p.points.size() != 3 * p.n) { delete f; return fail(err, errlen, "piece without <Points>"); }
      auto need = [&](const char* name, int nc) -> const std::vector<double>* {
        auto it = p.arrays.find(name);
        if (it == p.arrays.end()) return nullptr;
        if (it->second.first != nc) throw io_error(std::string("array ") + name + " has " + std::to_string(it->second.first) + " components, expected " + std::to_string(nc));
        return &it->second.second;
      };
      const auto* ep = need("errorProjectors", 9);   // GetTensors("errorProjectors") is dereferenced unconditionally (:178): required
      if (!ep) { delete f; return fail(err, errlen, "no point-data array `errorProjectors`"); }
      const auto* st = need("strengths", 1);          // optional, 0 when absent (:181-182)
      const auto* sh = need("shapes", 1);             // optional, 0 = corner when absent (:184-186)
      const auto* no = need("normals", 3);
      const auto* bi = need("binormals", 3);
      const auto* cv = f->has_cov ? need("covariances", 9) : nullptr;
      for (int64_t i = 0; i < p.n; ++i) {
        int64_t g = base + i;
        for (int d = 0; d < 3; ++d) f->xyz[d * n + g] = p.points[3 * i + d];
        for (int k = 0; k < 6; ++k) f->P6[k * n + g] = (*ep)[9 * i + sym[k]];
        if (cv) for (int k = 0; k < 6; ++k) f->C6[k * n + g] = (*cv)[9 * i + sym[k]];
        f->strength[g] = st ? (*st)[i] : 0.0;
        unsigned char shc = sh ? (unsigned char)(*sh)[i] : 0;
        if (shc > 2) throw io_error("shape value " + std::to_string((int)shc) + " at point " + std::to_string(g) + " (0 corner, 1 tube, 2 sheet)");
        f->shape[g] = shc;
        if (!no || (shc != 2 && !bi)) throw io_error("shape needs `normals`" + std::string(shc != 2 ? " and `binormals`" : ""));
        const double* nn = &(*no)[3 * i]; double bb[3] = {0, 0, 0};
        if (bi) std::memcpy(bb, &(*bi)[3 * i], sizeof bb);
        double d[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (shc == 0) { cross(nn, bb, d); std::memcpy(d + 3, bb, sizeof bb); std::memcpy(d + 6, nn, 3 * sizeof(double)); }       // tangent, binormal, normal (:193-203)
        else if (shc == 1) { std::memcpy(d, bb, sizeof bb); std::memcpy(d + 3, nn, 3 * sizeof(double)); }                         // binormal, normal (:207-214)
        else std::memcpy(d, nn, 3 * sizeof(double));                                                                              // normal (:218-221)
        for (int k = 0; k < 9; ++k) f->dirs[k * n + g] = d[k];
      }
      base += p.n;
    }
    *out = f;
    return 0;
  } catch (const std::exception& e) { return fail(err, errlen, std::string(path) + ": " + e.what()); }
}

void lrr_io_features_free(lrr_feature_file* f) { delete f; }
int64_t lrr_io_features_count(const lrr_feature_file* f) { return f ? f->n : 0; }
const double* lrr_io_features_xyz(const lrr_feature_file* f) { return f->xyz.data(); }
const double* lrr_io_features_P6(const lrr_feature_file* f) { return f->P6.data(); }
const double* lrr_io_features_C6(const lrr_feature_file* f) { return f->has_cov ? f->C6.data() : nullptr; }
const double* lrr_io_features_strength(const lrr_feature_file* f) { return f->strength.data(); }
const uint8_t* lrr_io_features_shape(const lrr_feature_file* f) { return f->shape.data(); }
const double* lrr_io_features_dirs(const lrr_feature_file* f) { return f->dirs.data(); }
void lrr_io_features_shape_counts(const lrr_feature_file* f, int64_t counts[3]) {
  counts[0] = counts[1] = counts[2] = 0;
  for (uint8_t s : f->shape) ++counts[s];
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ VTP writer
namespace {

struct OutArray { std::string name, type; int nc; std::vector<unsigned char> bytes; std::vector<double> values; };

OutArray make_array(const std::string& name, int nc, const std::vector<double>& v, bool f32) {
  OutArray a; a.name = name; a.nc = nc; a.values = v; a.type = f32 ? "Float32" : "Float64";
  a.bytes.resize(v.size() * (f32 ? 4 : 8));
  for (size_t i = 0; i < v.size(); ++i) { if (f32) { float x = (float)v[i]; std::memcpy(&a.bytes[4 * i], &x, 4); } else std::memcpy(&a.bytes[8 * i], &v[i], 8); }
  return a;
}

std::string header_words(const std::vector<uint32_t>& w) { return std::string((const char*)w.data(), w.size() * 4); }

// raw payload of one block in the file's binary convention (UInt32 headers, little endian)
void encode_block(const OutArray& a, bool zlib_on, std::string& header, std::string& payload) {
  if (!zlib_on) { header = header_words({(uint32_t)a.bytes.size()}); payload.assign((const char*)a.bytes.data(), a.bytes.size()); return; }
  const size_t bs = 32768;
  size_t nb = (a.bytes.size() + bs - 1) / bs, last = a.bytes.size() % bs;
  std::vector<uint32_t> h = {(uint32_t)nb, (uint32_t)bs, (uint32_t)last};
  payload.clear();
  for (size_t b = 0; b < nb; ++b) {
    size_t len = (b + 1 == nb && last) ? last : bs;
    uLongf cl = compressBound((uLong)len); std::vector<unsigned char> c(cl);
    if (compress2(c.data(), &cl, a.bytes.data() + b * bs, (uLong)len, Z_DEFAULT_COMPRESSION) != Z_OK) throw io_error("zlib compress failed");
    h.push_back((uint32_t)cl); payload.append((const char*)c.data(), cl);
  }
  header = header_words(h);
}

}  // namespace

extern "C" int lrr_io_write_features_vtp(const char* path, int64_t n, const double* xyz, const double* P6, const double* C6, const double* strength,
                                          const uint8_t* shape, const double* dirs, int mode, int as_float32, char* err, int errlen) {
  if (!path || n < 0 || (n && (!xyz || !P6 || !strength || !shape || !dirs))) return fail(err, errlen, "null argument");
  if (mode < 0 || mode > 3) return fail(err, errlen, "mode must be 0..3");
  try {
    bool f32 = as_float32 != 0;
    std::vector<double> pts(3 * n), ep(9 * n), cv, st(n), no(3 * n), bi(3 * n);
    static const int full[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
    if (C6) cv.resize(9 * n);
    OutArray sh; sh.name = "shapes"; sh.type = "UInt8"; sh.nc = 1; sh.bytes.resize(n); sh.values.resize(n);
    for (int64_t i = 0; i < n; ++i) {
      for (int d = 0; d < 3; ++d) pts[3 * i + d] = xyz[d * n + i];
      for (int k = 0; k < 9; ++k) { ep[9 * i + k] = P6[full[k] * n + i]; if (C6) cv[9 * i + k] = C6[full[k] * n + i]; }
      st[i] = strength[i];
      if (shape[i] > 2) throw io_error("shape value out of range");
      sh.bytes[i] = shape[i]; sh.values[i] = shape[i];
      int in = shape[i] == 0 ? 2 : shape[i] == 1 ? 1 : 0, ib = shape[i] == 0 ? 1 : shape[i] == 1 ? 0 : -1;
      for (int d = 0; d < 3; ++d) { no[3 * i + d] = dirs[(3 * in + d) * n + i]; bi[3 * i + d] = ib < 0 ? 0.0 : dirs[(3 * ib + d) * n + i]; }
    }
    std::vector<OutArray> pd;
    pd.push_back(make_array("strengths", 1, st, f32)); pd.push_back(std::move(sh));
    pd.push_back(make_array("normals", 3, no, f32)); pd.push_back(make_array("binormals", 3, bi, f32));
    if (C6) pd.push_back(make_array("covariances", 9, cv, f32));
    pd.push_back(make_array("errorProjectors", 9, ep, f32));
    OutArray points = make_array("Points", 3, pts, f32);
    bool zl = mode >= 2, appended = (mode == 1 || mode == 3), app_b64 = mode == 3;
    std::ostringstream o; std::string app;
    o << "<?xml version=\"1.0\"?>\n<VTKFile type=\"PolyData\" version=\"0.1\" byte_order=\"LittleEndian\"" << (mode ? " header_type=\"UInt32\"" : "")
      << (zl ? " compressor=\"vtkZLibDataCompressor\"" : "") << ">\n  <PolyData>\n    <Piece NumberOfPoints=\"" << n
      << "\" NumberOfVerts=\"0\" NumberOfLines=\"0\" NumberOfStrips=\"0\" NumberOfPolys=\"0\">\n";
    auto emit = [&](const OutArray& a, const char* ind) {
      o << ind << "<DataArray type=\"" << a.type << "\" Name=\"" << a.name << "\" NumberOfComponents=\"" << a.nc << "\" format=\"" << (mode == 0 ? "ascii" : appended ? "appended" : "binary") << "\"";
      if (mode == 0) {
        o << ">\n" << ind << "  " << std::setprecision(a.type == "Float32" ? 9 : 17);
        for (size_t i = 0; i < a.values.size(); ++i) { double v = a.type == "Float32" ? (double)(float)a.values[i] : a.values[i]; o << v << ((i + 1) % 9 == 0 ? "\n" + std::string(ind) + "  " : " "); }
        o << "\n" << ind << "</DataArray>\n"; return;
      }
      std::string h, p; encode_block(a, zl, h, p);
      if (appended) {
        o << " offset=\"" << app.size() << "\"/>\n";
        if (app_b64) { app += b64encode((const unsigned char*)h.data(), h.size()); app += b64encode((const unsigned char*)p.data(), p.size()); } else { app += h; app += p; }
      } else o << ">\n" << ind << "  " << b64encode((const unsigned char*)h.data(), h.size()) << b64encode((const unsigned char*)p.data(), p.size()) << "\n" << ind << "</DataArray>\n";
    };
    o << "      <PointData Scalars=\"shapes\" Vectors=\"binormals\" Tensors=\"errorProjectors\">\n";
    for (auto& a : pd) emit(a, "        ");
    o << "      </PointData>\n      <CellData>\n      </CellData>\n      <Points>\n";
    emit(points, "        ");
    o << "      </Points>\n    </Piece>\n  </PolyData>\n";
    if (appended) o << "  <AppendedData encoding=\"" << (app_b64 ? "base64" : "raw") << "\">\n   _" << app << "\n  </AppendedData>\n";
    o << "</VTKFile>\n";
    std::ofstream f(path, std::ios::binary); if (!f) throw io_error("cannot create file");
    std::string text = o.str(); f.write(text.data(), (std::streamsize)text.size());
    if (!f) throw io_error("write failed");
    return 0;
  } catch (const std::exception& e) { return fail(err, errlen, std::string(path) + ": " + e.what()); }
}

// ------------------------------------------------------------------------------------------------ MetaImage
struct lrr_label_volume { int64_t dims[3] = {0, 0, 0}; double spacing[3] = {1, 1, 1}, origin[3] = {0, 0, 0}; std::vector<int64_t> data; };

namespace {
struct MetType { const char* name; int size; bool is_signed; };
const MetType kMet[] = {{"MET_UCHAR", 1, false}, {"MET_CHAR", 1, true}, {"MET_USHORT", 2, false}, {"MET_SHORT", 2, true}, {"MET_UINT", 4, false},
                        {"MET_INT", 4, true}, {"MET_ULONG_LONG", 8, false}, {"MET_LONG_LONG", 8, true}};
const MetType& met_type(const std::string& n) { for (auto& t : kMet) if (n == t.name) return t; throw io_error("unsupported ElementType " + n + " (integer label volumes only)"); }
std::string trim(const std::string& s) { size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n"); return a == std::string::npos ? "" : s.substr(a, b - a + 1); }
std::string dir_of(const std::string& p) { size_t k = p.find_last_of('/'); return k == std::string::npos ? "" : p.substr(0, k + 1); }
}  // namespace

extern "C" {

int lrr_io_read_mhd(const char* path, lrr_label_volume** out, char* err, int errlen) {
  if (!path || !out) return fail(err, errlen, "null argument");
  try {
    std::string s = slurp(path);
    auto v = std::unique_ptr<lrr_label_volume>(new lrr_label_volume);
    std::string etype, dfile; bool msb = false, compressed = false; int ndims = 3; size_t pos = 0, data_off = 0; bool local = false;
    while (pos < s.size()) {
      size_t nl = s.find('\n', pos); if (nl == std::string::npos) nl = s.size();
      std::string line = s.substr(pos, nl - pos); pos = nl + 1;
      size_t eq = line.find('='); if (eq == std::string::npos) continue;
      std::string k = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
      std::istringstream vs(val);
      if (k == "NDims") vs >> ndims;
      else if (k == "DimSize") for (int d = 0; d < 3 && vs; ++d) vs >> v->dims[d];
      else if (k == "ElementSpacing" || k == "ElementSize") { double t[3] = {1, 1, 1}; for (int d = 0; d < 3; ++d) vs >> t[d]; if (k == "ElementSpacing" || v->spacing[0] == 1) std::memcpy(v->spacing, t, sizeof t); }
      else if (k == "Offset" || k == "Position" || k == "Origin") for (int d = 0; d < 3; ++d) vs >> v->origin[d];
      else if (k == "ElementType") etype = val;
      else if (k == "BinaryDataByteOrderMSB" || k == "ElementByteOrderMSB") msb = (val == "True" || val == "true");
      else if (k == "CompressedData") compressed = (val == "True" || val == "true");
      else if (k == "ElementDataFile") { dfile = val; if (val == "LOCAL") { local = true; data_off = pos; } break; }   // always the last header line
    }
    if (ndims != 3) throw io_error("NDims must be 3");
    if (dfile.empty()) throw io_error("no ElementDataFile");
    const MetType& mt = met_type(etype);
    int64_t count = v->dims[0] * v->dims[1] * v->dims[2];
    if (count <= 0) throw io_error("bad DimSize");
    std::string raw = local ? s.substr(data_off) : slurp((dfile[0] == '/' ? dfile : dir_of(path) + dfile).c_str());
    std::vector<unsigned char> buf((size_t)count * mt.size);
    if (compressed) { uLongf dl = (uLongf)buf.size(); if (uncompress(buf.data(), &dl, (const Bytef*)raw.data(), (uLong)raw.size()) != Z_OK || dl != buf.size()) throw io_error("compressed data do not inflate to DimSize x ElementType"); }
    else { if (raw.size() < buf.size()) throw io_error("data file shorter than DimSize x ElementType"); std::memcpy(buf.data(), raw.data(), buf.size()); }
    if (msb) swap_bytes(buf.data(), (size_t)count, mt.size);
    v->data.resize((size_t)count);
    for (int64_t i = 0; i < count; ++i) {
      const unsigned char* p = buf.data() + i * mt.size; int64_t x = 0;
      switch (mt.size) {
        case 1: x = mt.is_signed ? (int64_t)(int8_t)p[0] : (int64_t)p[0]; break;
        case 2: { uint16_t u; std::memcpy(&u, p, 2); x = mt.is_signed ? (int64_t)(int16_t)u : (int64_t)u; break; }
        case 4: { uint32_t u; std::memcpy(&u, p, 4); x = mt.is_signed ? (int64_t)(int32_t)u : (int64_t)u; break; }
        default: { uint64_t u; std::memcpy(&u, p, 8); x = (int64_t)u; }
      }
      v->data[(size_t)i] = x;
    }
    *out = v.release();
    return 0;
  } catch (const std::exception& e) { return fail(err, errlen, std::string(path) + ": " + e.what()); }
}

void lrr_io_volume_free(lrr_label_volume* v) { delete v; }
void lrr_io_volume_geometry(const lrr_label_volume* v, int64_t dims[3], double spacing[3], double origin[3]) {
  for (int d = 0; d < 3; ++d) { dims[d] = v->dims[d]; spacing[d] = v->spacing[d]; origin[d] = v->origin[d]; }
}
const int64_t* lrr_io_volume_data(const lrr_label_volume* v) { return v->data.data(); }

int lrr_io_labels_at(const lrr_label_volume* v, int64_t n, const double* xyz, int32_t outside, int32_t* labels_out) {
  if (!v || !xyz || !labels_out || n < 0) return 1;
  for (int64_t i = 0; i < n; ++i) {
    int64_t idx[3]; bool in = true;
    for (int d = 0; d < 3; ++d) {
      // itk::Image::TransformPhysicalPointToIndex rounds half up: floor(c + 0.5)
      idx[d] = (int64_t)std::floor((xyz[d * n + i] - v->origin[d]) / v->spacing[d] + 0.5);
      if (idx[d] < 0 || idx[d] >= v->dims[d]) in = false;
    }
    labels_out[i] = in ? (int32_t)v->data[(size_t)((idx[2] * v->dims[1] + idx[1]) * v->dims[0] + idx[0])] : outside;
  }
  return 0;
}

int lrr_io_write_mhd(const char* path, const int64_t dims[3], const double spacing[3], const double origin[3], const int64_t* data,
                     const char* element_type, int compressed, char* err, int errlen) {
  if (!path || !dims || !data || !element_type) return fail(err, errlen, "null argument");
  try {
    const MetType& mt = met_type(element_type);
    std::string p(path), stem = p.size() > 4 && p.compare(p.size() - 4, 4, ".mhd") == 0 ? p.substr(0, p.size() - 4) : p;
    std::string dname = stem + (compressed ? ".zraw" : ".raw"), dbase = dname.substr(dir_of(dname).size());
    int64_t count = dims[0] * dims[1] * dims[2];
    std::vector<unsigned char> buf((size_t)count * mt.size);
    for (int64_t i = 0; i < count; ++i) { uint64_t u = (uint64_t)data[i]; std::memcpy(buf.data() + i * mt.size, &u, (size_t)mt.size); }   // little endian host
    std::vector<unsigned char> z; size_t zn = 0;
    if (compressed) { uLongf cl = compressBound((uLong)buf.size()); z.resize(cl); if (compress2(z.data(), &cl, buf.data(), (uLong)buf.size(), Z_DEFAULT_COMPRESSION) != Z_OK) throw io_error("zlib compress failed"); zn = cl; }
    std::ofstream h(path); if (!h) throw io_error("cannot create header");
    h << std::setprecision(17) << "ObjectType = Image\nNDims = 3\nBinaryData = True\nBinaryDataByteOrderMSB = False\nCompressedData = " << (compressed ? "True" : "False") << "\n";
    if (compressed) h << "CompressedDataSize = " << zn << "\n";
    h << "TransformMatrix = 1 0 0 0 1 0 0 0 1\nOffset = " << (origin ? origin[0] : 0) << " " << (origin ? origin[1] : 0) << " " << (origin ? origin[2] : 0)
      << "\nCenterOfRotation = 0 0 0\nAnatomicalOrientation = RAI\nElementSpacing = " << (spacing ? spacing[0] : 1) << " " << (spacing ? spacing[1] : 1) << " " << (spacing ? spacing[2] : 1)
      << "\nDimSize = " << dims[0] << " " << dims[1] << " " << dims[2] << "\nElementType = " << mt.name << "\nElementDataFile = " << dbase << "\n";
    std::ofstream d(dname, std::ios::binary); if (!d) throw io_error("cannot create data file");
    if (compressed) d.write((const char*)z.data(), (std::streamsize)zn); else d.write((const char*)buf.data(), (std::streamsize)buf.size());
    if (!h || !d) throw io_error("write failed");
    return 0;
  } catch (const std::exception& e) { return fail(err, errlen, std::string(path) + ": " + e.what()); }
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ results text and transforms
namespace {

// ascending eigenvalues of a symmetric 3x3 (cyclic Jacobi) — vnl_symmetric_eigensystem_compute's order (:1952-1956)
void sym3_evalues(const double* m9, double* e3) {
  double a[3][3]; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = 0.5 * (m9[3 * i + j] + m9[3 * j + i]);
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
      if (a[p][q] == 0) continue;
      double th = (a[q][q] - a[p][p]) / (2 * a[p][q]), t = (th >= 0 ? 1 : -1) / (std::fabs(th) + std::sqrt(th * th + 1)), c = 1 / std::sqrt(t * t + 1), s = t * c;
      for (int k = 0; k < 3; ++k) { double x = a[k][p], y = a[k][q]; a[k][p] = c * x - s * y; a[k][q] = s * x + c * y; }
      for (int k = 0; k < 3; ++k) { double x = a[p][k], y = a[q][k]; a[p][k] = c * x - s * y; a[q][k] = s * x + c * y; }
    }
  }
  e3[0] = a[0][0]; e3[1] = a[1][1]; e3[2] = a[2][2]; std::sort(e3, e3 + 3);
}

void put_matrix(std::ostream& o, const double* m, int rows, int cols) {   // vnl_matrix operator<<: values separated by one blank, one row per line
  for (int r = 0; r < rows; ++r) { for (int c = 0; c < cols; ++c) o << (c ? " " : "") << m[r * cols + c]; o << "\n"; }
}
void put_vnl3(std::ostream& o, const double* v) { o << v[0] << " " << v[1] << " " << v[2]; }          // vnl_vector_fixed operator<<
void put_itk3(std::ostream& o, const double* v) { o << "[" << v[0] << ", " << v[1] << ", " << v[2] << "]"; }   // itk::Point operator<<

bool read_doubles(std::istream& in, double* out, int n) {
  for (int i = 0; i < n; ++i) {
    std::string tok; if (!(in >> tok)) return false;
    tok.erase(std::remove_if(tok.begin(), tok.end(), [](char c) { return c == '[' || c == ']' || c == ','; }), tok.end());
    if (tok.empty()) { --i; continue; }
    char* q = nullptr; out[i] = std::strtod(tok.c_str(), &q); if (q == tok.c_str()) return false;
  }
  return true;
}

}  // namespace

extern "C" {

void lrr_io_results_t_centered(const double* Ai, const double* ti, const double* ci, const double* f, const double* m, double* out) {
  for (int r = 0; r < 3; ++r) {
    double af = 0, ac = 0;
    for (int k = 0; k < 3; ++k) { af += Ai[3 * r + k] * f[k]; ac += Ai[3 * r + k] * ci[k]; }
    out[r] = ti[r] + af - ac - m[r];
  }
}

int lrr_io_results_name(char* out, int outlen, const char* dir, const char* prefix, const char* stem, unsigned k, unsigned candidate, const char* ext) {
  if (!out || outlen <= 0) return 1;
  int n = std::snprintf(out, (size_t)outlen, "%s//%s%s%06u_%02u%s", dir ? dir : "", prefix ? prefix : "", stem ? stem : "results", k, candidate, ext ? ext : ".txt");
  return n < 0 || n >= outlen;
}

int lrr_io_write_results(const char* path, const lrr_results_record* r, int precision, char* err, int errlen) {
  if (!path || !r) return fail(err, errlen, "null argument");
  std::ofstream o(path);
  if (!o) return fail(err, errlen, std::string("cannot create ") + path);
  if (precision > 0) o << std::setprecision(precision);
  double ev[3]; sym3_evalues(r->covarianceJ, ev);
  o << "\n" << "Final estimated transform: " << "\n"; put_matrix(o, r->A, 3, 3); o << "\n"; put_vnl3(o, r->t_centered); o << "\n\n";
  o << "Final estimated covariance: " << "\n"; put_matrix(o, r->covariance, 12, 12); o << "\n\n";
  o << "Final estimated covarianceJ from moving points: " << "\n"; put_matrix(o, r->covarianceJ, 3, 3); o << "\n";
  o << "covarianceJ from points evalues: " << ev[0] << "  " << ev[1] << "  " << ev[2] << "\n\n";
  o << "Fixed keypoint location: "; put_vnl3(o, r->fixed_keypoint); o << "\n";
  o << "Moving keypoint location: "; put_vnl3(o, r->moving_keypoint); o << "\n";
  o << "Converged: " << (r->converged ? 1 : 0) << "\n";
  o << "Initial RMS: " << r->initial_rms << "\n";
  o << "Final RMS: " << r->final_rms << "\n";
  o << "RMS reduced: " << ((r->initial_rms > r->final_rms) ? 1 : 0) << "\n";
  o << "Query point: "; put_itk3(o, r->query_point); o << "\n";
  o << "Query point mapped: "; put_itk3(o, r->query_point_mapped); o << "\n";
  o << "Final alignment error : Maximum trace : Sheet angles (deg): Tube angles (deg)" << "\n";
  o << "Classification:\t";
  for (int i = 0; i < 8; ++i) o << r->measurements[i] << "\t";
  o << "\n";
  o.close();
  return o ? 0 : fail(err, errlen, std::string("write failed: ") + path);
}

int lrr_io_read_results(const char* path, lrr_results_record* r, char* err, int errlen) {
  if (!path || !r) return fail(err, errlen, "null argument");
  std::ifstream in(path);
  if (!in) return fail(err, errlen, std::string("cannot open ") + path);
  std::memset(r, 0, sizeof *r);
  std::string line; int found = 0;
  auto starts = [&](const char* p) { return line.compare(0, std::strlen(p), p) == 0; };
  auto after = [&](const char* p) { return std::istringstream(line.substr(std::strlen(p))); };
  while (std::getline(in, line)) {
    if (starts("Final estimated transform:")) { if (!read_doubles(in, r->A, 9) || !read_doubles(in, r->t_centered, 3)) return fail(err, errlen, "bad transform section"); found |= 1; }
    else if (starts("Final estimated covarianceJ from moving points:")) { if (!read_doubles(in, r->covarianceJ, 9)) return fail(err, errlen, "bad covarianceJ section"); found |= 4; }
    else if (starts("Final estimated covariance:")) { if (!read_doubles(in, r->covariance, 144)) return fail(err, errlen, "bad covariance section"); found |= 2; }
    else if (starts("covarianceJ from points evalues:")) { auto s = after("covarianceJ from points evalues:"); read_doubles(s, r->covarianceJ_evalues, 3); }
    else if (starts("Fixed keypoint location:")) { auto s = after("Fixed keypoint location:"); read_doubles(s, r->fixed_keypoint, 3); }
    else if (starts("Moving keypoint location:")) { auto s = after("Moving keypoint location:"); read_doubles(s, r->moving_keypoint, 3); }
    else if (starts("Converged:")) { auto s = after("Converged:"); int c = 0; s >> c; r->converged = c; found |= 8; }
    else if (starts("Initial RMS:")) { auto s = after("Initial RMS:"); s >> r->initial_rms; }
    else if (starts("Final RMS:")) { auto s = after("Final RMS:"); s >> r->final_rms; }
    else if (starts("Query point mapped:")) { auto s = after("Query point mapped:"); read_doubles(s, r->query_point_mapped, 3); }
    else if (starts("Query point:")) { auto s = after("Query point:"); read_doubles(s, r->query_point, 3); }
    else if (starts("Classification:")) { auto s = after("Classification:"); if (!read_doubles(s, r->measurements, 8)) return fail(err, errlen, "bad Classification line"); found |= 16; }
  }
  if ((found & 1) == 0) return fail(err, errlen, std::string(path) + ": no 'Final estimated transform' section");
  return 0;
}

int lrr_io_write_transform(const char* path, const double* A9, const double* t3, const double* c3, char* err, int errlen) {
  if (!path || !A9 || !t3 || !c3) return fail(err, errlen, "null argument");
  std::ofstream o(path);
  if (!o) return fail(err, errlen, std::string("cannot create ") + path);
  o << std::setprecision(17) << "#Insight Transform File V1.0\n#Transform 0\nTransform: CenteredAffineTransform_double_3_3\nParameters:";
  for (int i = 0; i < 9; ++i) o << " " << A9[i];
  for (int i = 0; i < 3; ++i) o << " " << c3[i];
  for (int i = 0; i < 3; ++i) o << " " << (t3[i] - c3[i]);     // T = t - C (:1138-1146)
  o << "\nFixedParameters:";
  for (int i = 0; i < 3; ++i) o << " " << c3[i];
  o << "\n";
  o.close();
  return o ? 0 : fail(err, errlen, std::string("write failed: ") + path);
}

int lrr_io_read_transform(const char* path, double* A9, double* t3, double* c3, char* err, int errlen) {
  if (!path || !A9 || !t3 || !c3) return fail(err, errlen, "null argument");
  std::ifstream in(path);
  if (!in) return fail(err, errlen, std::string("cannot open ") + path);
  std::string line, kind; std::vector<double> par, fixed;
  while (std::getline(in, line)) {
    if (line.empty() || line[0] == '#') continue;
    size_t c = line.find(':'); if (c == std::string::npos) continue;
    std::string k = trim(line.substr(0, c)); std::istringstream v(line.substr(c + 1));
    if (k == "Transform") { if (!kind.empty()) break; v >> kind; }     // first transform of the file only
    else if (k == "Parameters") { double d; while (v >> d) par.push_back(d); }
    else if (k == "FixedParameters") { double d; while (v >> d) fixed.push_back(d); }
  }
  bool centered = kind.find("CenteredAffineTransform") == 0, affine = kind.find("AffineTransform") == 0 || kind.find("MatrixOffsetTransformBase") == 0;
  if (centered && par.size() == 15) {                                   // A, centre, translation (:1076-1106)
    for (int i = 0; i < 9; ++i) A9[i] = par[(size_t)i];
    for (int i = 0; i < 3; ++i) { c3[i] = par[9 + (size_t)i]; t3[i] = par[12 + (size_t)i] + c3[i]; }
    return 0;
  }
  if (affine && par.size() == 12) {                                     // A, translation; centre in FixedParameters
    for (int i = 0; i < 9; ++i) A9[i] = par[(size_t)i];
    for (int i = 0; i < 3; ++i) { c3[i] = fixed.size() >= 3 ? fixed[(size_t)i] : 0.0; t3[i] = par[9 + (size_t)i] + c3[i]; }
    return 0;
  }
  return fail(err, errlen, std::string(path) + ": transform '" + kind + "' with " + std::to_string(par.size()) + " parameters is not a 3-D (centered) affine transform");
}

}  // extern "C"
