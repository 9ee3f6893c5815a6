// swim_codec.cpp — wire codec of the reference's Types.hs, host side.
//
//   Envelope framing      Types.hs:96-119  (cereal: putWord8 / putWord16be / isolate)
//   Message body          Types.hs:147-155 (aeson generic ToJSON -> msgpack-aeson packAeson)
//   type indices          Types.hs:159-178
//
// The body is the msgpack form of aeson's default generic encoding of a sum of records:
// a map {"tag": <constructor>, <field>: <value>, ...}. aeson keeps objects in a HashMap, so
// the reference's key order is unspecified: this encoder emits "tag" first and then the
// fields in declaration order; the decoder accepts any order and ignores unknown keys.
// Integers use the shortest msgpack form (msgpack-1.0.0 `putInt`), strings fixstr/str8/
// str16, `payload :: [Word8]` is a msgpack array of integers.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/swim.h"

namespace swim {
extern thread_local std::string g_last_error;
}
namespace {

struct Out {
  uint8_t *p;
  size_t cap, len = 0;
  bool ok = true;
  void u8(uint8_t v) { if (len < cap) p[len] = v; else ok = false; ++len; }
  void bytes(const void *s, size_t n) { for (size_t i = 0; i < n; ++i) u8(((const uint8_t *)s)[i]); }
  void be16(uint16_t v) { u8(v >> 8); u8(v & 0xFF); }
  void be32(uint32_t v) { be16(v >> 16); be16(v & 0xFFFF); }
  void be64(uint64_t v) { be32((uint32_t)(v >> 32)); be32((uint32_t)v); }
};

void put_int(Out &o, int64_t n) {
  if (n >= -32 && n <= 127) { o.u8((uint8_t)n); return; }
  if (n >= 0) {
    if (n < 0x100) { o.u8(0xCC); o.u8((uint8_t)n); }
    else if (n < 0x10000) { o.u8(0xCD); o.be16((uint16_t)n); }
    else if (n < 0x100000000ll) { o.u8(0xCE); o.be32((uint32_t)n); }
    else { o.u8(0xCF); o.be64((uint64_t)n); }
  } else {
    if (n >= -0x80) { o.u8(0xD0); o.u8((uint8_t)n); }
    else if (n >= -0x8000) { o.u8(0xD1); o.be16((uint16_t)n); }
    else if (n >= -0x80000000ll) { o.u8(0xD2); o.be32((uint32_t)n); }
    else { o.u8(0xD3); o.be64((uint64_t)n); }
  }
}

void put_str(Out &o, const char *s, size_t n) {
  if (n < 32) o.u8(0xA0 | (uint8_t)n);
  else if (n < 0x100) { o.u8(0xD9); o.u8((uint8_t)n); }
  else { o.u8(0xDA); o.be16((uint16_t)n); }
  o.bytes(s, n);
}
void put_key(Out &o, const char *s) { put_str(o, s, strlen(s)); }

// Names are Haskell `String`s carried as msgpack str = UTF-8 text (Types.hs:70,122-145): a byte string that is
// not well-formed UTF-8 (overlong forms, surrogates, > U+10FFFF included) has no `String` and fails to decode.
bool utf8_ok(const char *p, size_t n) {
  const unsigned char *s = (const unsigned char *)p;
  for (size_t i = 0; i < n;) {
    const unsigned c = s[i];
    size_t len; unsigned cp, lo;
    if (c < 0x80) { ++i; continue; }
    else if ((c & 0xE0) == 0xC0) { len = 2; cp = c & 0x1F; lo = 0x80; }
    else if ((c & 0xF0) == 0xE0) { len = 3; cp = c & 0x0F; lo = 0x800; }
    else if ((c & 0xF8) == 0xF0) { len = 4; cp = c & 0x07; lo = 0x10000; }
    else return false;
    if (i + len > n) return false;
    for (size_t k = 1; k < len; ++k) { if ((s[i + k] & 0xC0) != 0x80) return false; cp = (cp << 6) | (s[i + k] & 0x3F); }
    if (cp < lo || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
    i += len;
  }
  return true;
}

// Message body (Types.hs:151-152 `put = putLazyByteString . packAeson`)
bool put_body(Out &o, const swim_wire_message_t &m) {
  const size_t nn = strnlen(m.node, SWIM_NAME_MAX + 1), nf = strnlen(m.dead_from, SWIM_NAME_MAX + 1);
  if (nn > SWIM_NAME_MAX || nf > SWIM_NAME_MAX || !utf8_ok(m.node, nn) || !utf8_ok(m.dead_from, nf)) return false;
  switch (m.kind) {
    case SWIM_MSG_PING:
      o.u8(0x83); put_key(o, "tag"); put_key(o, "Ping");
      put_key(o, "seqNo"); put_int(o, m.seq_no);
      put_key(o, "node"); put_str(o, m.node, nn);
      return true;
    case SWIM_MSG_INDIRECT_PING:
      o.u8(0x85); put_key(o, "tag"); put_key(o, "IndirectPing");
      put_key(o, "seqNo"); put_int(o, m.seq_no);
      put_key(o, "target"); put_int(o, m.target);
      put_key(o, "port"); put_int(o, m.port);
      put_key(o, "node"); put_str(o, m.node, nn);
      return true;
    case SWIM_MSG_ACK:
      if (m.payload_len > SWIM_ACK_PAYLOAD_MAX) return false;
      o.u8(0x83); put_key(o, "tag"); put_key(o, "Ack");
      put_key(o, "seqNo"); put_int(o, m.seq_no);
      put_key(o, "payload");
      if (m.payload_len < 16) o.u8(0x90 | m.payload_len); else { o.u8(0xDC); o.be16(m.payload_len); }
      for (unsigned i = 0; i < m.payload_len; ++i) put_int(o, m.payload[i]);
      return true;
    case SWIM_MSG_SUSPECT:
      o.u8(0x83); put_key(o, "tag"); put_key(o, "Suspect");
      put_key(o, "incarnation"); put_int(o, m.incarnation);
      put_key(o, "node"); put_str(o, m.node, nn);
      return true;
    case SWIM_MSG_ALIVE:
      o.u8(0x85); put_key(o, "tag"); put_key(o, "Alive");
      put_key(o, "incarnation"); put_int(o, m.incarnation);
      put_key(o, "node"); put_str(o, m.node, nn);
      put_key(o, "addr"); put_int(o, m.target);
      put_key(o, "port"); put_int(o, m.port);
      return true;
    case SWIM_MSG_DEAD:
      o.u8(0x84); put_key(o, "tag"); put_key(o, "Dead");
      put_key(o, "incarnation"); put_int(o, m.incarnation);
      put_key(o, "node"); put_str(o, m.node, nn);
      put_key(o, "deadFrom"); put_str(o, m.dead_from, nf);
      return true;
  }
  return false;
}

// ---------------------------------------------------------------- msgpack reader
struct In {
  const uint8_t *p;
  size_t len, pos = 0;
  bool ok = true;
  size_t left() const { return len - pos; }
  uint8_t u8() { if (pos < len) return p[pos++]; ok = false; return 0; }
  uint64_t be(int n) { uint64_t v = 0; for (int i = 0; i < n; ++i) v = (v << 8) | u8(); return v; }
};

struct Val { // the JSON values a Message body can hold
  enum { INT, STR, ARR, OTHER } t = OTHER;
  int64_t i = 0;
  bool big = false; // uint64 above INT64_MAX
  std::string s;
  std::vector<int64_t> a;
  bool arr_ints = true;
};

bool read_val(In &in, Val &v, int depth = 0);

bool skip_n(In &in, size_t n, int depth) {
  for (size_t x = 0; x < n && in.ok; ++x) { Val t; if (!read_val(in, t, depth + 1)) return false; }
  return in.ok;
}

bool read_val(In &in, Val &v, int depth) {
  if (depth > 8 || in.left() == 0) return false;
  const uint8_t b = in.u8();
  auto str = [&](size_t n) { if (in.left() < n) return false; v.t = Val::STR; v.s.assign((const char *)in.p + in.pos, n); in.pos += n; return true; };
  auto arr = [&](size_t n) {
    v.t = Val::ARR;
    for (size_t x = 0; x < n; ++x) {
      Val e;
      if (!read_val(in, e, depth + 1)) return false;
      if (e.t != Val::INT || e.big) v.arr_ints = false; else v.a.push_back(e.i);
    }
    return true;
  };
  if (b <= 0x7F) { v.t = Val::INT; v.i = b; return true; }
  if (b >= 0xE0) { v.t = Val::INT; v.i = (int8_t)b; return true; }
  if ((b & 0xE0) == 0xA0) return str(b & 0x1F);
  if ((b & 0xF0) == 0x90) return arr(b & 0x0F);
  if ((b & 0xF0) == 0x80) { v.t = Val::OTHER; return skip_n(in, 2 * (size_t)(b & 0x0F), depth); }
  switch (b) {
    case 0xC0: case 0xC2: case 0xC3: v.t = Val::OTHER; return true;
    case 0xCC: v.t = Val::INT; v.i = (int64_t)in.be(1); return in.ok;
    case 0xCD: v.t = Val::INT; v.i = (int64_t)in.be(2); return in.ok;
    case 0xCE: v.t = Val::INT; v.i = (int64_t)in.be(4); return in.ok;
    case 0xCF: { uint64_t u = in.be(8); v.t = Val::INT; v.big = u > 0x7FFFFFFFFFFFFFFFull; v.i = (int64_t)u; return in.ok; }
    case 0xD0: v.t = Val::INT; v.i = (int8_t)in.be(1); return in.ok;
    case 0xD1: v.t = Val::INT; v.i = (int16_t)in.be(2); return in.ok;
    case 0xD2: v.t = Val::INT; v.i = (int32_t)in.be(4); return in.ok;
    case 0xD3: v.t = Val::INT; v.i = (int64_t)in.be(8); return in.ok;
    case 0xD9: { size_t n = (size_t)in.be(1); return in.ok && str(n); }
    case 0xDA: { size_t n = (size_t)in.be(2); return in.ok && str(n); }
    case 0xDB: { size_t n = (size_t)in.be(4); return in.ok && str(n); }
    case 0xC4: { size_t n = (size_t)in.be(1); v.t = Val::OTHER; if (in.left() < n) return false; in.pos += n; return in.ok; }
    case 0xC5: { size_t n = (size_t)in.be(2); v.t = Val::OTHER; if (in.left() < n) return false; in.pos += n; return in.ok; }
    case 0xC6: { size_t n = (size_t)in.be(4); v.t = Val::OTHER; if (in.left() < n) return false; in.pos += n; return in.ok; }
    case 0xCA: in.be(4); v.t = Val::OTHER; return in.ok;
    case 0xCB: in.be(8); v.t = Val::OTHER; return in.ok;
    case 0xDC: { size_t n = (size_t)in.be(2); return in.ok && arr(n); }
    case 0xDD: { size_t n = (size_t)in.be(4); return in.ok && n <= in.left() && arr(n); }
    case 0xDE: { size_t n = (size_t)in.be(2); v.t = Val::OTHER; return in.ok && skip_n(in, 2 * n, depth); }
    case 0xDF: { size_t n = (size_t)in.be(4); v.t = Val::OTHER; return in.ok && n <= in.left() && skip_n(in, 2 * n, depth); }
  }
  return false; // ext types / reserved: not produced by aeson values
}

// `unpackAeson` + generic FromJSON (Types.hs:153-155): a map with a "tag" and the record fields
bool get_body(const uint8_t *p, size_t len, swim_wire_message_t &m) {
  In in{p, len};
  if (in.left() == 0) return false;
  const uint8_t b = in.u8();
  size_t n;
  if ((b & 0xF0) == 0x80) n = b & 0x0F;
  else if (b == 0xDE) n = (size_t)in.be(2);
  else if (b == 0xDF) n = (size_t)in.be(4);
  else return false;
  if (!in.ok || n > in.left()) return false;
  std::string tag;
  bool h_tag = false, h_seq = false, h_node = false, h_target = false, h_port = false, h_payload = false,
       h_inc = false, h_addr = false, h_from = false;
  int64_t seq = 0, target = 0, port = 0, inc = 0, addr = 0;
  std::string node, from;
  std::vector<int64_t> payload;
  for (size_t x = 0; x < n; ++x) {
    Val k, v;
    if (!read_val(in, k) || k.t != Val::STR || !read_val(in, v)) return false;
    const std::string &key = k.s;
    if (key == "tag") { if (v.t != Val::STR) return false; tag = v.s; h_tag = true; }
    else if (key == "seqNo") { if (v.t != Val::INT || v.big) return false; seq = v.i; h_seq = true; }
    else if (key == "node") { if (v.t != Val::STR) return false; node = v.s; h_node = true; }
    else if (key == "target") { if (v.t != Val::INT || v.big) return false; target = v.i; h_target = true; }
    else if (key == "port") { if (v.t != Val::INT || v.big) return false; port = v.i; h_port = true; }
    else if (key == "payload") { if (v.t != Val::ARR || !v.arr_ints) return false; payload = v.a; h_payload = true; }
    else if (key == "incarnation") { if (v.t != Val::INT || v.big) return false; inc = v.i; h_inc = true; }
    else if (key == "addr") { if (v.t != Val::INT || v.big) return false; addr = v.i; h_addr = true; }
    else if (key == "deadFrom") { if (v.t != Val::STR) return false; from = v.s; h_from = true; }
    // unknown keys are ignored, as aeson's generic parser does
  }
  if (!h_tag) return false;
  memset(&m, 0, sizeof m);
  auto u32ok = [](int64_t v) { return v >= 0 && v <= 0xFFFFFFFFll; }; // Word32 fields
  auto u16ok = [](int64_t v) { return v >= 0 && v <= 0xFFFF; };       // Word16 fields
  auto name = [](char *dst, const std::string &s) {
    if (s.size() > SWIM_NAME_MAX || memchr(s.data(), 0, s.size()) || !utf8_ok(s.data(), s.size())) return false;
    memcpy(dst, s.data(), s.size());
    return true;
  };
  if (tag == "Ping") {
    if (!h_seq || !h_node || !u32ok(seq) || !name(m.node, node)) return false;
    m.kind = SWIM_MSG_PING; m.seq_no = (uint32_t)seq;
  } else if (tag == "IndirectPing") {
    if (!h_seq || !h_target || !h_port || !h_node || !u32ok(seq) || !u32ok(target) || !u16ok(port) || !name(m.node, node)) return false;
    m.kind = SWIM_MSG_INDIRECT_PING; m.seq_no = (uint32_t)seq; m.target = (uint32_t)target; m.port = (uint16_t)port;
  } else if (tag == "Ack") {
    if (!h_seq || !h_payload || !u32ok(seq) || payload.size() > SWIM_ACK_PAYLOAD_MAX) return false;
    m.kind = SWIM_MSG_ACK; m.seq_no = (uint32_t)seq; m.payload_len = (uint8_t)payload.size();
    for (size_t i = 0; i < payload.size(); ++i) { if (payload[i] < 0 || payload[i] > 255) return false; m.payload[i] = (uint8_t)payload[i]; }
  } else if (tag == "Suspect") {
    if (!h_inc || !h_node || !name(m.node, node)) return false;
    m.kind = SWIM_MSG_SUSPECT; m.incarnation = inc;
  } else if (tag == "Alive") {
    if (!h_inc || !h_node || !h_addr || !h_port || !u32ok(addr) || !u16ok(port) || !name(m.node, node)) return false;
    m.kind = SWIM_MSG_ALIVE; m.incarnation = inc; m.target = (uint32_t)addr; m.port = (uint16_t)port;
  } else if (tag == "Dead") {
    if (!h_inc || !h_node || !h_from || !name(m.node, node) || !name(m.dead_from, from)) return false;
    m.kind = SWIM_MSG_DEAD; m.incarnation = inc;
  } else {
    return false;
  }
  return true;
}

int decode_fail(const std::string &why) {
  swim::g_last_error = why;
  return SWIM_EDECODE;
}

} // namespace

// `put (Envelope ...)` (Types.hs:96-103)
extern "C" int swim_envelope_encode(const swim_wire_message_t *msgs, size_t n, uint8_t *buf, size_t cap,
                                    size_t *len) {
  if (!msgs || !buf || !len || n == 0) { swim::g_last_error = "swim_envelope_encode: an Envelope is a NonEmpty list"; return SWIM_EINVAL; }
  if (n > 255) { swim::g_last_error = "swim_envelope_encode: more than 255 messages (count is one byte, Types.hs:100)"; return SWIM_ECAP; }
  Out o{buf, cap};
  if (n == 1) {
    // Envelope (msg :| []) = putWord8 (msgIndex msg) >> put msg
    if (msgs[0].kind > SWIM_MSG_DEAD) return SWIM_EINVAL;
    o.u8(msgs[0].kind);
    if (!put_body(o, msgs[0])) return SWIM_EINVAL;
  } else {
    o.u8(SWIM_MSG_COMPOUND);
    o.u8((uint8_t)n);
    std::vector<std::vector<uint8_t>> bodies(n);
    for (size_t x = 0; x < n; ++x) {
      bodies[x].resize(1024 + 2 * SWIM_NAME_MAX);
      Out b{bodies[x].data(), bodies[x].size()};
      if (!put_body(b, msgs[x]) || !b.ok) return SWIM_EINVAL;
      bodies[x].resize(b.len);
      o.be16((uint16_t)b.len); // putWord16be . fromIntegral . BS.length
    }
    for (size_t x = 0; x < n; ++x) o.bytes(bodies[x].data(), bodies[x].size()); // no per-body type byte
  }
  *len = o.len;
  if (!o.ok) { swim::g_last_error = "swim_envelope_encode: buffer too small"; return SWIM_ECAP; }
  return SWIM_OK;
}

// `get :: Get Envelope` (Types.hs:105-119)
extern "C" int swim_envelope_decode(const uint8_t *buf, size_t len, swim_wire_message_t *msgs, size_t cap,
                                    size_t *n_out) {
  if (!buf || !msgs || !n_out) return SWIM_EINVAL;
  *n_out = 0;
  if (len < 1) return decode_fail("too few bytes");
  const uint8_t typ = buf[0];
  if (typ > SWIM_MSG_COMPOUND) { // reference: `toEnum` out of range is a crash (Types.hs:110), not a parse error
    char tmp[64];
    snprintf(tmp, sizeof tmp, "invalid message type %u", typ);
    return decode_fail(tmp);
  }
  if (typ != SWIM_MSG_COMPOUND) {
    // the type byte is ignored; the constructor comes from the body (Types.hs:93-94,119)
    if (cap < 1) return SWIM_ECAP;
    if (!get_body(buf + 1, len - 1, msgs[0])) return decode_fail("Could not parse message body");
    *n_out = 1;
    return SWIM_OK;
  }
  if (len < 2) return decode_fail("too few bytes");
  const size_t num = buf[1];
  size_t pos = 2;
  if (len - pos < num * 2) return decode_fail("compound message is truncated"); // Types.hs:114-115
  if (num == 0) return decode_fail("compound mesage with zero messages");        // Types.hs:118 [sic]
  if (num > cap) return SWIM_ECAP;
  std::vector<size_t> lens(num);
  for (size_t x = 0; x < num; ++x) { lens[x] = ((size_t)buf[pos] << 8) | buf[pos + 1]; pos += 2; }
  for (size_t x = 0; x < num; ++x) {
    if (len - pos < lens[x]) return decode_fail("too few bytes"); // `isolate` runs out of input
    if (!get_body(buf + pos, lens[x], msgs[x])) return decode_fail("Could not parse message body");
    pos += lens[x];
  }
  *n_out = num;
  return SWIM_OK;
}
