// hdf5_io.cpp -- see hdf5_io.hpp.  Structures follow the HDF5 File Format Specification, "version 0" superblock
// family (what libhdf5 writes with default property lists); byte layouts cross-checked against the libhdf5-written
// fixtures in the reference tree (caffe-public/src/caffe/test/test_data/*.h5).
#include "hdf5_io.hpp"

#include <stdio.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <set>

namespace cosb {
namespace {

constexpr uint64_t kUndef = ~0ull;
constexpr int kLeafK = 4;       // symbol table node holds up to 2*kLeafK entries
constexpr int kInternalK = 16;  // group B-tree node holds up to 2*kInternalK children
constexpr uint64_t kBtreeBytes = 24 + (2 * kInternalK + 1) * 8 + 2 * kInternalK * 8;  // 544
constexpr uint64_t kSnodBytes = 8 + 2 * kLeafK * 40;                                   // 328
constexpr uint64_t kHeapFree = 64;  // free block kept at the end of every local heap (libhdf5 leaves one too)
const unsigned char kSignature[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};

uint64_t align8(uint64_t v) { return (v + 7) & ~7ull; }

// ------------------------------------------------------------------ writer
struct Bytes {
  std::string b;
  void u8(unsigned v) { b.push_back(static_cast<char>(v)); }
  void u16(unsigned v) { u8(v & 0xff); u8((v >> 8) & 0xff); }
  void u32(uint64_t v) { for (int i = 0; i < 4; ++i) u8((v >> (8 * i)) & 0xff); }
  void u64(uint64_t v) { for (int i = 0; i < 8; ++i) u8((v >> (8 * i)) & 0xff); }
  void raw(const void* p, size_t n) { b.append(static_cast<const char*>(p), n); }
  void zeros(size_t n) { b.append(n, '\0'); }
  void pad8() { b.append(align8(b.size()) - b.size(), '\0'); }
};

struct Plan {  // addresses of one node's on-disk pieces
  uint64_t ohdr = 0, btree = 0, heap = 0, heap_data = 0, heap_data_size = 0, raw = 0, raw_bytes = 0;
  std::vector<uint64_t> snods;
  std::vector<const H5Node*> sorted;           // children by name
  std::vector<uint64_t> name_off;              // heap offsets of the sorted children's names
  std::vector<std::pair<size_t, size_t>> cut;  // [first, last) child index per symbol table node
};

struct Writer {
  std::map<const H5Node*, Plan> plan;
  uint64_t top = 96;  // superblock (56) + root symbol table entry (40)
  std::string err;

  uint64_t alloc(uint64_t n) {
    top = align8(top);
    uint64_t a = top;
    top += n;
    return a;
  }
  static uint64_t elem_size(const H5Node& n) { return n.kind == H5Node::kString ? n.str.size() + 1 : 4; }
  static uint64_t dataset_header_bytes(const H5Node& n) {
    const uint64_t rank = n.kind == H5Node::kString ? 0 : n.shape.size();
    const uint64_t space = 8 + (rank ? 16 * rank : 0);
    const uint64_t type = n.kind == H5Node::kFloat32 ? 24 : (n.kind == H5Node::kInt32 ? 16 : 8);
    return (8 + space) + (8 + type) + (8 + 8) + (8 + 24) + (8 + 8);
  }

  bool layout(const H5Node& n) {
    Plan& p = plan[&n];
    if (n.kind == H5Node::kGroup) {
      if (n.children.size() > static_cast<size_t>(2 * kInternalK * 2 * kLeafK)) {
        err = "group '" + n.name + "' has more than 256 links (would need a two-level B-tree)";
        return false;
      }
      for (const auto& c : n.children) p.sorted.push_back(c.get());
      std::sort(p.sorted.begin(), p.sorted.end(),
                [](const H5Node* a, const H5Node* b) { return strcmp(a->name.c_str(), b->name.c_str()) < 0; });
      for (size_t i = 1; i < p.sorted.size(); ++i) {
        if (p.sorted[i]->name == p.sorted[i - 1]->name) {
          err = "duplicate link name '" + p.sorted[i]->name + "' in group '" + n.name + "'";
          return false;
        }
      }
      p.ohdr = alloc(16 + 8 + 16);
      p.btree = alloc(kBtreeBytes);
      uint64_t used = 8;  // offset 0: the empty string
      for (const H5Node* c : p.sorted) {
        if (c->name.empty() || c->name.find('/') != std::string::npos) {
          err = "invalid link name '" + c->name + "'";
          return false;
        }
        p.name_off.push_back(used);
        used += align8(c->name.size() + 1);
      }
      p.heap_data_size = used + kHeapFree;
      p.heap = alloc(32);
      p.heap_data = alloc(p.heap_data_size);
      const size_t nchild = p.sorted.size();
      const size_t nsnod = nchild == 0 ? 0 : (nchild + 2 * kLeafK - 1) / (2 * kLeafK);
      for (size_t s = 0, first = 0; s < nsnod; ++s) {  // spread evenly: every node holds >= kLeafK entries
        const size_t cnt = nchild / nsnod + (s < nchild % nsnod ? 1 : 0);
        p.cut.emplace_back(first, first + cnt);
        first += cnt;
        p.snods.push_back(alloc(kSnodBytes));
      }
      for (const H5Node* c : p.sorted)
        if (!layout(*c)) return false;
      return true;
    }
    uint64_t count = 1;
    for (int64_t d : n.shape) {
      if (d < 0) {
        err = "negative dimension in dataset '" + n.name + "'";
        return false;
      }
      count *= static_cast<uint64_t>(d);
    }
    if (n.kind != H5Node::kString && count != n.count) {
      err = "dataset '" + n.name + "': shape does not match its element count";
      return false;
    }
    p.ohdr = alloc(16 + dataset_header_bytes(n));
    p.raw_bytes = n.kind == H5Node::kString ? n.str.size() + 1 : count * 4;
    p.raw = p.raw_bytes ? alloc(p.raw_bytes) : kUndef;
    return true;
  }

  static void message(Bytes* o, unsigned type, unsigned flags, const Bytes& body) {
    const uint64_t size = align8(body.b.size());
    o->u16(type);
    o->u16(static_cast<unsigned>(size));
    o->u8(flags);
    o->zeros(3);
    o->raw(body.b.data(), body.b.size());
    o->zeros(size - body.b.size());
  }

  static void symbol_entry(Bytes* o, uint64_t name_off, const H5Node& c, const Plan& cp) {
    o->u64(name_off);
    o->u64(cp.ohdr);
    if (c.kind == H5Node::kGroup) {  // cache type 1: B-tree and heap addresses in the scratch pad
      o->u32(1);
      o->u32(0);
      o->u64(cp.btree);
      o->u64(cp.heap);
    } else {
      o->u32(0);
      o->u32(0);
      o->zeros(16);
    }
  }

  bool emit(FILE* f, uint64_t addr, const std::string& bytes) {
    return fseek(f, static_cast<long>(addr), SEEK_SET) == 0 && fwrite(bytes.data(), 1, bytes.size(), f) == bytes.size();
  }

  bool write_node(FILE* f, const H5Node& n, uint32_t now) {
    const Plan& p = plan[&n];
    if (n.kind == H5Node::kGroup) {
      Bytes oh;  // v1 object header: version, reserved, #messages, ref count, header size, pad to 8
      oh.u8(1); oh.u8(0); oh.u16(1); oh.u32(1); oh.u32(8 + 16); oh.u32(0);
      Bytes stab;
      stab.u64(p.btree);
      stab.u64(p.heap);
      message(&oh, 0x0011, 0, stab);
      Bytes bt;  // group B-tree node, level 0
      bt.raw("TREE", 4); bt.u8(0); bt.u8(0); bt.u16(static_cast<unsigned>(p.snods.size()));
      bt.u64(kUndef); bt.u64(kUndef);
      bt.u64(0);  // key 0: the empty string
      for (size_t s = 0; s < p.snods.size(); ++s) {
        bt.u64(p.snods[s]);
        bt.u64(p.name_off[p.cut[s].second - 1]);  // key s+1: largest name of child s
      }
      bt.zeros(kBtreeBytes - bt.b.size());
      Bytes hp;  // local heap header + data segment
      hp.raw("HEAP", 4); hp.u8(0); hp.zeros(3);
      hp.u64(p.heap_data_size);
      hp.u64(p.heap_data_size - kHeapFree);  // head of the free list
      hp.u64(p.heap_data);
      Bytes hd;
      hd.zeros(8);
      for (const H5Node* c : p.sorted) {
        hd.raw(c->name.c_str(), c->name.size() + 1);
        hd.pad8();
      }
      hd.u64(1);          // free block: next = H5HL_FREE_NULL
      hd.u64(kHeapFree);  //             size
      hd.zeros(p.heap_data_size - hd.b.size());
      if (!emit(f, p.ohdr, oh.b) || !emit(f, p.btree, bt.b) || !emit(f, p.heap, hp.b) || !emit(f, p.heap_data, hd.b))
        return false;
      for (size_t s = 0; s < p.snods.size(); ++s) {
        Bytes sn;
        sn.raw("SNOD", 4); sn.u8(1); sn.u8(0); sn.u16(static_cast<unsigned>(p.cut[s].second - p.cut[s].first));
        for (size_t i = p.cut[s].first; i < p.cut[s].second; ++i)
          symbol_entry(&sn, p.name_off[i], *p.sorted[i], plan[p.sorted[i]]);
        sn.zeros(kSnodBytes - sn.b.size());
        if (!emit(f, p.snods[s], sn.b)) return false;
      }
      for (const H5Node* c : p.sorted)
        if (!write_node(f, *c, now)) return false;
      return true;
    }
    Bytes oh;
    oh.u8(1); oh.u8(0); oh.u16(5); oh.u32(1); oh.u32(dataset_header_bytes(n)); oh.u32(0);
    Bytes space;
    if (n.kind == H5Node::kString) {
      space.u8(1); space.u8(0); space.u8(0); space.zeros(5);  // scalar
    } else {
      space.u8(1); space.u8(static_cast<unsigned>(n.shape.size())); space.u8(1); space.zeros(5);  // max dims present
      for (int64_t d : n.shape) space.u64(static_cast<uint64_t>(d));
      for (int64_t d : n.shape) space.u64(static_cast<uint64_t>(d));
    }
    message(&oh, 0x0001, 0, space);
    Bytes type;
    if (n.kind == H5Node::kFloat32) {  // class 1 v1, little endian, IEEE: sign 31, exponent 23/8 bias 127, mantissa 0/23
      type.u8(0x11); type.u8(0x20); type.u8(0x1f); type.u8(0x00); type.u32(4);
      type.u16(0); type.u16(32); type.u8(23); type.u8(8); type.u8(0); type.u8(23); type.u32(127);
    } else if (n.kind == H5Node::kInt32) {  // class 0 v1, little endian, signed two's complement
      type.u8(0x10); type.u8(0x08); type.u8(0x00); type.u8(0x00); type.u32(4);
      type.u16(0); type.u16(32);
    } else {  // class 3 v1: null-terminated ASCII, size = strlen + 1
      type.u8(0x13); type.u8(0x00); type.u8(0x00); type.u8(0x00); type.u32(n.str.size() + 1);
    }
    message(&oh, 0x0003, 1, type);
    Bytes fill;
    fill.u8(2); fill.u8(2); fill.u8(2); fill.u8(1); fill.u32(0);  // v2: late allocation, write if set, defined, size 0
    message(&oh, 0x0005, 1, fill);
    Bytes lay;
    lay.u8(3); lay.u8(1); lay.u64(p.raw); lay.u64(p.raw_bytes);  // v3, contiguous
    message(&oh, 0x0008, 1, lay);
    Bytes mt;
    mt.u8(1); mt.zeros(3); mt.u32(now);
    message(&oh, 0x0012, 0, mt);
    if (!emit(f, p.ohdr, oh.b)) return false;
    if (p.raw_bytes) {
      const void* src = n.kind == H5Node::kString ? static_cast<const void*>(n.str.c_str())
                        : n.data                  ? n.data
                        : n.kind == H5Node::kFloat32 ? static_cast<const void*>(n.f32.data())
                                                     : static_cast<const void*>(n.i32.data());
      if (fseek(f, static_cast<long>(p.raw), SEEK_SET) != 0 || fwrite(src, 1, p.raw_bytes, f) != p.raw_bytes)
        return false;
    }
    return true;
  }
};

// ------------------------------------------------------------------ reader
struct Reader {
  std::vector<unsigned char> b;
  std::string err;
  std::set<uint64_t> visiting;

  bool ok(uint64_t off, uint64_t n) const { return off <= b.size() && n <= b.size() - off; }
  uint64_t le(uint64_t off, int n) const {
    uint64_t v = 0;
    for (int i = n - 1; i >= 0; --i) v = (v << 8) | b[off + i];
    return v;
  }
  bool fail(const std::string& m) {
    if (err.empty()) err = m;
    return false;
  }

  struct Msg { unsigned type, flags; uint64_t off, size; };

  bool messages(uint64_t addr, std::vector<Msg>* out) {
    if (!ok(addr, 16) || b[addr] != 1) return fail("unsupported object header (only version 1 is handled)");
    const unsigned nmsg = static_cast<unsigned>(le(addr + 2, 2));
    std::vector<std::pair<uint64_t, uint64_t>> blocks{{addr + 16, le(addr + 8, 4)}};
    for (size_t bi = 0; bi < blocks.size() && out->size() < nmsg; ++bi) {
      uint64_t pos = blocks[bi].first;
      const uint64_t end = pos + blocks[bi].second;
      if (!ok(pos, blocks[bi].second)) return fail("object header block outside the file");
      while (pos + 8 <= end && out->size() < nmsg) {
        Msg m{static_cast<unsigned>(le(pos, 2)), b[pos + 4], pos + 8, le(pos + 2, 2)};
        if (!ok(m.off, m.size) || m.off + m.size > end) return fail("object header message outside its block");
        if (m.flags & 2) return fail("shared object header messages are not supported");
        if (m.type == 0x0010) {
          if (m.size < 16) return fail("short continuation message");
          blocks.emplace_back(le(m.off, 8), le(m.off + 8, 8));
          if (blocks.size() > 64) return fail("too many object header continuations");
        }
        out->push_back(m);
        pos = m.off + m.size;
      }
    }
    return true;
  }

  bool heap_name(uint64_t heap, uint64_t off, std::string* name) {
    if (!ok(heap, 32) || memcmp(&b[heap], "HEAP", 4) != 0) return fail("bad local heap");
    const uint64_t dsize = le(heap + 8, 8), daddr = le(heap + 24, 8);
    if (!ok(daddr, dsize) || off >= dsize) return fail("local heap data segment outside the file");
    const unsigned char* s = &b[daddr + off];
    const void* z = memchr(s, 0, dsize - off);
    if (!z) return fail("unterminated link name");
    name->assign(reinterpret_cast<const char*>(s), static_cast<const unsigned char*>(z) - s);
    return true;
  }

  bool btree(uint64_t node, uint64_t heap, H5Node* group, int depth) {
    if (depth > 8 || !ok(node, 24) || memcmp(&b[node], "TREE", 4) != 0 || b[node + 4] != 0)
      return fail("bad group B-tree node");
    const unsigned level = b[node + 5], used = static_cast<unsigned>(le(node + 6, 2));
    if (!ok(node + 24, static_cast<uint64_t>(used) * 16 + 8)) return fail("group B-tree node outside the file");
    for (unsigned i = 0; i < used; ++i) {
      const uint64_t child = le(node + 24 + 8 + static_cast<uint64_t>(i) * 16, 8);
      if (level > 0) {
        if (!btree(child, heap, group, depth + 1)) return false;
        continue;
      }
      if (!ok(child, 8) || memcmp(&b[child], "SNOD", 4) != 0) return fail("bad symbol table node");
      const unsigned nsym = static_cast<unsigned>(le(child + 6, 2));
      if (!ok(child + 8, static_cast<uint64_t>(nsym) * 40)) return fail("symbol table node outside the file");
      for (unsigned s = 0; s < nsym; ++s) {
        const uint64_t e = child + 8 + static_cast<uint64_t>(s) * 40;
        std::unique_ptr<H5Node> c(new H5Node());
        if (!heap_name(heap, le(e, 8), &c->name)) return false;
        if (!object(le(e + 8, 8), c.get(), depth + 1)) return false;
        group->children.push_back(std::move(c));
        if (group->children.size() > 1000000) return fail("too many links");
      }
    }
    return true;
  }

  bool object(uint64_t addr, H5Node* n, int depth) {
    if (depth > 32 || !visiting.insert(addr).second) return fail("cyclic or too deep group structure");
    std::vector<Msg> msgs;
    if (!messages(addr, &msgs)) return false;
    const Msg *stab = nullptr, *space = nullptr, *type = nullptr, *lay = nullptr;
    for (const Msg& m : msgs) {
      if (m.type == 0x0011) stab = &m;
      else if (m.type == 0x0001) space = &m;
      else if (m.type == 0x0003) type = &m;
      else if (m.type == 0x0008) lay = &m;
      else if (m.type == 0x000b) return fail("dataset '" + n->name + "' uses a filter pipeline (compression): not supported");
      else if (m.type == 0x0002 || m.type == 0x0006) return fail("new-style (link message) groups are not supported");
    }
    bool r;
    if (stab) {
      n->kind = H5Node::kGroup;
      r = stab->size >= 16 && btree(le(stab->off, 8), le(stab->off + 8, 8), n, depth);
      if (!r) fail("bad symbol table message");
    } else if (space && type && lay) {
      r = dataset(*space, *type, *lay, n);
    } else {
      r = fail("object '" + n->name + "' is neither an old-style group nor a simple dataset");
    }
    visiting.erase(addr);
    return r;
  }

  bool dataset(const Msg& space, const Msg& type, const Msg& lay, H5Node* n) {
    // dataspace: v1 = {version, rank, flags, reserved x5, dims...}, v2 = {version, rank, flags, type, dims...}
    if (space.size < 4) return fail("short dataspace message");
    const unsigned sv = b[space.off], rank = b[space.off + 1];
    const uint64_t dims_at = space.off + (sv == 1 ? 8 : 4);
    if ((sv != 1 && sv != 2) || rank > 32 || dims_at + 8ull * rank > space.off + space.size)
      return fail("unsupported dataspace message");
    uint64_t count = 1;
    for (unsigned i = 0; i < rank; ++i) {
      const uint64_t d = le(dims_at + 8ull * i, 8);
      if (d > (1ull << 40) || (d && count > (1ull << 40) / d)) return fail("dataset '" + n->name + "' is too large");
      n->shape.push_back(static_cast<int64_t>(d));
      count *= d;
    }
    if (sv == 2 && b[space.off + 3] == 2) count = 0;  // null dataspace
    // datatype
    if (type.size < 8) return fail("short datatype message");
    const unsigned cls = b[type.off] & 0x0f, bits0 = b[type.off + 1];
    const uint64_t esize = le(type.off + 4, 4);
    if (cls == 1 && esize == 4 && !(bits0 & 1)) n->kind = H5Node::kFloat32;
    else if (cls == 0 && esize == 4 && !(bits0 & 1)) n->kind = H5Node::kInt32;
    else if (cls == 3) n->kind = H5Node::kString;
    else return fail("dataset '" + n->name + "': only little-endian float32 / int32 and fixed strings are supported");
    // layout
    if (lay.size < 2 || b[lay.off] != 3) return fail("dataset '" + n->name + "': unsupported data layout version");
    const unsigned lclass = b[lay.off + 1];
    uint64_t addr = kUndef, size = 0;
    if (lclass == 1) {
      if (lay.size < 18) return fail("short layout message");
      addr = le(lay.off + 2, 8);
      size = le(lay.off + 10, 8);
    } else if (lclass == 0) {
      if (lay.size < 4) return fail("short layout message");
      size = le(lay.off + 2, 2);
      addr = lay.off + 4;
      if (addr + size > lay.off + lay.size) return fail("compact dataset outside its message");
    } else {
      return fail("dataset '" + n->name + "' is chunked: only contiguous / compact layouts are supported");
    }
    const uint64_t want = n->kind == H5Node::kString ? count * esize : count * 4;
    n->count = count;
    if (addr == kUndef || want == 0) {  // never written: zeros / empty
      if (count > (1ull << 28)) return fail("dataset '" + n->name + "' has no data but claims " + std::to_string(count) + " elements");
      if (n->kind == H5Node::kFloat32) n->f32.assign(count, 0.f);
      else if (n->kind == H5Node::kInt32) n->i32.assign(count, 0);
      return true;
    }
    if (size < want || !ok(addr, want)) return fail("dataset '" + n->name + "': raw data outside the file");
    if (n->kind == H5Node::kFloat32) {
      n->f32.resize(count);
      memcpy(n->f32.data(), &b[addr], want);
    } else if (n->kind == H5Node::kInt32) {
      n->i32.resize(count);
      memcpy(n->i32.data(), &b[addr], want);
    } else {
      n->str.assign(reinterpret_cast<const char*>(&b[addr]), strnlen(reinterpret_cast<const char*>(&b[addr]), esize));
    }
    return true;
  }
};

}  // namespace

H5Node* H5Node::add_group(const std::string& n) {
  children.emplace_back(new H5Node());
  children.back()->name = n;
  return children.back().get();
}
H5Node* H5Node::add_float(const std::string& n, const std::vector<int64_t>& shp, const float* d, uint64_t cnt) {
  H5Node* c = add_group(n);
  c->kind = kFloat32;
  c->shape = shp;
  c->data = d;
  c->count = cnt;
  return c;
}
H5Node* H5Node::add_int(const std::string& n, int32_t v) {
  H5Node* c = add_group(n);
  c->kind = kInt32;
  c->shape = {1};
  c->i32 = {v};
  c->count = 1;
  return c;
}
H5Node* H5Node::add_string(const std::string& n, const std::string& s) {
  H5Node* c = add_group(n);
  c->kind = kString;
  c->str = s;
  c->count = 1;
  return c;
}
const H5Node* H5Node::find(const std::string& n) const {
  for (const auto& c : children)
    if (c->name == n) return c.get();
  return nullptr;
}

bool h5_write(const std::string& path, const H5Node& root, std::string* err) {
  Writer w;
  if (root.kind != H5Node::kGroup) {
    *err = "the root must be a group";
    return false;
  }
  if (!w.layout(root)) {
    *err = "'" + path + "': " + w.err;
    return false;
  }
  const uint64_t eof = align8(w.top);
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) {
    *err = "cannot open '" + path + "' for writing";
    return false;
  }
  const Plan& rp = w.plan[&root];
  Bytes sb;  // superblock v0
  sb.raw(kSignature, 8);
  sb.u8(0); sb.u8(0); sb.u8(0); sb.u8(0); sb.u8(0);  // superblock, free-space, root entry, reserved, shared header versions
  sb.u8(8); sb.u8(8); sb.u8(0);                      // size of offsets, size of lengths, reserved
  sb.u16(kLeafK); sb.u16(kInternalK); sb.u32(0);     // group leaf K, group internal K, consistency flags
  sb.u64(0); sb.u64(kUndef); sb.u64(eof); sb.u64(kUndef);  // base, free-space info, end of file, driver info
  sb.u64(0); sb.u64(rp.ohdr); sb.u32(1); sb.u32(0); sb.u64(rp.btree); sb.u64(rp.heap);  // root symbol table entry
  bool ok = w.emit(f, 0, sb.b) && w.write_node(f, root, static_cast<uint32_t>(time(nullptr)));
  ok = ok && fflush(f) == 0 && ftruncate(fileno(f), static_cast<off_t>(eof)) == 0;
  ok = (fclose(f) == 0) && ok;
  if (!ok) *err = "short write to '" + path + "'";
  return ok;
}

bool h5_is_hdf5(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  unsigned char sig[8];
  const bool is = fread(sig, 1, 8, f) == 8 && memcmp(sig, kSignature, 8) == 0;
  fclose(f);
  return is;
}

bool h5_read(const std::string& path, H5Node* root, std::string* err) {
  Reader r;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    *err = "cannot open '" + path + "'";
    return false;
  }
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (n < 96) {
    fclose(f);
    *err = "'" + path + "' is not an HDF5 file";
    return false;
  }
  r.b.resize(static_cast<size_t>(n));
  const bool got = fread(r.b.data(), 1, r.b.size(), f) == r.b.size();
  fclose(f);
  if (!got || memcmp(r.b.data(), kSignature, 8) != 0) {
    *err = "'" + path + "' is not an HDF5 file";
    return false;
  }
  if (r.b[8] > 1 || r.b[13] != 8 || r.b[14] != 8) {
    *err = "'" + path + "': only superblock versions 0/1 with 8-byte offsets are supported";
    return false;
  }
  const uint64_t ste = r.b[8] == 0 ? 56 : 60;  // v1 inserts 4 bytes (indexed storage K + reserved)
  if (r.le(24, 8) != 0) {
    *err = "'" + path + "': non-zero base address is not supported";
    return false;
  }
  root->kind = H5Node::kGroup;
  root->name = "/";
  if (!r.object(r.le(ste + 8, 8), root, 0)) {
    *err = "'" + path + "': " + r.err;
    return false;
  }
  return true;
}

}  // namespace cosb
