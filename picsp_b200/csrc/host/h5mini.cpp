// Minimal HDF5 writer for exactly the subset of the format that picsp's output uses
// (src/main.cpp:21-36, 187-199, 348-353, 1142-1247): one file, the root group with scalar
// attributes (f64 / i32), first-level groups, and contiguous little-endian f64 datasets of rank 2.
//
// libhdf5 is not available in this image, so the file is emitted directly in the HDF5 File Format
// Specification's oldest, universally readable encodings: superblock version 0, version-1 object
// headers, "old style" groups (symbol-table message -> v1 B-tree -> symbol-table node + local heap),
// dataspace message v1, datatype message v1, data layout message v3 (contiguous), attribute
// message v1.  Each group uses ONE symbol-table node: the "group leaf node K" recorded in the
// superblock is chosen large enough for the fullest group, so its B-tree is a single leaf-level
// node with one child.  Raw data is streamed as datasets are written; all metadata is written by
// close().  A reader for the same subset lives in tests/h5mini.py.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "host.hpp"

namespace picsp_host {

namespace {

const uint64_t UNDEF = 0xFFFFFFFFFFFFFFFFull;

struct Buf {
    std::vector<uint8_t> b;
    void u8(uint8_t v) { b.push_back(v); }
    void u16(uint16_t v) { for (int i = 0; i < 2; i++) b.push_back((uint8_t)(v >> (8 * i))); }
    void u32(uint32_t v) { for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i))); }
    void u64(uint64_t v) { for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i))); }
    void raw(const void *p, size_t n) { const uint8_t *q = (const uint8_t *)p; b.insert(b.end(), q, q + n); }
    void zeros(size_t n) { b.insert(b.end(), n, 0); }
    void pad8() { while (b.size() % 8) b.push_back(0); }
    size_t size() const { return b.size(); }
};

// header message: type(2) size(2) flags(1) reserved(3) data(padded to 8)
void message(Buf &out, uint16_t type, const Buf &data) {
    Buf d = data; d.pad8();
    out.u16(type); out.u16((uint16_t)d.size()); out.u8(0); out.zeros(3);
    out.raw(d.b.data(), d.size());
}

// version-1 object header around a block of messages
Buf object_header(uint16_t nmsgs, const Buf &msgs) {
    Buf h;
    h.u8(1); h.u8(0); h.u16(nmsgs); h.u32(1); h.u32((uint32_t)msgs.size()); h.zeros(4);   // prefix padded to 16
    h.raw(msgs.b.data(), msgs.size());
    return h;
}

Buf datatype_f64() {
    Buf d;
    d.u8(0x11);            // version 1, class 1 (floating point)
    d.u8(0x20);            // little endian, no padding bits, mantissa normalisation = implied msb (2 << 4)
    d.u8(0x3F);            // sign bit location 63
    d.u8(0x00);
    d.u32(8);              // size in bytes
    d.u16(0); d.u16(64);   // bit offset, precision
    d.u8(52); d.u8(11);    // exponent location, size
    d.u8(0); d.u8(52);     // mantissa location, size
    d.u32(1023);           // exponent bias
    return d;
}
Buf datatype_i32() {
    Buf d;
    d.u8(0x10);            // version 1, class 0 (fixed point)
    d.u8(0x08);            // little endian, signed (bit 3)
    d.u8(0); d.u8(0);
    d.u32(4);
    d.u16(0); d.u16(32);   // bit offset, precision
    return d;
}
Buf dataspace(int rank, uint64_t d0, uint64_t d1) {
    Buf d;
    d.u8(1); d.u8((uint8_t)rank); d.u8(0); d.u8(0); d.u32(0);   // version 1, rank, flags (no max dims), reserved
    if (rank >= 1) d.u64(d0);
    if (rank >= 2) d.u64(d1);
    return d;
}

}  // namespace

bool H5Writer::open(const std::string &path, std::string *err, bool shadow) {
    shadow_ = shadow;
    fp_ = std::fopen(path.c_str(), shadow ? "r+b" : "wb");
    if (!fp_) { if (err) *err = (shadow ? "cannot open " : "cannot create ") + path; return false; }
    objs_.clear(); attrs_.clear();
    Obj root; root.name = "/"; root.is_group = true;
    objs_.push_back(root);
    std::vector<uint8_t> sb(96, 0);    // superblock placeholder (96 bytes), rewritten by close()
    eof_ = 0;
    if (shadow) eof_ = sb.size(); else append(sb);
    return true;
}

uint64_t H5Writer::reserve_dataset_f64(const std::string &abs_name, uint64_t d0, uint64_t d1) {
    while (eof_ % 8) { if (!shadow_) std::fputc(0, fp_); eof_++; }
    const uint64_t at = eof_;
    eof_ += sizeof(double) * d0 * d1;
    if (!shadow_) {
        size_t slash = abs_name.rfind('/');
        size_t parent = find_or_make_group(abs_name.substr(0, slash));
        Obj o; o.name = abs_name.substr(slash + 1); o.d0 = d0; o.d1 = d1; o.data_addr = at;
        objs_.push_back(o);
        objs_[parent].children.push_back(objs_.size() - 1);
        std::fflush(fp_);
        std::fseek(fp_, (long)eof_, SEEK_SET);        // the block is filled by write_rows() of all ranks
    }
    return at;
}

bool H5Writer::write_rows(uint64_t data_addr, uint64_t row_lo, uint64_t nrows, uint64_t d1, const double *data) {
    if (!fp_) return false;
    if (nrows == 0) return true;
    std::fflush(fp_);
    const long keep = std::ftell(fp_);
    if (std::fseek(fp_, (long)(data_addr + sizeof(double) * row_lo * d1), SEEK_SET) != 0) return false;
    const size_t n = (size_t)(nrows * d1);
    const bool ok = std::fwrite(data, sizeof(double), n, fp_) == n;
    std::fflush(fp_);
    std::fseek(fp_, keep, SEEK_SET);
    return ok;
}

uint64_t H5Writer::append(const std::vector<uint8_t> &bytes) {
    while (eof_ % 8) { std::fputc(0, fp_); eof_++; }
    uint64_t at = eof_;
    if (!bytes.empty()) std::fwrite(bytes.data(), 1, bytes.size(), fp_);
    eof_ += bytes.size();
    return at;
}

size_t H5Writer::find_or_make_group(const std::string &abs_name) {
    if (abs_name == "/" || abs_name.empty()) return 0;
    std::string name = abs_name[0] == '/' ? abs_name.substr(1) : abs_name;
    for (size_t c : objs_[0].children)
        if (objs_[c].is_group && objs_[c].name == name) return c;
    Obj g; g.name = name; g.is_group = true;
    objs_.push_back(g);
    objs_[0].children.push_back(objs_.size() - 1);
    return objs_.size() - 1;
}

void H5Writer::create_group(const std::string &abs_name) { find_or_make_group(abs_name); }

void H5Writer::write_dataset_f64(const std::string &abs_name, const double *data, uint64_t d0, uint64_t d1) {
    // "/group/name" (picsp only ever writes one level below a first-level group)
    size_t slash = abs_name.rfind('/');
    if (shadow_) { reserve_dataset_f64(abs_name, d0, d1); return; }      // rank 0 writes this one; keep the offsets in step
    size_t parent = find_or_make_group(abs_name.substr(0, slash));
    Obj o; o.name = abs_name.substr(slash + 1); o.d0 = d0; o.d1 = d1;
    while (eof_ % 8) { std::fputc(0, fp_); eof_++; }
    o.data_addr = eof_;
    std::fwrite(data, sizeof(double), (size_t)(d0 * d1), fp_);
    eof_ += sizeof(double) * d0 * d1;
    objs_.push_back(o);
    objs_[parent].children.push_back(objs_.size() - 1);
}

void H5Writer::write_attr_f64(const std::string &name, double v) { Attr a{name, false, v, 0}; attrs_.push_back(a); }
void H5Writer::write_attr_i32(const std::string &name, int32_t v) { Attr a{name, true, 0.0, v}; attrs_.push_back(a); }

// writes heap, symbol-table node, B-tree node and object header of one group; returns the header address.
// children must already have header_addr set.
uint64_t H5Writer::write_group(size_t idx, uint16_t leaf_k) {
    Obj &g = objs_[idx];
    std::vector<size_t> kids = g.children;
    std::sort(kids.begin(), kids.end(), [&](size_t a, size_t b) { return objs_[a].name < objs_[b].name; });

    // local heap data segment: offset 0 = "" then the names, each NUL-terminated and padded to 8
    Buf seg; seg.zeros(8);
    std::vector<uint64_t> name_off;
    for (size_t k : kids) {
        name_off.push_back(seg.size());
        seg.raw(objs_[k].name.c_str(), objs_[k].name.size() + 1);
        seg.pad8();
    }
    uint64_t seg_addr = append(seg.b);
    Buf heap;
    heap.raw("HEAP", 4); heap.u8(0); heap.zeros(3);
    heap.u64(seg.size());      // data segment size
    heap.u64(1);               // head of free list: 1 == H5HL_FREE_NULL (no free block)
    heap.u64(seg_addr);
    uint64_t heap_addr = append(heap.b);

    // symbol table node with 2*leaf_k entry slots
    Buf snod;
    snod.raw("SNOD", 4); snod.u8(1); snod.u8(0); snod.u16((uint16_t)kids.size());
    for (size_t n = 0; n < kids.size(); n++) {
        const Obj &c = objs_[kids[n]];
        snod.u64(name_off[n]); snod.u64(c.header_addr);
        snod.u32(0); snod.u32(0);          // cache type 0 (nothing cached), reserved
        snod.zeros(16);                    // scratch pad
    }
    snod.zeros((size_t)(2 * leaf_k - kids.size()) * 40);
    uint64_t snod_addr = append(snod.b);

    // v1 B-tree, group node, level 0, one child; capacity 2K children with K = 16
    const int K = 16;
    Buf bt;
    bt.raw("TREE", 4); bt.u8(0); bt.u8(0); bt.u16(kids.empty() ? 0 : 1);
    bt.u64(UNDEF); bt.u64(UNDEF);
    bt.u64(0);                                            // key 0: the empty string
    bt.u64(kids.empty() ? UNDEF : snod_addr);             // child 0
    bt.u64(kids.empty() ? 0 : name_off.back());           // key 1: the largest name in child 0
    bt.zeros((size_t)(2 * K - 1) * 16);                   // unused child/key pairs
    uint64_t bt_addr = append(bt.b);

    // object header: symbol table message (+ the root's attributes)
    Buf msgs; uint16_t nm = 0;
    { Buf st; st.u64(bt_addr); st.u64(heap_addr); message(msgs, 0x0011, st); nm++; }
    if (idx == 0) {
        for (const Attr &a : attrs_) {
            Buf dt = a.is_int ? datatype_i32() : datatype_f64();
            Buf ds = dataspace(0, 0, 0);
            Buf m;
            m.u8(1); m.u8(0);
            m.u16((uint16_t)(a.name.size() + 1)); m.u16((uint16_t)dt.size()); m.u16((uint16_t)ds.size());
            m.raw(a.name.c_str(), a.name.size() + 1); m.pad8();
            m.raw(dt.b.data(), dt.size()); m.pad8();
            m.raw(ds.b.data(), ds.size()); m.pad8();
            if (a.is_int) m.u32((uint32_t)a.i); else m.raw(&a.f, 8);
            message(msgs, 0x000C, m); nm++;
        }
    }
    { Buf nil; nil.zeros(8); message(msgs, 0x0000, nil); nm++; }    // NIL message: keeps chunk 0 above the minimum size
    Buf oh = object_header(nm, msgs);
    g.header_addr = append(oh.b);
    // remember B-tree / heap for the superblock's root entry
    g.d0 = bt_addr; g.d1 = heap_addr;
    return g.header_addr;
}

bool H5Writer::close(std::string *err) {
    if (!fp_) { if (err) *err = "file not open"; return false; }
    if (shadow_) { bool ok = std::fclose(fp_) == 0; fp_ = nullptr; return ok; }      // the metadata belongs to rank 0
    // dataset object headers
    for (Obj &o : objs_) {
        if (o.is_group) continue;
        Buf msgs;
        message(msgs, 0x0001, dataspace(2, o.d0, o.d1));
        message(msgs, 0x0003, datatype_f64());
        { Buf f; f.u8(2); f.u8(2); f.u8(2); f.u8(0); message(msgs, 0x0005, f); }   // fill value v2: late alloc, write if set, undefined
        { Buf l; l.u8(3); l.u8(1); l.u64(o.data_addr); l.u64(8 * o.d0 * o.d1); message(msgs, 0x0008, l); }   // layout v3, contiguous
        Buf oh = object_header(4, msgs);
        o.header_addr = append(oh.b);
    }
    size_t fullest = objs_[0].children.size();
    for (size_t c : objs_[0].children) fullest = std::max(fullest, objs_[c].children.size());
    uint16_t leaf_k = (uint16_t)std::max<size_t>(4, (fullest + 1) / 2);
    if (fullest > 65000) { if (err) *err = "too many objects in one group for the single-node layout"; return false; }
    for (size_t c : objs_[0].children) write_group(c, leaf_k);
    write_group(0, leaf_k);
    while (eof_ % 8) { std::fputc(0, fp_); eof_++; }

    Buf sb;
    const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    sb.raw(sig, 8);
    sb.u8(0); sb.u8(0); sb.u8(0); sb.u8(0);     // superblock, free-space, root-entry versions; reserved
    sb.u8(0); sb.u8(8); sb.u8(8); sb.u8(0);     // shared-header version; size of offsets, of lengths; reserved
    sb.u16(leaf_k); sb.u16(16);                 // group leaf node K, group internal node K
    sb.u32(0);                                  // file consistency flags
    sb.u64(0); sb.u64(UNDEF); sb.u64(eof_); sb.u64(UNDEF);   // base, free-space info, end of file, driver info
    sb.u64(0); sb.u64(objs_[0].header_addr); sb.u32(1); sb.u32(0);   // root entry: name offset, header, cache type 1
    sb.u64(objs_[0].d0); sb.u64(objs_[0].d1);                        // scratch: B-tree, heap
    std::fseek(fp_, 0, SEEK_SET);
    std::fwrite(sb.b.data(), 1, sb.size(), fp_);
    bool ok = std::fclose(fp_) == 0;
    fp_ = nullptr;
    if (!ok && err) *err = "write error";
    return ok;
}

}  // namespace picsp_host
