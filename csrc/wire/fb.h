// Minimal FlatBuffers writer + verifying reader for the four control-plane schemas.
//
// There is no flatc / flatbuffers header on the build image, so the subset of the format
// that the store needs (tables, strings, vectors of scalars / offsets / inline structs)
// is implemented here from the format specification.  The four messages are those of the
// reference (src/meta_request.fbs, src/allocate_response.fbs, src/local_meta_request.fbs,
// src/get_match_last_index.fbs); field slots are listed in messages.h.
//
// Writer: builds back-to-front into a caller supplied fixed buffer (so a message can be
// serialised straight into a socket / send buffer, the role FixedBufferAllocator plays in
// the reference, src/protocol.h:95-106).  The finished message sits at the END of the
// buffer: data() .. data()+size().
// Reader: every offset is validated against the buffer length before it is followed
// (the reference never runs a Verifier on untrusted input; this reader always does).
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string_view>
#include <vector>

namespace istore::fb {

using uoffset_t = uint32_t;
using soffset_t = int32_t;
using voffset_t = uint16_t;

struct Overflow : std::runtime_error {
    Overflow() : std::runtime_error("flatbuffer builder: fixed buffer overflow") {}
};

class Builder {
   public:
    // `buf` must be 8-byte aligned and `cap` a multiple of 8.
    Builder(void* buf, size_t cap) : buf_(static_cast<uint8_t*>(buf)), cap_(cap) {
        if ((reinterpret_cast<uintptr_t>(buf) & 7) || (cap & 7))
            throw std::invalid_argument("flatbuffer builder: buffer must be 8-byte aligned");
    }

    size_t size() const { return used_; }
    const uint8_t* data() const { return buf_ + cap_ - used_; }
    uint8_t* data() { return buf_ + cap_ - used_; }

    uoffset_t create_string(std::string_view s) {
        pre_align(s.size() + 1, sizeof(uoffset_t));
        push_zero(1);
        push_bytes(s.data(), s.size());
        push_scalar<uoffset_t>(static_cast<uoffset_t>(s.size()));
        return static_cast<uoffset_t>(used_);
    }

    template <typename T>
    uoffset_t create_vector(const T* v, size_t n) {
        start_vector(n * sizeof(T), sizeof(T));
        push_bytes(v, n * sizeof(T));
        push_scalar<uoffset_t>(static_cast<uoffset_t>(n));
        return static_cast<uoffset_t>(used_);
    }

    // vector of inline structs (elem_size bytes each, `align` alignment)
    uoffset_t create_struct_vector(const void* v, size_t n, size_t elem_size, size_t align) {
        start_vector(n * elem_size, align);
        push_bytes(v, n * elem_size);
        push_scalar<uoffset_t>(static_cast<uoffset_t>(n));
        return static_cast<uoffset_t>(used_);
    }

    // vector of offsets to previously created objects
    uoffset_t create_offset_vector(const uoffset_t* offs, size_t n) {
        start_vector(n * sizeof(uoffset_t), sizeof(uoffset_t));
        for (size_t i = n; i > 0; --i) push_scalar<uoffset_t>(refer_to(offs[i - 1]));
        push_scalar<uoffset_t>(static_cast<uoffset_t>(n));
        return static_cast<uoffset_t>(used_);
    }

    void start_table() {
        nfields_ = 0;
        table_start_ = used_;
    }
    template <typename T>
    void add_scalar(voffset_t slot, T v, T def) {
        if (v == def) return;  // defaults are not stored
        align(sizeof(T));
        push_scalar<T>(v);
        track(slot);
    }
    void add_offset(voffset_t slot, uoffset_t off) {
        if (!off) return;
        align(sizeof(uoffset_t));
        push_scalar<uoffset_t>(refer_to(off));
        track(slot);
    }
    uoffset_t end_table() {
        align(sizeof(soffset_t));
        push_scalar<soffset_t>(0);  // patched below
        const size_t table_off = used_;
        voffset_t max_slot = 0;
        for (int i = 0; i < nfields_; ++i) max_slot = std::max(max_slot, fields_[i].slot);
        const size_t vt_bytes = std::max<size_t>(max_slot + sizeof(voffset_t), 4);
        // vtable sits right below the table
        push_zero(vt_bytes);
        uint8_t* vt = data();
        write_le<voffset_t>(vt, static_cast<voffset_t>(vt_bytes));
        write_le<voffset_t>(vt + 2, static_cast<voffset_t>(table_off - table_start_));
        for (int i = 0; i < nfields_; ++i)
            write_le<voffset_t>(vt + fields_[i].slot,
                                static_cast<voffset_t>(table_off - fields_[i].off));
        const size_t vt_off = used_;
        write_le<soffset_t>(buf_ + cap_ - table_off, static_cast<soffset_t>(vt_off - table_off));
        return static_cast<uoffset_t>(table_off);
    }

    void finish(uoffset_t root) {
        pre_align(sizeof(uoffset_t), min_align_);
        push_scalar<uoffset_t>(refer_to(root));
    }

   private:
    struct Field {
        size_t off;
        voffset_t slot;
    };
    template <typename T>
    static void write_le(uint8_t* p, T v) {
        std::memcpy(p, &v, sizeof(T));  // x86-64 / aarch64-le only
    }
    void grow(size_t n) {
        if (used_ + n > cap_) throw Overflow();
        used_ += n;
    }
    void push_zero(size_t n) {
        grow(n);
        std::memset(data(), 0, n);
    }
    void push_bytes(const void* p, size_t n) {
        grow(n);
        if (n) std::memcpy(data(), p, n);
    }
    template <typename T>
    void push_scalar(T v) {
        grow(sizeof(T));
        write_le<T>(data(), v);
    }
    void align(size_t a) {
        if (a > min_align_) min_align_ = a;
        push_zero((~used_ + 1) & (a - 1));
    }
    // make sure that after writing `len` more bytes the cursor is `a`-aligned
    void pre_align(size_t len, size_t a) {
        if (a > min_align_) min_align_ = a;
        push_zero((~(used_ + len) + 1) & (a - 1));
    }
    void start_vector(size_t bytes, size_t elem_align) {
        pre_align(bytes, sizeof(uoffset_t));
        pre_align(bytes, elem_align);
    }
    uoffset_t refer_to(uoffset_t off) {
        align(sizeof(uoffset_t));
        return static_cast<uoffset_t>(used_ - off + sizeof(uoffset_t));
    }
    void track(voffset_t slot) {
        if (nfields_ >= kMaxFields) throw std::logic_error("flatbuffer builder: too many fields");
        fields_[nfields_++] = Field{used_, slot};
    }

    static constexpr int kMaxFields = 16;
    uint8_t* buf_;
    size_t cap_;
    size_t used_ = 0;
    size_t min_align_ = 1;
    size_t table_start_ = 0;
    Field fields_[kMaxFields];
    int nfields_ = 0;
};

// ---------------------------------------------------------------- reader

struct Malformed : std::runtime_error {
    explicit Malformed(const char* what) : std::runtime_error(what) {}
};

class Buf {
   public:
    Buf(const void* p, size_t n) : p_(static_cast<const uint8_t*>(p)), n_(n) {}
    size_t size() const { return n_; }
    const uint8_t* at(size_t off, size_t len) const {
        if (off > n_ || len > n_ - off) throw Malformed("flatbuffer: offset out of bounds");
        return p_ + off;
    }
    template <typename T>
    T read(size_t off) const {
        T v;
        std::memcpy(&v, at(off, sizeof(T)), sizeof(T));
        return v;
    }

   private:
    const uint8_t* p_;
    size_t n_;
};

template <typename T>
class ScalarVec {
   public:
    ScalarVec() = default;
    ScalarVec(const uint8_t* p, uint32_t n) : p_(p), n_(n) {}
    uint32_t size() const { return n_; }
    bool present() const { return p_ != nullptr; }
    T operator[](uint32_t i) const {
        T v;
        std::memcpy(&v, p_ + size_t(i) * sizeof(T), sizeof(T));
        return v;
    }
    const uint8_t* raw() const { return p_; }

   private:
    const uint8_t* p_ = nullptr;
    uint32_t n_ = 0;
};

class Table;

class OffsetVec {
   public:
    OffsetVec() = default;
    OffsetVec(const Buf* b, size_t pos, uint32_t n) : b_(b), pos_(pos), n_(n) {}
    uint32_t size() const { return n_; }
    bool present() const { return b_ != nullptr; }
    std::string_view str(uint32_t i) const;
    Table table(uint32_t i) const;

   private:
    size_t target(uint32_t i) const {
        const size_t loc = pos_ + size_t(i) * sizeof(uoffset_t);
        return loc + b_->read<uoffset_t>(loc);
    }
    const Buf* b_ = nullptr;
    size_t pos_ = 0;
    uint32_t n_ = 0;
};

class Table {
   public:
    Table(const Buf* b, size_t pos) : b_(b), pos_(pos) {
        const soffset_t so = b_->read<soffset_t>(pos_);
        const int64_t vt = int64_t(pos_) - so;
        if (vt < 0) throw Malformed("flatbuffer: vtable before buffer start");
        vt_ = size_t(vt);
        vt_bytes_ = b_->read<voffset_t>(vt_);
        if (vt_bytes_ < 4 || (vt_bytes_ & 1)) throw Malformed("flatbuffer: bad vtable size");
        b_->at(vt_, vt_bytes_);
        tbl_bytes_ = b_->read<voffset_t>(vt_ + 2);
        b_->at(pos_, tbl_bytes_);
    }
    static Table root(const Buf* b) {
        const uoffset_t r = b->read<uoffset_t>(0);
        return Table(b, r);
    }

    template <typename T>
    T scalar(voffset_t slot, T def) const {
        const size_t f = field(slot, sizeof(T));
        return f ? b_->read<T>(f) : def;
    }
    std::string_view str(voffset_t slot) const {
        const size_t f = field(slot, sizeof(uoffset_t));
        return f ? read_string(b_, f + b_->read<uoffset_t>(f)) : std::string_view();
    }
    bool has(voffset_t slot) const { return field(slot, 1) != 0; }
    template <typename T>
    ScalarVec<T> vec(voffset_t slot) const {
        return struct_vec_impl<T>(slot, sizeof(T));
    }
    // vector of inline structs of `elem` bytes; returned as raw bytes
    ScalarVec<uint8_t> struct_vec(voffset_t slot, size_t elem, uint32_t* count) const {
        const size_t f = field(slot, sizeof(uoffset_t));
        *count = 0;
        if (!f) return {};
        const size_t v = f + b_->read<uoffset_t>(f);
        const uint32_t n = b_->read<uoffset_t>(v);
        const uint8_t* p = b_->at(v + 4, size_t(n) * elem);
        *count = n;
        return ScalarVec<uint8_t>(p, uint32_t(size_t(n) * elem));
    }
    OffsetVec offset_vec(voffset_t slot) const {
        const size_t f = field(slot, sizeof(uoffset_t));
        if (!f) return {};
        const size_t v = f + b_->read<uoffset_t>(f);
        const uint32_t n = b_->read<uoffset_t>(v);
        b_->at(v + 4, size_t(n) * sizeof(uoffset_t));
        return OffsetVec(b_, v + 4, n);
    }

    static std::string_view read_string(const Buf* b, size_t pos) {
        const uint32_t n = b->read<uoffset_t>(pos);
        const uint8_t* p = b->at(pos + 4, size_t(n) + 1);
        if (p[n] != 0) throw Malformed("flatbuffer: string not NUL terminated");
        return std::string_view(reinterpret_cast<const char*>(p), n);
    }

   private:
    template <typename T>
    ScalarVec<T> struct_vec_impl(voffset_t slot, size_t elem) const {
        const size_t f = field(slot, sizeof(uoffset_t));
        if (!f) return {};
        const size_t v = f + b_->read<uoffset_t>(f);
        const uint32_t n = b_->read<uoffset_t>(v);
        const uint8_t* p = b_->at(v + 4, size_t(n) * elem);
        return ScalarVec<T>(p, n);
    }
    // absolute position of a field, or 0 when absent
    size_t field(voffset_t slot, size_t width) const {
        if (size_t(slot) + sizeof(voffset_t) > vt_bytes_) return 0;
        const voffset_t fo = b_->read<voffset_t>(vt_ + slot);
        if (!fo) return 0;
        if (size_t(fo) + width > tbl_bytes_) throw Malformed("flatbuffer: field outside table");
        return pos_ + fo;
    }
    const Buf* b_;
    size_t pos_;
    size_t vt_ = 0;
    voffset_t vt_bytes_ = 0;
    voffset_t tbl_bytes_ = 0;
};

inline std::string_view OffsetVec::str(uint32_t i) const {
    if (i >= n_) throw Malformed("flatbuffer: vector index out of range");
    return Table::read_string(b_, target(i));
}
inline Table OffsetVec::table(uint32_t i) const {
    if (i >= n_) throw Malformed("flatbuffer: vector index out of range");
    return Table(b_, target(i));
}

}  // namespace istore::fb
