#include "messages.h"

namespace istore {

const char* op_name(char op) {
    switch (op) {
        case kOpLocalRead: return "LOCAL_READ";
        case kOpLocalWrite: return "LOCAL_WRITE";
        case kOpSync: return "SYNC";
        case kOpExchange: return "EXCHANGE";
        case kOpAllocate: return "ALLOCATE";
        case kOpReadLookup: return "READ_LOOKUP";
        case kOpCommit: return "COMMIT";
        case kOpCheckExist: return "CHECK_EXIST";
        case kOpMatchLastIdx: return "MATCH_LAST_INDEX";
        case kOpPoolMap: return "POOL_MAP";
        case kOpStageCommit: return "STAGE_COMMIT";
        case kOpTouch: return "TOUCH";
        default: return "UNKNOWN";
    }
}

bool op_known(char op) { return std::strcmp(op_name(op), "UNKNOWN") != 0; }
bool op_has_body(char op) { return op != kOpSync; }

namespace {
constexpr fb::voffset_t kRmKeys = 4, kRmBlockSize = 6, kRmRkey = 8, kRmAddrs = 10, kRmOp = 12,
                        kRmHint = 14;
constexpr fb::voffset_t kArBlocks = 4;
constexpr fb::voffset_t kLmDevice = 4, kLmIpc = 6, kLmBlockSize = 8, kLmBlocks = 10;
constexpr fb::voffset_t kBlKey = 4, kBlOffset = 6;
constexpr fb::voffset_t kGmKeys = 4;

fb::uoffset_t build_keys(fb::Builder& b, const std::vector<std::string_view>& keys) {
    std::vector<fb::uoffset_t> offs(keys.size());
    for (size_t i = 0; i < keys.size(); ++i) offs[i] = b.create_string(keys[i]);
    return b.create_offset_vector(offs.data(), offs.size());
}
}  // namespace

void encode_remote_meta(fb::Builder& b, const std::vector<std::string_view>& keys,
                        int32_t block_size, uint32_t rkey, const uint64_t* addrs, size_t naddrs,
                        char op, int32_t hint) {
    const fb::uoffset_t keys_off = keys.empty() ? 0 : build_keys(b, keys);
    const fb::uoffset_t addrs_off = naddrs ? b.create_vector<uint64_t>(addrs, naddrs) : 0;
    b.start_table();
    b.add_offset(kRmAddrs, addrs_off);
    b.add_scalar<uint32_t>(kRmRkey, rkey, 0);
    b.add_scalar<int32_t>(kRmBlockSize, block_size, 0);
    b.add_offset(kRmKeys, keys_off);
    b.add_scalar<int32_t>(kRmHint, hint + 1, 0);
    b.add_scalar<int8_t>(kRmOp, static_cast<int8_t>(op), 0);
    b.finish(b.end_table());
}

void encode_allocate_response(fb::Builder& b, const RemoteBlock* blocks, size_t n) {
    // an empty vector is still emitted so that readers see "0 blocks", not "absent"
    const fb::uoffset_t v = b.create_struct_vector(blocks, n, sizeof(RemoteBlock), 8);
    b.start_table();
    b.add_offset(kArBlocks, v);
    b.finish(b.end_table());
}

void encode_local_meta(fb::Builder& b, int32_t device, std::string_view ipc_handle,
                       int32_t block_size, const std::vector<LocalBlock>& blocks) {
    std::vector<fb::uoffset_t> offs(blocks.size());
    for (size_t i = 0; i < blocks.size(); ++i) {
        const fb::uoffset_t key = b.create_string(blocks[i].key);
        b.start_table();
        b.add_scalar<uint64_t>(kBlOffset, blocks[i].offset, 0);
        b.add_offset(kBlKey, key);
        offs[i] = b.end_table();
    }
    const fb::uoffset_t blocks_off = b.create_offset_vector(offs.data(), offs.size());
    const fb::uoffset_t ipc_off =
        ipc_handle.empty()
            ? 0
            : b.create_vector<uint8_t>(reinterpret_cast<const uint8_t*>(ipc_handle.data()),
                                       ipc_handle.size());
    b.start_table();
    b.add_offset(kLmBlocks, blocks_off);
    b.add_scalar<int32_t>(kLmBlockSize, block_size, 0);
    b.add_offset(kLmIpc, ipc_off);
    b.add_scalar<int32_t>(kLmDevice, device, 0);
    b.finish(b.end_table());
}

void encode_match_request(fb::Builder& b, const std::vector<std::string_view>& keys) {
    const fb::uoffset_t keys_off = build_keys(b, keys);
    b.start_table();
    b.add_offset(kGmKeys, keys_off);
    b.finish(b.end_table());
}

RemoteMetaRequest decode_remote_meta(const void* data, size_t len) {
    fb::Buf buf(data, len);
    fb::Table t = fb::Table::root(&buf);
    RemoteMetaRequest r;
    fb::OffsetVec keys = t.offset_vec(kRmKeys);
    r.keys.reserve(keys.size());
    for (uint32_t i = 0; i < keys.size(); ++i) r.keys.push_back(keys.str(i));
    r.block_size = t.scalar<int32_t>(kRmBlockSize, 0);
    r.rkey = t.scalar<uint32_t>(kRmRkey, 0);
    fb::ScalarVec<uint64_t> addrs = t.vec<uint64_t>(kRmAddrs);
    r.remote_addrs.resize(addrs.size());
    for (uint32_t i = 0; i < addrs.size(); ++i) r.remote_addrs[i] = addrs[i];
    r.op = t.scalar<int8_t>(kRmOp, 0);
    r.hint = t.scalar<int32_t>(kRmHint, 0) - 1;
    return r;
}

std::vector<RemoteBlock> decode_allocate_response(const void* data, size_t len) {
    fb::Buf buf(data, len);
    fb::Table t = fb::Table::root(&buf);
    uint32_t n = 0;
    fb::ScalarVec<uint8_t> raw = t.struct_vec(kArBlocks, sizeof(RemoteBlock), &n);
    std::vector<RemoteBlock> out(n);
    if (n) std::memcpy(out.data(), raw.raw(), size_t(n) * sizeof(RemoteBlock));
    return out;
}

LocalMetaRequest decode_local_meta(const void* data, size_t len) {
    fb::Buf buf(data, len);
    fb::Table t = fb::Table::root(&buf);
    LocalMetaRequest r;
    r.device = t.scalar<int32_t>(kLmDevice, 0);
    fb::ScalarVec<uint8_t> ipc = t.vec<uint8_t>(kLmIpc);
    if (ipc.present())
        r.ipc_handle = std::string_view(reinterpret_cast<const char*>(ipc.raw()), ipc.size());
    r.block_size = t.scalar<int32_t>(kLmBlockSize, 0);
    fb::OffsetVec blocks = t.offset_vec(kLmBlocks);
    r.blocks.reserve(blocks.size());
    for (uint32_t i = 0; i < blocks.size(); ++i) {
        fb::Table bt = blocks.table(i);
        r.blocks.push_back(LocalBlock{bt.str(kBlKey), bt.scalar<uint64_t>(kBlOffset, 0)});
    }
    return r;
}

std::vector<std::string_view> decode_match_request(const void* data, size_t len) {
    fb::Buf buf(data, len);
    fb::Table t = fb::Table::root(&buf);
    fb::OffsetVec keys = t.offset_vec(kGmKeys);
    std::vector<std::string_view> out;
    out.reserve(keys.size());
    for (uint32_t i = 0; i < keys.size(); ++i) out.push_back(keys.str(i));
    return out;
}

size_t remote_meta_bound(const std::vector<std::string_view>& keys, size_t naddrs) {
    size_t n = 128 + naddrs * 8 + keys.size() * 4;
    for (auto k : keys) n += k.size() + 12;  // len + NUL + padding
    return (n + 7) & ~size_t(7);
}

size_t local_meta_bound(const std::vector<LocalBlock>& blocks) {
    size_t n = 256 + blocks.size() * 4;
    for (auto& b : blocks) n += b.key.size() + 12 + 32;  // string + table + vtable
    return (n + 7) & ~size_t(7);
}

}  // namespace istore
