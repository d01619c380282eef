// Typed encode/decode of the four control-plane messages on top of fb.h.
//
// Schemas (field -> vtable slot), matching the reference's .fbs files:
//   RemoteMetaRequest   keys:[string]=4 block_size:int=6 rkey:uint=8 remote_addrs:[ulong]=10
//                       op:byte=12                       (reference: src/meta_request.fbs)
//                       + extension  hint:int=14  (pool/device placement hint, -1 = any)
//   RdmaAllocateResponse blocks:[RemoteBlock]=4, RemoteBlock = inline 16-byte struct
//                                                        (reference: src/allocate_response.fbs)
//   LocalMetaRequest    device:int=4 ipc_handle:[ubyte]=6 block_size:int=8 blocks:[Block]=10
//   Block               key:string=4 offset:ulong=6      (reference: src/local_meta_request.fbs)
//   GetMatchLastIndexRequest keys:[string]=4             (reference: src/get_match_last_index.fbs)
// Extension fields use new, higher vtable slots, which FlatBuffers readers that do not know
// them simply ignore.
#pragma once

#include <string>
#include <string_view>
#include <vector>

#include "fb.h"
#include "protocol.h"

namespace istore {

struct RemoteMetaRequest {
    std::vector<std::string_view> keys;
    int32_t block_size = 0;
    uint32_t rkey = 0;
    std::vector<uint64_t> remote_addrs;
    int8_t op = 0;
    int32_t hint = 0;  // stored +1 on the wire so that "absent" (0) decodes as -1 = any
};

struct LocalBlock {
    std::string_view key;
    uint64_t offset = 0;
};

struct LocalMetaRequest {
    int32_t device = 0;
    std::string_view ipc_handle;  // 64 raw bytes or empty
    int32_t block_size = 0;
    std::vector<LocalBlock> blocks;
};

// Each encoder serialises into `b` and returns after finish(); payload = b.data(), b.size().
void encode_remote_meta(fb::Builder& b, const std::vector<std::string_view>& keys,
                        int32_t block_size, uint32_t rkey, const uint64_t* addrs, size_t naddrs,
                        char op, int32_t hint = -1);
void encode_allocate_response(fb::Builder& b, const RemoteBlock* blocks, size_t n);
void encode_local_meta(fb::Builder& b, int32_t device, std::string_view ipc_handle,
                       int32_t block_size, const std::vector<LocalBlock>& blocks);
void encode_match_request(fb::Builder& b, const std::vector<std::string_view>& keys);

// Decoders validate every offset; they throw fb::Malformed on bad input.  The returned
// string_views point into `data`.
RemoteMetaRequest decode_remote_meta(const void* data, size_t len);
std::vector<RemoteBlock> decode_allocate_response(const void* data, size_t len);
LocalMetaRequest decode_local_meta(const void* data, size_t len);
std::vector<std::string_view> decode_match_request(const void* data, size_t len);

// Upper bound of the encoded size, used to size send buffers.
size_t remote_meta_bound(const std::vector<std::string_view>& keys, size_t naddrs);
size_t local_meta_bound(const std::vector<LocalBlock>& blocks);

}  // namespace istore
