// Pool segments: the memory behind the block pool and how client processes map it.
//
// The reference's pool is one posix_memalign'd region that is cudaHostRegister'ed and
// ibv_reg_mr'ed (src/mempool.cpp:29-40) and reached through a NIC.  Here a segment is
//   * kSegDeviceIpc : HBM on a pool GPU.  Clients on any GPU of the NVSwitch domain map it
//     through a CUDA IPC handle and their kv_write / kv_read kernels address it directly
//     with peer loads/stores over NVLink.  The tail of the mapping holds the device
//     resident key index (8-way buckets, 32 bytes per entry, see kernels/index.cuh);
//   * kSegHostShm   : POSIX shared memory, the CPU-only plumbing backend and a host tier.
//     Clients mmap it (and cudaHostRegister it when they own a GPU).
// `rkey` of the reference becomes the segment id, `remote_addr` the byte offset.
#pragma once

#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../wire/protocol.h"
#include "fdpass.h"
#include "nvls.h"

namespace istore::fabric {

// 16 random bytes identifying this process (used to skip IPC when client and server
// share a process: a process cannot open its own IPC handle).
const uint8_t* process_uuid();

bool cuda_available();          // at least one usable CUDA device in this process
int cuda_device_count();
std::string cuda_last_error();  // description of the most recent failure in this module

// ---------------------------------------------------------------- server side
class SegmentOwner {
   public:
    ~SegmentOwner();
    SegmentOwner(const SegmentOwner&) = delete;
    SegmentOwner& operator=(const SegmentOwner&) = delete;

    static std::unique_ptr<SegmentOwner> create_host(uint32_t id, size_t bytes, uint32_t granule,
                                                     int port, std::string* err);
    static std::unique_ptr<SegmentOwner> create_device(uint32_t id, int device, size_t bytes,
                                                       uint32_t granule, size_t index_slots,
                                                       std::string* err);
    // NVLS-replicated region over `devices` (needs multicast support); the VMM handles are
    // handed to client processes through a unix socket named after `port`.
    static std::unique_ptr<SegmentOwner> create_replica(uint32_t id, const std::vector<int>& devices,
                                                        size_t bytes, uint32_t granule, int port,
                                                        std::string* err);
    const SegmentInfo& info() const { return info_; }
    void* base() const { return base_; }
    // Base through which the owner can READ and WRITE block bytes itself (checkpointing): the
    // replica of the first GPU for an NVLS segment (its base() is the write-only multicast VA).
    void* rw_base() const {
        return group_ ? reinterpret_cast<void*>(group_->uc_ptr(0)) : base_;
    }
    int rw_device() const { return group_ ? group_->devices()[0] : info_.device; }
    // Zero the device index (purge).  No-op for host segments.
    void clear_index();

   private:
    SegmentOwner() = default;
    SegmentInfo info_{};
    void* base_ = nullptr;
    int shm_fd_ = -1;
    std::string shm_name_;
    std::shared_ptr<NvlsGroup> group_;
    std::unique_ptr<FdServer> fd_server_;
};

// ---------------------------------------------------------------- client side
// One mapping of a segment into this process, usable from `device` (-1: host only).
struct Mapping {
    SegmentInfo info{};
    int device = -1;
    uint8_t* host_ptr = nullptr;  // CPU-addressable base (host segments only)
    uint8_t* dev_ptr = nullptr;   // GPU-addressable base for kernels / cudaMemcpy on `device`
    uint8_t* mc_ptr = nullptr;    // replica segments: multicast address (writes go here)
    bool ipc_opened = false;
    bool host_registered = false;
    std::shared_ptr<NvlsImport> nvls;  // replica segment of another process
    ~Mapping();
};

// Process-wide cache: an IPC handle can be opened only once per context, and several
// connections of one process share the mapping.
std::shared_ptr<Mapping> map_segment(const SegmentInfo& info, int device, std::string* err);

}  // namespace istore::fabric
